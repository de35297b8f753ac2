// Compiles the C++ host mirror against the C ABI and exercises the no-GPU error path (CPU test) or one detection (GPU).
#include <cstdio>
#include <cstring>
#include <vector>

#include "chalkydri_b200.hpp"

int main(int argc, char **argv)
{
    std::printf("version %s devices %d\n", cb_version(), cb_device_count());
    try {
        auto det = chalkydri::DetectorBuilder::default_().add_family_bits("tag36h11", 3).capacity(640, 480, 1, 16).build();
        std::vector<uint8_t> img(640 * 480, 128);
        chalkydri::Image im{img.data(), 640, 480, 640};
        auto d = det.detect(im);
        std::printf("detections on a flat frame: %zu\n", d.size());
        return d.empty() ? 0 : 2;
    } catch (const chalkydri::Error &e) {
        std::printf("no usable GPU: %s\n", e.what());
        return cb_device_count() == 0 ? 0 : 3;    // loud failure is the expected behaviour on a CPU-only host
    }
}
