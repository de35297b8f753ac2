// Compiles the C++ host mirror against the C ABI and exercises the no-GPU error path (CPU test) or one detection (GPU).
#include <cstdio>
#include <cstring>
#include <vector>

#include "chalkydri_b200.hpp"

int main()
{
    std::printf("version %s devices %d\n", cb_version(), cb_device_count());
    try {
        auto det = chalkydri::DetectorBuilder::default_().add_family_bits("tag36h11", 3).capacity(640, 480, 1, 16).build();
        std::vector<uint8_t> img(640 * 480, 128);
        chalkydri::Image im{img.data(), 640, 480, 640};
        auto d = det.detect(im);
        std::printf("detections on a flat frame: %zu\n", d.size());
        // streaming form: two batches in flight, collected oldest first
        std::vector<cb_detection> lists(16);
        int32_t n1 = -1, n2 = -1;
        det.submit(img.data(), 640, 480, 640, (size_t)640 * 480, 1);
        det.submit(img.data(), 640, 480, 640, (size_t)640 * 480, 1);
        det.collect(lists.data(), &n1);
        det.collect(lists.data(), &n2);
        std::printf("streaming: %d + %d detections, %d pending\n", n1, n2, det.pending());
        if (n1 != 0 || n2 != 0 || det.pending() != 0) return 4;
        // the task mirror on the fused device call: a flat frame publishes the heartbeat once
        int published = 0;
        chalkydri::Comm comm;
        comm.gyro_angle = [] { return std::optional<double>(0.25); };
        comm.publish = [&](uint8_t, uint8_t tags, uint64_t, const chalkydri::RobotPose &, const chalkydri::VisionUncertainty &) { published += tags == 0; };
        cb_iso3 tag{};
        tag.q[0] = 1.0;
        chalkydri::AprilTags task(chalkydri::DetectorBuilder::default_().add_family_bits("tag36h11", 3).capacity(640, 480, 1, 16), {{1, tag}},
                                  {500, 500, 320, 240, 0, 0, 0, 0, 0}, std::nullopt, 7, comm);
        const auto r = task.process(1000000, 990000, im);
        std::printf("task on a flat frame: pose %s, heartbeats %d\n", r ? "some" : "none", published);
        // the task's streaming form: two frames in flight; 20 ms later the heartbeat is due again, the second one is not
        task.submit(1010000, im);
        task.submit(1011000, im);
        const auto r1 = task.collect(1020000), r2 = task.collect(1021000);
        std::printf("task, streaming: %s %s, heartbeats %d\n", r1 ? "some" : "none", r2 ? "some" : "none", published);
        if (r1 || r2 || published != 2) return 5;
        return d.empty() && !r ? 0 : 2;
    } catch (const chalkydri::Error &e) {
        std::printf("no usable GPU: %s\n", e.what());
        return cb_device_count() == 0 ? 0 : 3;    // loud failure is the expected behaviour on a CPU-only host
    }
}
