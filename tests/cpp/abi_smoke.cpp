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
        // the CAT detector mirror: Detector::new + process_frame (the reference's bench shape, bench.rs:8-26) against the stage calls
        {
            const size_t W = 96, H = 72;
            std::vector<uint8_t> rgb(W * H * 3);
            for (size_t y = 0; y < H; y++)
                for (size_t x = 0; x < W; x++) {
                    const bool black = (x / 12 + y / 12) % 2 == 0;
                    for (int c = 0; c < 3; c++) rgb[(y * W + x) * 3 + c] = (uint8_t)(black ? 20 + (x * 7 + y * 3 + c) % 9 : 220 - (x + y + c) % 11);
                }
            chalkydri::cat::Detector cat(W, H, {});
            cat.process_frame(rgb.data(), rgb.size());
            const auto pts = cat.points(), lns = cat.lines();
            const auto map = cat.buf;
            chalkydri::cat::Detector stagewise(cat);                 // Clone: a fresh detector of the same size
            stagewise.calc_otsu(rgb.data());
            stagewise.detect_corners();
            stagewise.check_edges();
            const auto uf = stagewise.connected_components();
            std::printf("CAT: %zu corners, %zu lines, root of pixel 0 has %zu pixels\n", pts.size() / 2, lns.size() / 4, uf.get_size(uf.find(0)));
            if (pts != stagewise.points() || lns != stagewise.lines() || map != stagewise.buf || pts.empty()) return 6;
            bool threw = false;
            try { cat.process_frame(rgb.data(), rgb.size() - 3); } catch (const std::invalid_argument &) { threw = true; }
            if (!threw) return 7;
            // the decode path on the same (tag-free) frame: runs every stage, finds nothing
            const auto tags = cat.detect_tags(rgb.data(), rgb.size());
            std::printf("CAT decode: %zu tags on a checkerboard\n", tags.size());
            if (!tags.empty()) return 12;
        }
        // Family::from_str, SqPnP: Clone
        if (chalkydri::family_from_str("tag36h11") != chalkydri::Family::Tag36h11) return 8;
        chalkydri::SqPnP solver;
        solver.max_iter(10).tolerance(1e-9);
        chalkydri::SqPnP copy(solver);
        // one process, one context per listed GPU (the same GPU twice here): lists land in slices of one array
        {
            std::vector<uint8_t> frames((size_t)5 * 640 * 480, 128);
            std::vector<cb_detection> lists((size_t)5 * 16);
            std::vector<int32_t> counts(5, -1);
            chalkydri::DetectorPool pool({0, 0}, 640, 480, 2, 16);
            pool.detect(frames.data(), 640, 480, 640, (size_t)640 * 480, 5, lists.data(), counts.data());
            for (int c : counts) if (c != 0) return 9;
            std::printf("pool of %d contexts: ok\n", pool.size());
        }
        return d.empty() && !r ? 0 : 2;
    } catch (const chalkydri::Error &e) {
        std::printf("no usable GPU: %s\n", e.what());
        return cb_device_count() == 0 ? 0 : 3;    // loud failure is the expected behaviour on a CPU-only host
    }
}
