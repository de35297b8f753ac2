// chalkydri_b200.hpp -- header-only C++ host mirror of the reference interfaces, above the C ABI (chalkydri_b200.h).
//
// The reference's host code is Rust; where its toolchain is absent the host side above the C ABI is C++ (this file) and
// Python (chalkydri_b200/*.py).  Names, argument meaning and failure behaviour follow
//   apriltag::{DetectorBuilder, Detector, Detection}   /root/reference/crates/apriltags/src/lib.rs:19,258-261,301-314
//   chalkydri_sqpnp::SqPnP                              /root/reference/crates/chalkydri_sqpnp/src/lib.rs:183-304,430-461
//   the Copper task AprilTags (new / process)           /root/reference/crates/apriltags/src/lib.rs:166-379
//   chalkydri-apriltags ("CAT") Detector                 /root/reference/crates/chalkydri-apriltags/src/lib.rs:142-181,265-287,501-549
// Config errors throw (the reference unwrap()s / panics), the solver returns std::optional (the reference: Option).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "chalkydri_b200.h"

namespace chalkydri {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error("chalkydri_b200 error " + std::to_string(c) + ": " + what), code(c) {}
};

// image_u8_t view {buf, width, height, stride}: borrows the caller's pixels (image_from_cuimage, lib.rs:197-213)
struct Image {
    const uint8_t *buf;
    int width, height, stride;
};

class Detection {
  public:
    explicit Detection(const cb_detection &d) : d_(d) {}
    size_t id() const { return (size_t)d_.id; }
    size_t hamming() const { return (size_t)d_.hamming; }
    float decision_margin() const { return d_.decision_margin; }
    std::array<std::array<double, 2>, 4> corners() const
    {
        std::array<std::array<double, 2>, 4> c;
        for (int i = 0; i < 4; i++) c[i] = {d_.p[i][0], d_.p[i][1]};
        return c;
    }
    std::array<double, 2> center() const { return {d_.c[0], d_.c[1]}; }
    const double *homography() const { return d_.H; }   // row-major 3x3
    const cb_detection &raw() const { return d_; }

  private:
    cb_detection d_;
};

class Detector {
  public:
    Detector(int device, int max_w, int max_h, int max_batch, int max_dets, int bits_corrected) : max_dets_(max_dets), max_batch_(max_batch)
    {
        ctx_ = cb_create(device, max_w, max_h, max_batch, max_dets);
        if (!ctx_) throw Error(CB_ERR_CUDA, cb_last_error(nullptr));
        check(cb_set_family_tag36h11(ctx_, bits_corrected));
    }
    Detector(const Detector &) = delete;
    Detector &operator=(const Detector &) = delete;
    ~Detector() { cb_destroy(ctx_); }

    // Detector::detect(&Image) -> Vec<Detection>   (lib.rs:301)
    std::vector<Detection> detect(const Image &im)
    {
        std::vector<cb_detection> out((size_t)max_dets_);
        int32_t n = 0;
        check(cb_detect_gray(ctx_, im.buf, im.width, im.height, im.stride, (size_t)im.stride * im.height, 1, out.data(), &n));
        std::vector<Detection> r;
        for (int i = 0; i < n; i++) r.emplace_back(out[i]);
        return r;
    }
    // batched form: frames laid out frame_stride bytes apart; out[b * max_dets + k], counts[b]
    void detect_batch(const uint8_t *frames, int w, int h, int stride, size_t frame_stride, int batch, cb_detection *out, int32_t *counts)
    {
        check(cb_detect_gray(ctx_, frames, w, h, stride, frame_stride, batch, out, counts));
    }
    // streaming form for a continuous feed (the camera loop's 4-slot host pool, gst_to_cu.rs:66,72): submit batch k+1, then
    // collect batch k -- its H2D copy runs under batch k's kernels.  At most two batches in flight; frames stay valid until collected.
    void submit(const uint8_t *frames, int w, int h, int stride, size_t frame_stride, int batch)
    {
        check(cb_detect_gray_submit(ctx_, frames, w, h, stride, frame_stride, batch));
    }
    void collect(cb_detection *out, int32_t *counts) { check(cb_detect_gray_collect(ctx_, out, counts)); }
    int pending() const { return cb_detect_gray_pending(ctx_); }
    cb_timing timing() const { cb_timing t; cb_get_timing(ctx_, &t); return t; }
    cb_ctx *ctx() { return ctx_; }
    int max_dets() const { return max_dets_; }

  private:
    void check(int rc) { if (rc != CB_OK) throw Error(rc, cb_last_error(ctx_)); }
    cb_ctx *ctx_ = nullptr;
    int max_dets_, max_batch_;
};

// apriltag::Family (the reference parses its `family` config string with Family::from_str and unwraps, lib.rs:229)
enum class Family { Tag36h11 };
inline Family family_from_str(const std::string &name)
{
    if (name == "tag36h11") return Family::Tag36h11;
    throw Error(CB_ERR_UNSUPPORTED, "unknown family " + name + " (this build carries tag36h11, the reference's FAMILY)");
}

class DetectorBuilder {
  public:
    static DetectorBuilder default_() { return DetectorBuilder(); }
    // DetectorBuilder::add_family_bits(family, bits) (lib.rs:259, 280)
    DetectorBuilder &add_family_bits(Family, size_t bits_corrected)
    {
        bits_ = (int)bits_corrected;
        return *this;
    }
    DetectorBuilder &add_family_bits(const std::string &family, size_t bits_corrected) { return add_family_bits(family_from_str(family), bits_corrected); }
    DetectorBuilder &device(int d) { device_ = d; return *this; }
    DetectorBuilder &capacity(int max_w, int max_h, int max_batch = 1, int max_dets = 64) { w_ = max_w; h_ = max_h; batch_ = max_batch; dets_ = max_dets; return *this; }
    Detector build() const
    {
        if (bits_ < 0) throw Error(CB_ERR_STATE, "no tag family added");
        return Detector(device_, w_, h_, batch_, dets_, bits_);
    }

  private:
    int bits_ = -1, device_ = 0, w_ = 1600, h_ = 1304, batch_ = 1, dets_ = 64;
};

struct RobotPoseResult {
    std::array<double, 9> rotation;   // column-major 3x3 (Rot3)
    std::array<double, 3> position, std_devs;
};

class SqPnP {
  public:
    explicit SqPnP(int device = 0)
    {
        ctx_ = cb_create(device, 8, 8, 1, 1);
        if (!ctx_) throw Error(CB_ERR_CUDA, cb_last_error(nullptr));
    }
    // Clone (lib.rs:182: #[derive(Clone, Debug, Default)]): a fresh solver with the same settings -- the scratch is per instance
    SqPnP(const SqPnP &o) : SqPnP() { max_iter_ = o.max_iter_; tol_ = o.tol_; cb_sqpnp_set(ctx_, max_iter_, tol_); }
    SqPnP &operator=(const SqPnP &o) { max_iter_ = o.max_iter_; tol_ = o.tol_; cb_sqpnp_set(ctx_, max_iter_, tol_); return *this; }
    ~SqPnP() { cb_destroy(ctx_); }
    SqPnP &max_iter(size_t n) { max_iter_ = (int)n; cb_sqpnp_set(ctx_, max_iter_, tol_); return *this; }      // lib.rs:214-217
    SqPnP &tolerance(double t) { tol_ = t; cb_sqpnp_set(ctx_, max_iter_, tol_); return *this; }               // lib.rs:219-222

    // solve_robot_pose(points_isometry, points_2d, robot_to_cam, gyro, sign_change_error) -> Option<(Rot3, Vec3, Vec3)>  (lib.rs:297-304)
    std::optional<RobotPoseResult> solve_robot_pose(const std::vector<cb_iso3> &points_isometry, const std::vector<std::array<double, 3>> &points_2d,
                                                    const cb_iso3 &robot_to_cam, double gyro, double sign_change_error)
    {
        const size_t n = points_isometry.size();
        if (n > 32) throw std::invalid_argument("SqPnP: at most 32 tags per problem (csrc/sqpnp.cuh SQ_MAX_TAGS)");   // the reference has no cap: do not answer "no pose"
        if (n * 4 < 3 || n * 4 != points_2d.size()) return std::nullopt;      // lib.rs:255-257
        const int32_t nt = (int32_t)n;
        cb_pose out;
        uint8_t ok = 0;
        int rc = cb_sqpnp_batch(ctx_, points_isometry.data(), &points_2d[0][0], &nt, (int)n, &robot_to_cam, &gyro, sign_change_error, 1, &out, &ok);
        if (rc != CB_OK) throw Error(rc, cb_last_error(ctx_));
        if (!ok) return std::nullopt;
        RobotPoseResult r;
        for (int i = 0; i < 9; i++) r.rotation[i] = out.rot[i];
        for (int i = 0; i < 3; i++) { r.position[i] = out.pos[i]; r.std_devs[i] = out.std_devs[i]; }
        return r;
    }
    static cb_iso3 create_solver_camera_transform(double fwd_m, double left_m, double up_m, double roll_deg, double pitch_deg, double yaw_deg)
    {
        cb_iso3 o;
        cb_create_solver_camera_transform(fwd_m, left_m, up_m, roll_deg, pitch_deg, yaw_deg, &o);
        return o;
    }

  private:
    cb_ctx *ctx_ = nullptr;
    int max_iter_ = 15;
    double tol_ = 1e-8;
};

// ---- whacknet/src/lib.rs:17-38 ----
struct RobotPose { double x = 0, y = 0, rot = 0; };
struct VisionUncertainty { double x = 0, y = 0, rot = 0; };

// What the task needs from whacknet::Comm (whacknet/src/lib.rs:152-178): the gyro reading and the publish call.
struct Comm {
    std::function<std::optional<double>()> gyro_angle;
    std::function<void(uint8_t cam_id, uint8_t tag_count, uint64_t ts_us, const RobotPose &, const VisionUncertainty &)> publish;
};

// Mirror of the Copper sink task `AprilTags` (crates/apriltags/src/lib.rs:166-379) on the fused device call: detect, field
// lookup, un-projection and the multi-tag SQPnP run inside cb_detect_pose_gray; this class keeps the task's publish / heartbeat
// rules (lib.rs:340-376).
class AprilTags {
  public:
    // field: tag id -> pose (field.json, field_layout.rs:18-44); calib: OpenCVModel5 {fx, fy, cx, cy, k1, k2, p1, p2, k3}
    // (lib.rs:232-233); robot_to_cam: SqPnP::create_solver_camera_transform(...) or nullopt (lib.rs:277-290)
    AprilTags(const DetectorBuilder &builder, const std::vector<std::pair<int32_t, cb_iso3>> &field, const std::array<double, 9> &calib,
              std::optional<cb_iso3> robot_to_cam, uint8_t cam_id, Comm comm)
        : det_(builder.build()), cam_id_(cam_id), comm_(std::move(comm))
    {
        std::vector<int32_t> ids;
        std::vector<cb_iso3> poses;
        for (const auto &kv : field) { ids.push_back(kv.first); poses.push_back(kv.second); }
        check(cb_set_field(det_.ctx(), ids.data(), poses.data(), (int)ids.size()));
        check(cb_set_camera(det_.ctx(), calib.data(), robot_to_cam ? &*robot_to_cam : nullptr));
    }

    // process(clock.now(), tov, image) -> the pose it published, if any
    std::optional<std::pair<RobotPose, VisionUncertainty>> process(uint64_t now_us, uint64_t frame_time_us, const Image &image)
    {
        constexpr double SIGN_FLIP_CONST = 600.0;                                 // crates/apriltags/src/lib.rs:6
        const std::optional<double> gyro = comm_.gyro_angle ? comm_.gyro_angle() : std::nullopt;
        const double g = gyro ? *gyro : std::nan("");
        std::vector<cb_detection> dets((size_t)det_.max_dets());
        int32_t count = 0, used = 0;
        cb_pose pose{};
        uint8_t ok = 0;
        // a frame so cluttered that a fixed-size device table overflows must not take the task down (upstream has no such limit):
        // it counts as "nothing detected" and the heartbeat goes out
        if (!check_or_overflow(cb_detect_pose_gray(det_.ctx(), image.buf, image.width, image.height, image.stride, (size_t)image.stride * image.height, 1,
                                                   &g, SIGN_FLIP_CONST, dets.data(), &count, &pose, &ok, &used))) { count = 0; ok = 0; }
        return publish(now_us, now_us - frame_time_us, pose, ok, count);
    }

    // continuous feed: submit(frame k+1) before collect(frame k); the frame stays valid until its collect (the camera pool
    // slot is handed back afterwards).  Same publish rules as process().
    void submit(uint64_t frame_time_us, const Image &image)
    {
        constexpr double SIGN_FLIP_CONST = 600.0;
        const std::optional<double> gyro = comm_.gyro_angle ? comm_.gyro_angle() : std::nullopt;
        const double g = gyro ? *gyro : std::nan("");
        check(cb_detect_pose_gray_submit(det_.ctx(), image.buf, image.width, image.height, image.stride, (size_t)image.stride * image.height, 1, &g,
                                         SIGN_FLIP_CONST));
        times_[(head_ + pending_++) & 1] = frame_time_us;
    }
    std::optional<std::pair<RobotPose, VisionUncertainty>> collect(uint64_t now_us)
    {
        std::vector<cb_detection> dets((size_t)det_.max_dets());
        int32_t count = 0, used = 0;
        cb_pose pose{};
        uint8_t ok = 0;
        if (!check_or_overflow(cb_detect_pose_gray_collect(det_.ctx(), dets.data(), &count, &pose, &ok, &used))) { count = 0; ok = 0; }
        const uint64_t ts = now_us - times_[head_];
        head_ ^= 1; pending_--;
        return publish(now_us, ts, pose, ok, count);
    }

  private:
    std::optional<std::pair<RobotPose, VisionUncertainty>> publish(uint64_t now_us, uint64_t ts, const cb_pose &pose, uint8_t ok, int32_t count)
    {
        if (ok) {
            cb_vision_measurement m{};
            check(cb_pack_vision_measurements(&pose, &ok, &count, &ts, cam_id_, 1, &m));
            const RobotPose rp{m.x, m.y, m.rot};
            const VisionUncertainty vu{m.std_x, m.std_y, m.std_rot};
            if (comm_.publish) comm_.publish(cam_id_, m.tag_count, ts, rp, vu);
            return std::make_pair(rp, vu);
        }
        const uint64_t now_ms = now_us / 1000;
        if (!last_time_ || now_ms - *last_time_ > 5) {                            // empty heartbeat at most every 5 ms (lib.rs:365-376)
            if (comm_.publish) comm_.publish(cam_id_, 0, ts, RobotPose{}, VisionUncertainty{});
            last_time_ = now_ms;
        }
        return std::nullopt;
    }
    void check(int rc) { if (rc != CB_OK) throw Error(rc, cb_last_error(det_.ctx())); }
    bool check_or_overflow(int rc) { if (rc == CB_ERR_OVERFLOW) { overflowed_++; return false; } check(rc); return true; }
    uint64_t overflowed_ = 0;                  // frames answered with a heartbeat because a device table overflowed
    uint64_t times_[2] = {0, 0};
    int head_ = 0, pending_ = 0;
    Detector det_;
    uint8_t cam_id_;
    Comm comm_;
    std::optional<uint64_t> last_time_;
};

// ---- the in-house CAT detector (crates/chalkydri-apriltags/src/lib.rs) ----
namespace cat {

enum Color : uint8_t { Black = 0, White = 1, Other = 2 };                        // utils.rs:1-6

// the UnionFind connected_components() returns (lib.rs:42-113): parent[] = smallest pixel index of the component
struct UnionFind {
    std::vector<uint32_t> parent, cluster_sizes;
    size_t find(size_t idx) const { return parent[idx]; }
    size_t get_size(size_t idx) const { return cluster_sizes[idx]; }
};

class Detector {
  public:
    // Detector::new(width, height, valid_tags) (lib.rs:158-181)
    Detector(size_t width, size_t height, std::vector<size_t> valid_tags = {}, int device = 0)
        : width_(width), height_(height), valid_tags_(std::move(valid_tags)), buf(width * height, (uint8_t)Black)
    {
        ctx_ = cb_create(device, 8, 8, 1, 1);
        if (!ctx_) throw Error(CB_ERR_CUDA, cb_last_error(nullptr));
    }
    Detector(const Detector &o) : Detector(o.width_, o.height_, o.valid_tags_) {}          // Clone = a fresh, empty detector (lib.rs:663-667)
    Detector &operator=(const Detector &) = delete;
    ~Detector() { cb_destroy(ctx_); if (det_ctx_) cb_destroy(det_ctx_); }

    // process_frame(&mut self, input: &[u8]) (lib.rs:265-287); panics on a wrong length (lib.rs:267)
    void process_frame(const uint8_t *input, size_t len)
    {
        if (len != width_ * height_ * 3) throw std::invalid_argument("process_frame: input must be width * height * 3 bytes of packed RGB");
        points_.resize(cap_ * 2);
        lines_.resize(cap_ * 4);
        int64_t n = 0, m = 0;
        check(cb_cat_process_frame(ctx_, input, (int)width_, (int)height_, buf.data(), points_.data(), (int64_t)cap_, &n, lines_.data(), (int64_t)cap_, &m));
        points_.resize((size_t)n * 2);
        lines_.resize((size_t)m * 4);
    }
    void calc_otsu(const uint8_t *input) { check(cb_cat_calc_otsu(ctx_, input, (int)width_, (int)height_, buf.data())); }       // lib.rs:191
    void thresh(const uint8_t *input) { check(cb_cat_thresh(ctx_, input, (int)width_, (int)height_, buf.data())); }             // lib.rs:319
    void detect_corners()                                                                                                        // lib.rs:291
    {
        points_.resize(cap_ * 2);
        int64_t n = 0;
        check(cb_cat_detect_corners(ctx_, buf.data(), (int)width_, (int)height_, points_.data(), (int64_t)cap_, &n));
        if ((size_t)n > cap_) throw Error(CB_ERR_OVERFLOW, "corner list capacity");
        points_.resize((size_t)n * 2);
    }
    void check_edges()                                                                                                           // lib.rs:480
    {
        lines_.resize(cap_ * 4);
        int64_t n = 0;
        check(cb_cat_check_edges(ctx_, buf.data(), (int)width_, (int)height_, points_.data(), (int64_t)(points_.size() / 2), lines_.data(), (int64_t)cap_, &n));
        if ((size_t)n > cap_) throw Error(CB_ERR_OVERFLOW, "line list capacity");
        lines_.resize((size_t)n * 4);
    }
    // The decode the reference intends for CAT (book/src/maintenance/apriltags.md:58-60, lib.rs:551-613): CAT's own ternary map (thresh, or
    // calc_otsu), then the C library's stages on it, in one library call.  With valid_tags given to the constructor only those ids.
    std::vector<cb_detection> detect_tags(const uint8_t *input, size_t len, bool use_otsu = false, int max_dets = 64)
    {
        if (len != width_ * height_ * 3) throw std::invalid_argument("detect_tags: input must be width * height * 3 bytes of packed RGB");
        if (!det_ctx_) {                                       // the decode stages run undecimated: capacity of twice the frame size
            det_ctx_ = cb_create(0, 2 * (int)width_, 2 * (int)height_, 1, max_dets);
            if (!det_ctx_) throw Error(CB_ERR_CUDA, cb_last_error(nullptr));
            det_cap_ = max_dets;
            const int rc = cb_set_family_tag36h11(det_ctx_, 3);
            if (rc != CB_OK) throw Error(rc, cb_last_error(det_ctx_));
        }
        std::vector<cb_detection> out((size_t)det_cap_);
        int32_t n = 0;
        const int rc = cb_cat_detect_tags(det_ctx_, input, (int)width_, (int)height_, use_otsu ? 1 : 0, out.data(), &n);
        if (rc != CB_OK) throw Error(rc, cb_last_error(det_ctx_));
        out.resize((size_t)n);
        if (!valid_tags_.empty())
            out.erase(std::remove_if(out.begin(), out.end(), [&](const cb_detection &d) {
                          return std::find(valid_tags_.begin(), valid_tags_.end(), (size_t)d.id) == valid_tags_.end(); }), out.end());
        return out;
    }
    UnionFind connected_components() const                                                                                        // lib.rs:501
    {
        UnionFind uf;
        uf.parent.resize(width_ * height_);
        uf.cluster_sizes.resize(width_ * height_);
        const int rc = cb_cat_connected_components(ctx_, buf.data(), (int)width_, (int)height_, uf.parent.data(), uf.cluster_sizes.data());
        if (rc != CB_OK) throw Error(rc, cb_last_error(ctx_));
        return uf;
    }
    const std::vector<int32_t> &points() const { return points_; }      // (x, y) pairs, the reference's scan order
    const std::vector<int32_t> &lines() const { return lines_; }        // (x1, y1, x2, y2)
    cb_timing timing() const { cb_timing t; cb_get_timing(ctx_, &t); return t; }
    std::vector<uint8_t> buf;                                           // the Color map (bufs.buf)

  private:
    void check(int rc) { if (rc != CB_OK) throw Error(rc, cb_last_error(ctx_)); }
    cb_ctx *ctx_ = nullptr;
    cb_ctx *det_ctx_ = nullptr;                 // detect_tags(): a second context sized for the undecimated decode stages
    int det_cap_ = 0;
    size_t width_, height_, cap_ = (size_t)1 << 20;
    std::vector<size_t> valid_tags_;
    std::vector<int32_t> points_, lines_;
};

}  // namespace cat

// ---- several GPUs, one process (cb_pool_*): one context + host thread per GPU, lists into slices of one array ----
class DetectorPool {
  public:
    DetectorPool(const std::vector<int> &devices, int max_w, int max_h, int max_batch, int max_dets, int bits_corrected = 3) : max_dets_(max_dets)
    {
        pool_ = cb_pool_create(devices.empty() ? nullptr : devices.data(), (int)devices.size(), max_w, max_h, max_batch, max_dets);
        if (!pool_) throw Error(CB_ERR_CUDA, cb_pool_last_error(nullptr));
        check(cb_pool_set_family_tag36h11(pool_, bits_corrected));
    }
    DetectorPool(const DetectorPool &) = delete;
    ~DetectorPool() { cb_pool_destroy(pool_); }
    int size() const { return cb_pool_size(pool_); }
    // out[f * max_dets + k], counts[f] for frame f of `frames`
    void detect(const uint8_t *frames, int w, int h, int stride, size_t frame_stride, int n_frames, cb_detection *out, int32_t *counts)
    {
        check(cb_pool_detect_gray(pool_, frames, w, h, stride, frame_stride, n_frames, out, counts));
    }
    cb_pool_timing timing() const { cb_pool_timing t; cb_pool_get_timing(pool_, &t); return t; }

  private:
    void check(int rc) { if (rc != CB_OK) throw Error(rc, cb_pool_last_error(pool_)); }
    cb_pool *pool_ = nullptr;
    int max_dets_;
};

}  // namespace chalkydri
