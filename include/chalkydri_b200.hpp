// chalkydri_b200.hpp -- header-only C++ host mirror of the reference interfaces, above the C ABI (chalkydri_b200.h).
//
// The reference's host code is Rust; where its toolchain is absent the host side above the C ABI is C++ (this file) and
// Python (chalkydri_b200/*.py).  Names, argument meaning and failure behaviour follow
//   apriltag::{DetectorBuilder, Detector, Detection}   /root/reference/crates/apriltags/src/lib.rs:19,258-261,301-314
//   chalkydri_sqpnp::SqPnP                              /root/reference/crates/chalkydri_sqpnp/src/lib.rs:183-304,430-461
// Config errors throw (the reference unwrap()s / panics), the solver returns std::optional (the reference: Option).
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
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
    cb_timing timing() const { cb_timing t; cb_get_timing(ctx_, &t); return t; }
    cb_ctx *ctx() { return ctx_; }

  private:
    void check(int rc) { if (rc != CB_OK) throw Error(rc, cb_last_error(ctx_)); }
    cb_ctx *ctx_ = nullptr;
    int max_dets_, max_batch_;
};

class DetectorBuilder {
  public:
    static DetectorBuilder default_() { return DetectorBuilder(); }
    DetectorBuilder &add_family_bits(const std::string &family, size_t bits_corrected)
    {
        if (family != "tag36h11") throw Error(CB_ERR_UNSUPPORTED, "unknown family " + family + " (this build carries tag36h11, the reference's FAMILY)");
        bits_ = (int)bits_corrected;
        return *this;
    }
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
    SqPnP(const SqPnP &) = delete;
    ~SqPnP() { cb_destroy(ctx_); }
    SqPnP &max_iter(size_t n) { max_iter_ = (int)n; cb_sqpnp_set(ctx_, max_iter_, tol_); return *this; }      // lib.rs:214-217
    SqPnP &tolerance(double t) { tol_ = t; cb_sqpnp_set(ctx_, max_iter_, tol_); return *this; }               // lib.rs:219-222

    // solve_robot_pose(points_isometry, points_2d, robot_to_cam, gyro, sign_change_error) -> Option<(Rot3, Vec3, Vec3)>  (lib.rs:297-304)
    std::optional<RobotPoseResult> solve_robot_pose(const std::vector<cb_iso3> &points_isometry, const std::vector<std::array<double, 3>> &points_2d,
                                                    const cb_iso3 &robot_to_cam, double gyro, double sign_change_error)
    {
        const size_t n = points_isometry.size();
        if (n * 4 < 3 || n * 4 != points_2d.size() || n > 16) return std::nullopt;      // lib.rs:255-257
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

}  // namespace chalkydri
