//! Raw bindings: one declaration per symbol of `include/chalkydri_b200.h`, in the header's order.
//! `tests/test_abi.py::test_rust_extern_block_lists_the_header` checks the two lists against each other (no rustc needed).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const CB_OK: c_int = 0;
pub const CB_ERR_ARG: c_int = -1;
pub const CB_ERR_CUDA: c_int = -2;
pub const CB_ERR_UNSUPPORTED: c_int = -3;
pub const CB_ERR_OVERFLOW: c_int = -4;
pub const CB_ERR_STATE: c_int = -5;

#[repr(C)]
pub struct cb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct cb_pool {
    _private: [u8; 0],
}

/// `apriltag_detection_t` as `apriltag::Detection` exposes it, plus the frame index inside the batch.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cb_detection {
    pub frame: i32,
    pub id: i32,
    pub hamming: i32,
    pub decision_margin: f32,
    pub h: [f64; 9], // row-major 3x3
    pub c: [f64; 2],
    pub p: [[f64; 2]; 4],
}

/// nalgebra `Isometry3<f64>`: translation, unit quaternion (w, x, y, z).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb_iso3 {
    pub t: [f64; 3],
    pub q: [f64; 4],
}

/// `Some((Rot3, Vec3 position, Vec3 std_devs))` of `SqPnP::solve_robot_pose`.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cb_pose {
    pub rot: [f64; 9], // column-major
    pub pos: [f64; 3],
    pub std_devs: [f64; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb_timing {
    pub h2d_ms: f32,
    pub preprocess_ms: f32,
    pub threshold_ms: f32,
    pub ccl_ms: f32,
    pub cluster_ms: f32,
    pub quad_ms: f32,
    pub decode_ms: f32,
    pub d2h_ms: f32,
    pub total_ms: f32,
    pub threshold_launches: i32,
    pub kernel_launches: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb_pool_timing {
    pub wall_ms: f32,
    pub max_device_ms: f32,
    pub min_device_ms: f32,
    pub n_devices: i32,
}

/// whacknet's 64-byte wire record (crates/whacknet/src/lib.rs:40-66).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb_vision_measurement {
    pub x: f64,
    pub y: f64,
    pub rot: f64,
    pub std_x: f64,
    pub std_y: f64,
    pub std_rot: f64,
    pub ts: u64,
    pub camera_id: u8,
    pub tag_count: u8,
    pub reserved: [u8; 6],
}
const _: () = assert!(std::mem::size_of::<cb_vision_measurement>() == 64); // the reference's one test (whacknet/src/lib.rs:92-95)
const _: () = assert!(std::mem::size_of::<cb_detection>() == 168);

unsafe extern "C" {
    // ---- lifetime ----
    pub fn cb_create(device: c_int, max_width: c_int, max_height: c_int, max_batch: c_int, max_dets_per_frame: c_int) -> *mut cb_ctx;
    pub fn cb_destroy(ctx: *mut cb_ctx);
    pub fn cb_last_error(ctx: *const cb_ctx) -> *const c_char;
    pub fn cb_set_family_tag36h11(ctx: *mut cb_ctx, bits_corrected: c_int) -> c_int;
    pub fn cb_set_params(ctx: *mut cb_ctx, quad_decimate: f32, quad_sigma: f32, refine_edges: c_int, decode_sharpening: f64,
                         min_cluster_pixels: c_int, max_nmaxima: c_int, critical_rad: f32, max_line_fit_mse: f32,
                         min_white_black_diff: c_int) -> c_int;
    // ---- Detector::detect ----
    pub fn cb_detect_gray(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                          batch: c_int, out: *mut cb_detection, out_counts: *mut i32) -> c_int;
    pub fn cb_detect_gray_device(ctx: *mut cb_ctx, frames_dev: *const u8, width: c_int, height: c_int, stride: c_int,
                                 frame_stride: usize, batch: c_int, out: *mut cb_detection, out_counts: *mut i32) -> c_int;
    pub fn cb_detect_gray_submit(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                                 batch: c_int) -> c_int;
    pub fn cb_detect_gray_collect(ctx: *mut cb_ctx, out: *mut cb_detection, out_counts: *mut i32) -> c_int;
    pub fn cb_detect_gray_pending(ctx: *const cb_ctx) -> c_int;
    pub fn cb_detect_rgb(ctx: *mut cb_ctx, frames_rgb: *const u8, width: c_int, height: c_int, batch: c_int, out: *mut cb_detection,
                         out_counts: *mut i32) -> c_int;
    pub fn cb_detect_yuyv(ctx: *mut cb_ctx, frames_yuyv: *const u8, width: c_int, height: c_int, batch: c_int, out: *mut cb_detection,
                          out_counts: *mut i32) -> c_int;
    pub fn cb_detect_yuv420(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, batch: c_int, out: *mut cb_detection,
                            out_counts: *mut i32) -> c_int;
    // ---- stage taps ----
    pub fn cb_rgb_to_gray(ctx: *mut cb_ctx, frames_rgb: *const u8, width: c_int, height: c_int, batch: c_int, gray_out: *mut u8) -> c_int;
    pub fn cb_yuyv_to_gray(ctx: *mut cb_ctx, frames_yuyv: *const u8, width: c_int, height: c_int, batch: c_int, gray_out: *mut u8) -> c_int;
    pub fn cb_decimated_size(ctx: *const cb_ctx, width: c_int, height: c_int, w: *mut c_int, h: *mut c_int) -> c_int;
    pub fn cb_threshold(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize, batch: c_int,
                        out: *mut u8) -> c_int;
    pub fn cb_labels(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize, batch: c_int,
                     labels: *mut u32, sizes: *mut u32) -> c_int;
    pub fn cb_clusters(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize, batch: c_int,
                       pts: *mut i16, cluster_of: *mut i32, cap: i64, npoints: *mut i64, nclusters: *mut i32) -> c_int;
    pub fn cb_quads(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize, batch: c_int,
                    quads: *mut f32, cap: c_int, counts: *mut i32, npoints_total: *mut i64) -> c_int;
    pub fn cb_frame_flags(ctx: *const cb_ctx, flags: *mut u32, n: c_int) -> c_int;
    pub fn cb_get_timing(ctx: *const cb_ctx, t: *mut cb_timing) -> c_int;
    // ---- solver ----
    pub fn cb_sqpnp_set(ctx: *mut cb_ctx, max_iter: c_int, tolerance: f64) -> c_int;
    pub fn cb_sqpnp_batch(ctx: *mut cb_ctx, tags: *const cb_iso3, bearings: *const f64, n_tags: *const i32, max_tags: c_int,
                          robot_to_cam: *const cb_iso3, gyro: *const f64, sign_change_error: f64, n: i64, out: *mut cb_pose,
                          ok: *mut u8) -> c_int;
    pub fn cb_sqpnp_batch_device(ctx: *mut cb_ctx, tags: *const cb_iso3, bearings: *const f64, n_tags: *const i32, max_tags: c_int,
                                 robot_to_cam: *const cb_iso3, gyro: *const f64, sign_change_error: f64, n: i64, out: *mut cb_pose,
                                 ok: *mut u8) -> c_int;
    // ---- AprilTags::process on the device ----
    pub fn cb_set_field(ctx: *mut cb_ctx, ids: *const i32, poses: *const cb_iso3, n: c_int) -> c_int;
    pub fn cb_set_camera(ctx: *mut cb_ctx, params9: *const f64, robot_to_cam: *const cb_iso3) -> c_int;
    pub fn cb_detect_pose_gray(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                               batch: c_int, gyro: *const f64, sign_change_error: f64, out: *mut cb_detection, out_counts: *mut i32,
                               poses: *mut cb_pose, pose_ok: *mut u8, pose_tags: *mut i32) -> c_int;
    pub fn cb_detect_pose_gray_submit(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                                      batch: c_int, gyro: *const f64, sign_change_error: f64) -> c_int;
    pub fn cb_detect_pose_gray_collect(ctx: *mut cb_ctx, out: *mut cb_detection, out_counts: *mut i32, poses: *mut cb_pose,
                                       pose_ok: *mut u8, pose_tags: *mut i32) -> c_int;
    pub fn cb_pack_vision_measurements(poses: *const cb_pose, pose_ok: *const u8, det_counts: *const i32, ts_us: *const u64, camera_id: u8,
                                       n: c_int, out: *mut cb_vision_measurement) -> c_int;
    pub fn cb_create_solver_camera_transform(fwd_m: f64, left_m: f64, up_m: f64, roll_deg: f64, pitch_deg: f64, yaw_deg: f64,
                                             out: *mut cb_iso3) -> c_int;
    pub fn cb_unproject_opencv5(ctx: *mut cb_ctx, params9: *const f64, px: *const f64, n: i64, bearings: *mut f64, ok: *mut u8) -> c_int;
    // ---- CAT stages ----
    pub fn cb_cat_calc_otsu(ctx: *mut cb_ctx, rgb: *const u8, width: c_int, height: c_int, color: *mut u8) -> c_int;
    pub fn cb_cat_thresh(ctx: *mut cb_ctx, rgb: *const u8, width: c_int, height: c_int, color: *mut u8) -> c_int;
    pub fn cb_cat_detect_corners(ctx: *mut cb_ctx, color: *const u8, width: c_int, height: c_int, xy: *mut i32, cap: i64, n: *mut i64) -> c_int;
    pub fn cb_cat_check_edges(ctx: *mut cb_ctx, color: *const u8, width: c_int, height: c_int, xy: *const i32, npts: i64, lines: *mut i32,
                              cap: i64, n: *mut i64) -> c_int;
    pub fn cb_cat_process_frame(ctx: *mut cb_ctx, rgb: *const u8, width: c_int, height: c_int, color: *mut u8, xy: *mut i32, xy_cap: i64,
                                n_points: *mut i64, lines: *mut i32, lines_cap: i64, n_lines: *mut i64) -> c_int;
    pub fn cb_cat_detect_tags(ctx: *mut cb_ctx, rgb: *const u8, width: c_int, height: c_int, use_otsu: c_int, out: *mut cb_detection,
                              out_count: *mut i32) -> c_int;
    pub fn cb_cat_connected_components(ctx: *mut cb_ctx, color: *const u8, width: c_int, height: c_int, labels: *mut u32,
                                       sizes: *mut u32) -> c_int;
    // ---- several GPUs, one process ----
    pub fn cb_pool_create(devices: *const c_int, n_devices: c_int, max_width: c_int, max_height: c_int, max_batch: c_int,
                          max_dets_per_frame: c_int) -> *mut cb_pool;
    pub fn cb_pool_destroy(pool: *mut cb_pool);
    pub fn cb_pool_last_error(pool: *const cb_pool) -> *const c_char;
    pub fn cb_pool_size(pool: *const cb_pool) -> c_int;
    pub fn cb_pool_context(pool: *mut cb_pool, i: c_int) -> *mut cb_ctx;
    pub fn cb_pool_set_family_tag36h11(pool: *mut cb_pool, bits_corrected: c_int) -> c_int;
    pub fn cb_pool_detect_gray(pool: *mut cb_pool, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                               n_frames: c_int, out: *mut cb_detection, out_counts: *mut i32) -> c_int;
    pub fn cb_pool_get_timing(pool: *const cb_pool, t: *mut cb_pool_timing) -> c_int;
    // ---- plumbing ----
    pub fn cb_host_alloc(bytes: usize) -> *mut c_void;
    pub fn cb_host_free(p: *mut c_void);
    pub fn cb_device_alloc(ctx: *mut cb_ctx, bytes: usize) -> *mut c_void;
    pub fn cb_device_free(ctx: *mut cb_ctx, p: *mut c_void);
    pub fn cb_memcpy_h2d(ctx: *mut cb_ctx, dst_dev: *mut c_void, src_host: *const c_void, bytes: usize) -> c_int;
    pub fn cb_memcpy_d2h(ctx: *mut cb_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: usize) -> c_int;
    pub fn cb_device_count() -> c_int;
    pub fn cb_version() -> *const c_char;
}
