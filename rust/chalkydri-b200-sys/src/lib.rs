//! FFI crate for `libchalkydri_b200.so` (`include/chalkydri_b200.h`) with wrappers shaped like the types the
//! reference task uses: `apriltag::{DetectorBuilder, Detector, Detection}` (crates/apriltags/src/lib.rs:19,258-261,301-314)
//! and `chalkydri_sqpnp::SqPnP` (crates/chalkydri_sqpnp/src/lib.rs:183-304).
//!
//! SOURCE ONLY: written against the C header, never compiled in the build image (no rustc there).
#![allow(non_camel_case_types)]

use nalgebra::{Isometry3, Matrix3, Rotation3, Vector3};
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct cb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cb_detection {
    pub frame: i32,
    pub id: i32,
    pub hamming: i32,
    pub decision_margin: f32,
    pub h: [f64; 9],
    pub c: [f64; 2],
    pub p: [[f64; 2]; 4],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb_iso3 {
    pub t: [f64; 3],
    pub q: [f64; 4], // w, x, y, z
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cb_pose {
    pub rot: [f64; 9], // column-major
    pub pos: [f64; 3],
    pub std_devs: [f64; 3],
}

unsafe extern "C" {
    pub fn cb_create(device: c_int, max_width: c_int, max_height: c_int, max_batch: c_int, max_dets: c_int) -> *mut cb_ctx;
    pub fn cb_destroy(ctx: *mut cb_ctx);
    pub fn cb_last_error(ctx: *const cb_ctx) -> *const c_char;
    pub fn cb_set_family_tag36h11(ctx: *mut cb_ctx, bits_corrected: c_int) -> c_int;
    pub fn cb_detect_gray(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                          batch: c_int, out: *mut cb_detection, out_counts: *mut i32) -> c_int;
    pub fn cb_detect_gray_submit(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                                 batch: c_int) -> c_int;
    pub fn cb_detect_gray_collect(ctx: *mut cb_ctx, out: *mut cb_detection, out_counts: *mut i32) -> c_int;
    pub fn cb_detect_gray_pending(ctx: *const cb_ctx) -> c_int;
    pub fn cb_detect_pose_gray_submit(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                                      batch: c_int, gyro: *const f64, sign_change_error: f64) -> c_int;
    pub fn cb_detect_pose_gray_collect(ctx: *mut cb_ctx, out: *mut cb_detection, out_counts: *mut i32, poses: *mut cb_pose,
                                       pose_ok: *mut u8, pose_tags: *mut i32) -> c_int;
    pub fn cb_sqpnp_set(ctx: *mut cb_ctx, max_iter: c_int, tolerance: f64) -> c_int;
    pub fn cb_sqpnp_batch(ctx: *mut cb_ctx, tags: *const cb_iso3, bearings: *const f64, n_tags: *const i32, max_tags: c_int,
                          robot_to_cam: *const cb_iso3, gyro: *const f64, sign_change_error: f64, n: i64, out: *mut cb_pose,
                          ok: *mut u8) -> c_int;
    pub fn cb_create_solver_camera_transform(fwd: f64, left: f64, up: f64, roll_deg: f64, pitch_deg: f64, yaw_deg: f64,
                                             out: *mut cb_iso3) -> c_int;
    pub fn cb_unproject_opencv5(ctx: *mut cb_ctx, params9: *const f64, px: *const f64, n: i64, bearings: *mut f64, ok: *mut u8) -> c_int;
    pub fn cb_pack_vision_measurements(poses: *const cb_pose, pose_ok: *const u8, det_counts: *const i32, ts_us: *const u64, camera_id: u8,
                                       n: c_int, out: *mut cb_vision_measurement) -> c_int;
    // AprilTags::process on the device: field layout + camera once, then frames in -> detections and poses out
    pub fn cb_set_field(ctx: *mut cb_ctx, ids: *const i32, poses: *const cb_iso3, n: c_int) -> c_int;
    pub fn cb_set_camera(ctx: *mut cb_ctx, params9: *const f64, robot_to_cam: *const cb_iso3) -> c_int;
    pub fn cb_detect_pose_gray(ctx: *mut cb_ctx, frames: *const u8, width: c_int, height: c_int, stride: c_int, frame_stride: usize,
                               batch: c_int, gyro: *const f64, sign_change_error: f64, out: *mut cb_detection, out_counts: *mut i32,
                               poses: *mut cb_pose, pose_ok: *mut u8, pose_tags: *mut i32) -> c_int;
}

/// whacknet's 64-byte wire record (crates/whacknet/src/lib.rs:40-66)
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct cb_vision_measurement {
    pub x: f64, pub y: f64, pub rot: f64,
    pub std_x: f64, pub std_y: f64, pub std_rot: f64,
    pub ts: u64,
    pub camera_id: u8,
    pub tag_count: u8,
    pub reserved: [u8; 6],
}

#[derive(Debug)]
pub struct Error(pub i32, pub String);

fn last_error(ctx: *const cb_ctx) -> String {
    unsafe { CStr::from_ptr(cb_last_error(ctx)).to_string_lossy().into_owned() }
}

/// `apriltag::Detection` look-alike.
#[derive(Clone, Copy, Debug)]
pub struct Detection(cb_detection);
impl Detection {
    pub fn id(&self) -> usize { self.0.id as usize }
    pub fn hamming(&self) -> usize { self.0.hamming as usize }
    pub fn decision_margin(&self) -> f32 { self.0.decision_margin }
    pub fn corners(&self) -> [[f64; 2]; 4] { self.0.p }
    pub fn center(&self) -> [f64; 2] { self.0.c }
    pub fn homography(&self) -> Matrix3<f64> { Matrix3::from_row_slice(&self.0.h) }
}

/// Borrowed gray frame, the `image_u8_t` view built by `image_from_cuimage` (crates/apriltags/src/lib.rs:197-213).
pub struct Image<'a> {
    pub buf: &'a [u8],
    pub width: i32,
    pub height: i32,
    pub stride: i32,
}

pub struct DetectorBuilder {
    bits: Option<usize>,
    device: i32,
    max_w: i32,
    max_h: i32,
    max_dets: i32,
}
impl Default for DetectorBuilder {
    fn default() -> Self { Self { bits: None, device: 0, max_w: 1600, max_h: 1304, max_dets: 64 } }
}
impl DetectorBuilder {
    /// `Family` is tag36h11 (the reference's FAMILY, lib.rs:45); other families are rejected by the caller's parse.
    pub fn add_family_bits(mut self, _family_tag36h11: (), bits_corrected: usize) -> Self { self.bits = Some(bits_corrected); self }
    pub fn capacity(mut self, max_w: i32, max_h: i32, max_dets: i32) -> Self { self.max_w = max_w; self.max_h = max_h; self.max_dets = max_dets; self }
    pub fn device(mut self, device: i32) -> Self { self.device = device; self }
    pub fn build(self) -> Result<Detector, Error> {
        let bits = self.bits.ok_or_else(|| Error(-5, "no tag family added".into()))?;
        let ctx = unsafe { cb_create(self.device, self.max_w, self.max_h, 1, self.max_dets) };
        if ctx.is_null() { return Err(Error(-2, last_error(std::ptr::null()))); }
        let rc = unsafe { cb_set_family_tag36h11(ctx, bits as c_int) };
        if rc != 0 { let e = Error(rc, last_error(ctx)); unsafe { cb_destroy(ctx) }; return Err(e); }
        Ok(Detector { ctx, max_dets: self.max_dets as usize, out: Vec::new() })
    }
}

pub struct Detector { ctx: *mut cb_ctx, max_dets: usize, out: Vec<cb_detection> }
unsafe impl Send for Detector {}
impl Detector {
    /// `Detector::detect(&mut self, &Image) -> Vec<Detection>` (lib.rs:301)
    pub fn detect(&mut self, image: &Image) -> Vec<Detection> {
        self.out.resize(self.max_dets, unsafe { std::mem::zeroed() });
        let mut count: i32 = 0;
        let rc = unsafe {
            cb_detect_gray(self.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                           (image.stride as usize) * (image.height as usize), 1, self.out.as_mut_ptr(), &mut count)
        };
        if rc != 0 { panic!("chalkydri_b200: {}", last_error(self.ctx)); } // the reference unwraps as well
        self.out[..count as usize].iter().map(|d| Detection(*d)).collect()
    }
    /// Streaming form: enqueue a batch of frames (`frame_stride` bytes apart) and return at once; at most two batches in
    /// flight.  The borrow keeps the frames alive: the caller holds `frames` until the matching `collect`.
    pub fn submit(&mut self, frames: &[u8], width: i32, height: i32, stride: i32, frame_stride: usize, batch: i32) -> Result<(), String> {
        assert!(frames.len() >= frame_stride * (batch as usize - 1) + (stride as usize) * (height as usize));
        let rc = unsafe { cb_detect_gray_submit(self.ctx, frames.as_ptr(), width, height, stride, frame_stride, batch) };
        if rc != 0 { Err(last_error(self.ctx)) } else { Ok(()) }
    }
    /// Wait for the oldest submitted batch: `out[b * max_dets + k]`, `counts[b]`.
    pub fn collect(&mut self, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), String> {
        assert!(out.len() >= counts.len() * self.max_dets);
        let rc = unsafe { cb_detect_gray_collect(self.ctx, out.as_mut_ptr(), counts.as_mut_ptr()) };
        if rc != 0 { Err(last_error(self.ctx)) } else { Ok(()) }
    }
    pub fn raw(&self) -> *mut cb_ctx { self.ctx }
}
impl Drop for Detector { fn drop(&mut self) { unsafe { cb_destroy(self.ctx) } } }

fn to_iso(i: &Isometry3<f64>) -> cb_iso3 {
    let q = i.rotation.quaternion();
    cb_iso3 { t: [i.translation.x, i.translation.y, i.translation.z], q: [q.w, q.i, q.j, q.k] }
}

/// `chalkydri_sqpnp::SqPnP` look-alike (lib.rs:183-304).
pub struct SqPnP { ctx: *mut cb_ctx, owns: bool, max_iter: usize, tol: f64 }
impl SqPnP {
    pub fn new() -> Self {
        let ctx = unsafe { cb_create(0, 8, 8, 1, 1) };
        assert!(!ctx.is_null(), "chalkydri_b200: {}", last_error(std::ptr::null()));
        Self { ctx, owns: true, max_iter: 15, tol: 1e-8 }
    }
    pub fn max_iter(mut self, max_iter: usize) -> Self { self.max_iter = max_iter; unsafe { cb_sqpnp_set(self.ctx, max_iter as c_int, self.tol) }; self }
    pub fn tolerance(mut self, tol: f64) -> Self { self.tol = tol; unsafe { cb_sqpnp_set(self.ctx, self.max_iter as c_int, tol) }; self }

    pub fn solve_robot_pose(&mut self, points_isometry: &[Isometry3<f64>], points_2d: &[Vector3<f64>], robot_to_cam: &Isometry3<f64>,
                            gyro: f64, sign_change_error: f64) -> Option<(Rotation3<f64>, Vector3<f64>, Vector3<f64>)> {
        let n = points_isometry.len();
        if n * 4 < 3 || n * 4 != points_2d.len() || n > 16 { return None; }
        let tags: Vec<cb_iso3> = points_isometry.iter().map(to_iso).collect();
        let bearings: Vec<f64> = points_2d.iter().flat_map(|v| [v.x, v.y, v.z]).collect();
        let r2c = to_iso(robot_to_cam);
        let (nt, g) = ([n as i32], [gyro]);
        let mut out: cb_pose = unsafe { std::mem::zeroed() };
        let mut ok: u8 = 0;
        let rc = unsafe { cb_sqpnp_batch(self.ctx, tags.as_ptr(), bearings.as_ptr(), nt.as_ptr(), n as c_int, &r2c, g.as_ptr(), sign_change_error, 1, &mut out, &mut ok) };
        if rc != 0 || ok == 0 { return None; }
        Some((Rotation3::from_matrix_unchecked(Matrix3::from_column_slice(&out.rot)), Vector3::from(out.pos), Vector3::from(out.std_devs)))
    }

    pub fn create_solver_camera_transform(fwd_m: f64, left_m: f64, up_m: f64, roll_deg: f64, pitch_deg: f64, yaw_deg: f64) -> Isometry3<f64> {
        let mut o = cb_iso3::default();
        unsafe { cb_create_solver_camera_transform(fwd_m, left_m, up_m, roll_deg, pitch_deg, yaw_deg, &mut o) };
        Isometry3::from_parts(nalgebra::Translation3::new(o.t[0], o.t[1], o.t[2]),
                              nalgebra::UnitQuaternion::new_unchecked(nalgebra::Quaternion::new(o.q[0], o.q[1], o.q[2], o.q[3])))
    }
}
impl Drop for SqPnP { fn drop(&mut self) { if self.owns { unsafe { cb_destroy(self.ctx) } } } }
