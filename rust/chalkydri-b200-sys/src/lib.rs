//! FFI crate for `libchalkydri_b200.so` (`include/chalkydri_b200.h`): raw bindings in [`ffi`], and on top of them wrappers shaped
//! like the types the reference task uses, so that `crates/apriltags/src/lib.rs` changes its `use` lines and little else:
//!
//! * `apriltag::{Family, DetectorBuilder, Detector, Detection, Image}` -- crates/apriltags/src/lib.rs:19,229,258-261,279-282,301-314
//! * `chalkydri_sqpnp::SqPnP` (`Clone + Debug + Default`, const builders) -- crates/chalkydri_sqpnp/src/lib.rs:182-222,297-304,430-461
//! * `chalkydri_apriltags::Detector` (CAT) -- crates/chalkydri-apriltags/src/lib.rs:158,265,501
//! * [`DetectorPool`]: one process, one context per GPU, lists into slices of one array (no collective)
//!
//! SOURCE ONLY: written against the C header; the build image has no rustc, so this crate has not been compiled there.
//! The C++ (`include/chalkydri_b200.hpp`) and Python (`chalkydri_b200/*.py`) mirrors of the same wrappers are what the tests run.

pub mod ffi;

use ffi::*;
use nalgebra::{Isometry3, Matrix3, Quaternion, Rotation3, Translation3, UnitQuaternion, Vector3};
use std::ffi::CStr;
use std::os::raw::c_int;
use std::str::FromStr;

#[derive(Debug, Clone)]
pub struct Error(pub i32, pub String);
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "chalkydri_b200 error {}: {}", self.0, self.1) }
}
impl std::error::Error for Error {}

fn last_error(ctx: *const cb_ctx) -> String {
    unsafe { CStr::from_ptr(cb_last_error(ctx)).to_string_lossy().into_owned() }
}
fn check(ctx: *const cb_ctx, rc: c_int) -> Result<(), Error> {
    if rc == CB_OK { Ok(()) } else { Err(Error(rc, last_error(ctx))) }
}

/// `apriltag::Family`.  The reference parses its `family` config string and unwraps (crates/apriltags/src/lib.rs:229);
/// this build carries tag36h11, the reference's `FAMILY` (lib.rs:45).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Family {
    Tag36h11,
}
impl FromStr for Family {
    type Err = Error;
    fn from_str(s: &str) -> Result<Self, Error> {
        match s {
            "tag36h11" => Ok(Family::Tag36h11),
            other => Err(Error(CB_ERR_UNSUPPORTED, format!("unknown family {other:?}: this build carries tag36h11"))),
        }
    }
}

/// `apriltag::Detection`.
#[derive(Clone, Copy, Debug)]
pub struct Detection(cb_detection);
impl Detection {
    pub fn id(&self) -> usize { self.0.id as usize }
    pub fn hamming(&self) -> usize { self.0.hamming as usize }
    pub fn decision_margin(&self) -> f32 { self.0.decision_margin }
    pub fn corners(&self) -> [[f64; 2]; 4] { self.0.p }
    pub fn center(&self) -> [f64; 2] { self.0.c }
    pub fn homography(&self) -> Matrix3<f64> { Matrix3::from_row_slice(&self.0.h) }
    pub fn raw(&self) -> &cb_detection { &self.0 }
}

/// Borrowed gray frame: the `image_u8_t` view `image_from_cuimage` builds (crates/apriltags/src/lib.rs:197-213), without the
/// boxed header that function has to leak.
pub struct Image<'a> {
    pub buf: &'a [u8],
    pub width: i32,
    pub height: i32,
    pub stride: i32,
}

/// `apriltag::DetectorBuilder`; `capacity` / `device` / `max_batch` are the additions a device context needs (buffers are sized once).
pub struct DetectorBuilder {
    bits: Option<usize>,
    device: i32,
    max_w: i32,
    max_h: i32,
    max_batch: i32,
    max_dets: i32,
}
impl Default for DetectorBuilder {
    fn default() -> Self { Self { bits: None, device: 0, max_w: 1600, max_h: 1304, max_batch: 1, max_dets: 64 } }
}
impl DetectorBuilder {
    /// `DetectorBuilder::add_family_bits(family, bits_corrected)` (lib.rs:259, 280).
    pub fn add_family_bits(mut self, family: Family, bits_corrected: usize) -> Self {
        let Family::Tag36h11 = family;
        self.bits = Some(bits_corrected);
        self
    }
    pub fn capacity(mut self, max_w: i32, max_h: i32, max_dets: i32) -> Self { self.max_w = max_w; self.max_h = max_h; self.max_dets = max_dets; self }
    pub fn max_batch(mut self, max_batch: i32) -> Self { self.max_batch = max_batch; self }
    pub fn device(mut self, device: i32) -> Self { self.device = device; self }
    pub fn build(self) -> Result<Detector, Error> {
        let bits = self.bits.ok_or_else(|| Error(CB_ERR_STATE, "no tag family added".into()))?;
        let ctx = unsafe { cb_create(self.device, self.max_w, self.max_h, self.max_batch, self.max_dets) };
        if ctx.is_null() { return Err(Error(CB_ERR_CUDA, last_error(std::ptr::null()))); }
        let rc = unsafe { cb_set_family_tag36h11(ctx, bits as c_int) };
        if rc != CB_OK { let e = Error(rc, last_error(ctx)); unsafe { cb_destroy(ctx) }; return Err(e); }
        Ok(Detector { ctx, max_dets: self.max_dets as usize, out: Vec::new() })
    }
}

/// `apriltag::Detector`.
pub struct Detector { ctx: *mut cb_ctx, max_dets: usize, out: Vec<cb_detection> }
unsafe impl Send for Detector {}
impl Detector {
    /// `Detector::detect(&mut self, &Image) -> Vec<Detection>` (lib.rs:301).  Panics on a library error, like the reference's unwraps.
    pub fn detect(&mut self, image: &Image) -> Vec<Detection> {
        self.out.resize(self.max_dets, unsafe { std::mem::zeroed() });
        let mut count: i32 = 0;
        let rc = unsafe {
            cb_detect_gray(self.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                           (image.stride as usize) * (image.height as usize), 1, self.out.as_mut_ptr(), &mut count)
        };
        // a frame so cluttered that a fixed-size device table overflowed is "nothing detected" (upstream has no such limit)
        if rc == CB_ERR_OVERFLOW { return Vec::new(); }
        if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
        self.out[..count as usize].iter().map(|d| Detection(*d)).collect()
    }
    /// Overflow bits per frame of the last completed call (a flagged frame reported an empty list); returns how many are flagged.
    pub fn frame_flags(&self, flags: &mut [u32]) -> usize {
        unsafe { cb_frame_flags(self.ctx, flags.as_mut_ptr(), flags.len() as c_int).max(0) as usize }
    }
    /// Batched form: `batch` frames `frame_stride` bytes apart; `out[b * max_dets + k]`, `counts[b]`.
    pub fn detect_batch(&mut self, frames: &[u8], width: i32, height: i32, stride: i32, frame_stride: usize, out: &mut [cb_detection],
                        counts: &mut [i32]) -> Result<(), Error> {
        assert!(out.len() >= counts.len() * self.max_dets);
        check(self.ctx, unsafe { cb_detect_gray(self.ctx, frames.as_ptr(), width, height, stride, frame_stride, counts.len() as c_int,
                                                out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
    /// NV12 / I420 buffers: the Y plane of every frame is the gray image (gst_to_cu.rs:152-188).
    pub fn detect_yuv420(&mut self, frames: &[u8], width: i32, height: i32, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), Error> {
        check(self.ctx, unsafe { cb_detect_yuv420(self.ctx, frames.as_ptr(), width, height, counts.len() as c_int, out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
    pub fn detect_yuyv(&mut self, frames: &[u8], width: i32, height: i32, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), Error> {
        check(self.ctx, unsafe { cb_detect_yuyv(self.ctx, frames.as_ptr(), width, height, counts.len() as c_int, out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
    pub fn detect_rgb(&mut self, frames: &[u8], width: i32, height: i32, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), Error> {
        check(self.ctx, unsafe { cb_detect_rgb(self.ctx, frames.as_ptr(), width, height, counts.len() as c_int, out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
    /// Streaming form: enqueue a batch and return at once; at most two batches in flight.  The caller keeps `frames` alive and
    /// unchanged until the matching `collect` (exactly where the reference drops its `CuImage` handle today).
    pub fn submit(&mut self, frames: &[u8], width: i32, height: i32, stride: i32, frame_stride: usize, batch: i32) -> Result<(), Error> {
        assert!(frames.len() >= frame_stride * (batch as usize - 1) + (stride as usize) * (height as usize));
        check(self.ctx, unsafe { cb_detect_gray_submit(self.ctx, frames.as_ptr(), width, height, stride, frame_stride, batch) })
    }
    /// Wait for the oldest submitted batch.
    pub fn collect(&mut self, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), Error> {
        assert!(out.len() >= counts.len() * self.max_dets);
        check(self.ctx, unsafe { cb_detect_gray_collect(self.ctx, out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
    pub fn pending(&self) -> i32 { unsafe { cb_detect_gray_pending(self.ctx) } }
    /// `apriltag_detector_t` fields the reference leaves at their defaults.
    #[allow(clippy::too_many_arguments)]
    pub fn set_params(&mut self, quad_decimate: f32, quad_sigma: f32, refine_edges: bool, decode_sharpening: f64, min_cluster_pixels: i32,
                      max_nmaxima: i32, critical_rad: f32, max_line_fit_mse: f32, min_white_black_diff: i32) -> Result<(), Error> {
        check(self.ctx, unsafe { cb_set_params(self.ctx, quad_decimate, quad_sigma, refine_edges as c_int, decode_sharpening, min_cluster_pixels,
                                               max_nmaxima, critical_rad, max_line_fit_mse, min_white_black_diff) })
    }
    /// Stage taps (parity tests): threshold map, component labels + sizes, candidate quads.
    pub fn threshold(&mut self, image: &Image, out: &mut [u8]) -> Result<(), Error> {
        check(self.ctx, unsafe { cb_threshold(self.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                                              (image.stride as usize) * (image.height as usize), 1, out.as_mut_ptr()) })
    }
    pub fn labels(&mut self, image: &Image, labels: &mut [u32], sizes: &mut [u32]) -> Result<(), Error> {
        check(self.ctx, unsafe { cb_labels(self.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                                           (image.stride as usize) * (image.height as usize), 1, labels.as_mut_ptr(), sizes.as_mut_ptr()) })
    }
    pub fn quads(&mut self, image: &Image, quads: &mut [[f32; 8]]) -> Result<(usize, i64), Error> {
        let (mut count, mut npoints) = (0i32, 0i64);
        check(self.ctx, unsafe { cb_quads(self.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                                          (image.stride as usize) * (image.height as usize), 1, quads.as_mut_ptr() as *mut f32,
                                          quads.len() as c_int, &mut count, &mut npoints) })?;
        Ok((count as usize, npoints))
    }
    /// Stage tap of gradient_clusters(): `pts[k] = [x, y, gx, gy]` as upstream stores a point, `cluster_of[k]`; returns (points, clusters).
    pub fn clusters(&mut self, image: &Image, pts: &mut [[i16; 4]], cluster_of: &mut [i32]) -> Result<(i64, i32), Error> {
        let (mut n, mut ncl) = (0i64, 0i32);
        check(self.ctx, unsafe { cb_clusters(self.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                                             (image.stride as usize) * (image.height as usize), 1, pts.as_mut_ptr() as *mut i16,
                                             cluster_of.as_mut_ptr(), pts.len().min(cluster_of.len()) as i64, &mut n, &mut ncl) })?;
        Ok((n, ncl))
    }
    pub fn decimated_size(&self, width: i32, height: i32) -> (i32, i32) {
        let (mut w, mut h) = (0, 0);
        unsafe { cb_decimated_size(self.ctx, width, height, &mut w, &mut h) };
        (w, h)
    }
    pub fn timing(&self) -> cb_timing { let mut t = cb_timing::default(); unsafe { cb_get_timing(self.ctx, &mut t) }; t }
    pub fn raw(&self) -> *mut cb_ctx { self.ctx }
}
impl Drop for Detector { fn drop(&mut self) { unsafe { cb_destroy(self.ctx) } } }

fn to_iso(i: &Isometry3<f64>) -> cb_iso3 {
    let q = i.rotation.quaternion();
    cb_iso3 { t: [i.translation.x, i.translation.y, i.translation.z], q: [q.w, q.i, q.j, q.k] }
}
fn from_iso(o: &cb_iso3) -> Isometry3<f64> {
    Isometry3::from_parts(Translation3::new(o.t[0], o.t[1], o.t[2]), UnitQuaternion::new_unchecked(Quaternion::new(o.q[0], o.q[1], o.q[2], o.q[3])))
}

/// `chalkydri_sqpnp::SqPnP` (crates/chalkydri_sqpnp/src/lib.rs:182-222): `Clone + Debug + Default`, `new` / `max_iter` / `tolerance`
/// are `const fn` there, so the device context is created on first use.
#[derive(Debug)]
pub struct SqPnP { ctx: *mut cb_ctx, max_iter: usize, tol: f64 }
unsafe impl Send for SqPnP {}
impl Default for SqPnP { fn default() -> Self { Self::new() } }
impl Clone for SqPnP {
    /// A fresh solver with the same settings (the reference's clone copies scratch vectors that every solve clears first).
    fn clone(&self) -> Self { Self { ctx: std::ptr::null_mut(), max_iter: self.max_iter, tol: self.tol } }
}
impl SqPnP {
    pub const fn new() -> Self { Self { ctx: std::ptr::null_mut(), max_iter: 15, tol: 1e-8 } }      // DEFAULT_MAX_ITER, tol_sq 1e-16 (lib.rs:203-204)
    pub const fn max_iter(mut self, max_iter: usize) -> Self { self.max_iter = max_iter; self }
    pub const fn tolerance(mut self, tol: f64) -> Self { self.tol = tol; self }
    fn ctx(&mut self) -> *mut cb_ctx {
        if self.ctx.is_null() {
            self.ctx = unsafe { cb_create(0, 8, 8, 1, 1) };
            assert!(!self.ctx.is_null(), "chalkydri_b200: {}", last_error(std::ptr::null()));
            unsafe { cb_sqpnp_set(self.ctx, self.max_iter as c_int, self.tol) };
        }
        self.ctx
    }
    /// `solve_robot_pose(&mut self, points_isometry, points_2d, robot_to_cam, gyro, sign_change_error)` (lib.rs:297-304): `None` for
    /// fewer than 3 points, a length mismatch, or no candidate with every point in front of the camera.
    pub fn solve_robot_pose(&mut self, points_isometry: &[Isometry3<f64>], points_2d: &[Vector3<f64>], robot_to_cam: &Isometry3<f64>,
                            gyro: f64, sign_change_error: f64) -> Option<(Rotation3<f64>, Vector3<f64>, Vector3<f64>)> {
        let n = points_isometry.len();
        if n * 4 < 3 || n * 4 != points_2d.len() { return None; }
        let tags: Vec<cb_iso3> = points_isometry.iter().map(to_iso).collect();
        let bearings: Vec<f64> = points_2d.iter().flat_map(|v| [v.x, v.y, v.z]).collect();
        let r2c = to_iso(robot_to_cam);
        let (nt, g) = ([n as i32], [gyro]);
        let mut out: cb_pose = unsafe { std::mem::zeroed() };
        let mut ok: u8 = 0;
        let ctx = self.ctx();
        let rc = unsafe { cb_sqpnp_batch(ctx, tags.as_ptr(), bearings.as_ptr(), nt.as_ptr(), n as c_int, &r2c, g.as_ptr(), sign_change_error, 1, &mut out, &mut ok) };
        if rc != CB_OK || ok == 0 { return None; }
        Some((Rotation3::from_matrix_unchecked(Matrix3::from_column_slice(&out.rot)), Vector3::from(out.pos), Vector3::from(out.std_devs)))
    }
    /// Many independent problems in one launch: problem `i` uses `tags[i * max_tags ..][.. n_tags[i]]` and 4 bearings per tag.
    #[allow(clippy::too_many_arguments)]
    pub fn solve_robot_pose_batch(&mut self, tags: &[cb_iso3], bearings: &[f64], n_tags: &[i32], max_tags: usize, robot_to_cam: &Isometry3<f64>,
                                  gyro: &[f64], sign_change_error: f64, out: &mut [cb_pose], ok: &mut [u8]) -> Result<(), Error> {
        let n = n_tags.len();
        assert!(tags.len() >= n * max_tags && bearings.len() >= n * max_tags * 12 && gyro.len() >= n && out.len() >= n && ok.len() >= n);
        let r2c = to_iso(robot_to_cam);
        let ctx = self.ctx();
        check(ctx, unsafe { cb_sqpnp_batch(ctx, tags.as_ptr(), bearings.as_ptr(), n_tags.as_ptr(), max_tags as c_int, &r2c, gyro.as_ptr(),
                                           sign_change_error, n as i64, out.as_mut_ptr(), ok.as_mut_ptr()) })
    }
    /// `SqPnP::create_solver_camera_transform` (lib.rs:430-461).
    pub fn create_solver_camera_transform(fwd_m: f64, left_m: f64, up_m: f64, roll_deg: f64, pitch_deg: f64, yaw_deg: f64) -> Isometry3<f64> {
        let mut o = cb_iso3::default();
        unsafe { cb_create_solver_camera_transform(fwd_m, left_m, up_m, roll_deg, pitch_deg, yaw_deg, &mut o) };
        from_iso(&o)
    }
    /// `GenericModel::unproject` for `OpenCVModel5` (crates/apriltags/src/lib.rs:316-321): `None` where the iteration fails.
    pub fn unproject_opencv5(&mut self, params9: &[f64; 9], px: &[[f64; 2]]) -> Vec<Option<Vector3<f64>>> {
        let mut b = vec![0.0f64; px.len() * 3];
        let mut ok = vec![0u8; px.len()];
        let ctx = self.ctx();
        unsafe { cb_unproject_opencv5(ctx, params9.as_ptr(), px.as_ptr() as *const f64, px.len() as i64, b.as_mut_ptr(), ok.as_mut_ptr()) };
        (0..px.len()).map(|i| if ok[i] != 0 { Some(Vector3::new(b[3 * i], b[3 * i + 1], b[3 * i + 2])) } else { None }).collect()
    }
}
impl Drop for SqPnP { fn drop(&mut self) { if !self.ctx.is_null() { unsafe { cb_destroy(self.ctx) } } } }

/// `AprilTags::process` on the device (crates/apriltags/src/lib.rs:293-379): set the field layout and the camera once, then every
/// call returns the detection list AND `Some / None` robot pose per frame; detections never leave the device in between.
pub struct PoseDetector { det: Detector }
impl PoseDetector {
    pub fn new(det: Detector, field: &[(i32, Isometry3<f64>)], calib9: &[f64; 9], robot_to_cam: Option<&Isometry3<f64>>) -> Result<Self, Error> {
        let ids: Vec<i32> = field.iter().map(|(i, _)| *i).collect();
        let poses: Vec<cb_iso3> = field.iter().map(|(_, p)| to_iso(p)).collect();
        check(det.ctx, unsafe { cb_set_field(det.ctx, ids.as_ptr(), poses.as_ptr(), ids.len() as c_int) })?;
        let r2c = robot_to_cam.map(to_iso);
        check(det.ctx, unsafe { cb_set_camera(det.ctx, calib9.as_ptr(), r2c.as_ref().map_or(std::ptr::null(), |r| r as *const cb_iso3)) })?;
        Ok(Self { det })
    }
    /// One frame: `(detections, Some((rot, pos, std_devs)) | None)`; `gyro = None` is `comm.gyro_angle() == None` (lib.rs:329).
    pub fn process(&mut self, image: &Image, gyro: Option<f64>, sign_change_error: f64)
                   -> Result<(Vec<Detection>, Option<(Rotation3<f64>, Vector3<f64>, Vector3<f64>)>), Error> {
        let d = &mut self.det;
        d.out.resize(d.max_dets, unsafe { std::mem::zeroed() });
        let (mut count, mut used, mut ok) = (0i32, 0i32, 0u8);
        let mut pose: cb_pose = unsafe { std::mem::zeroed() };
        let g = gyro.unwrap_or(f64::NAN);
        let rc = unsafe { cb_detect_pose_gray(d.ctx, image.buf.as_ptr(), image.width, image.height, image.stride,
                                              (image.stride as usize) * (image.height as usize), 1, &g, sign_change_error,
                                              d.out.as_mut_ptr(), &mut count, &mut pose, &mut ok, &mut used) };
        if rc == CB_ERR_OVERFLOW { return Ok((Vec::new(), None)); }      // cluttered frame: the caller publishes its heartbeat
        check(d.ctx, rc)?;
        let dets = d.out[..count as usize].iter().map(|x| Detection(*x)).collect();
        let res = (ok != 0).then(|| (Rotation3::from_matrix_unchecked(Matrix3::from_column_slice(&pose.rot)), Vector3::from(pose.pos), Vector3::from(pose.std_devs)));
        Ok((dets, res))
    }
    /// Streaming form with two batches in flight (same rules as `Detector::submit` / `collect`).
    #[allow(clippy::too_many_arguments)]
    pub fn submit(&mut self, frames: &[u8], width: i32, height: i32, stride: i32, frame_stride: usize, gyro: &[f64], sign_change_error: f64) -> Result<(), Error> {
        check(self.det.ctx, unsafe { cb_detect_pose_gray_submit(self.det.ctx, frames.as_ptr(), width, height, stride, frame_stride,
                                                                gyro.len() as c_int, gyro.as_ptr(), sign_change_error) })
    }
    pub fn collect(&mut self, out: &mut [cb_detection], counts: &mut [i32], poses: &mut [cb_pose], pose_ok: &mut [u8], pose_tags: &mut [i32]) -> Result<(), Error> {
        check(self.det.ctx, unsafe { cb_detect_pose_gray_collect(self.det.ctx, out.as_mut_ptr(), counts.as_mut_ptr(), poses.as_mut_ptr(),
                                                                 pose_ok.as_mut_ptr(), pose_tags.as_mut_ptr()) })
    }
    /// The 64-byte records `AprilTags::process` publishes (lib.rs:340-376), one per frame.
    pub fn pack(poses: &[cb_pose], pose_ok: &[u8], det_counts: &[i32], ts_us: &[u64], camera_id: u8) -> Vec<cb_vision_measurement> {
        let mut out = vec![cb_vision_measurement::default(); poses.len()];
        unsafe { cb_pack_vision_measurements(poses.as_ptr(), pose_ok.as_ptr(), det_counts.as_ptr(), ts_us.as_ptr(), camera_id, poses.len() as c_int, out.as_mut_ptr()) };
        out
    }
}

/// The in-house CAT detector (crates/chalkydri-apriltags/src/lib.rs:142-181).
pub mod cat {
    use super::*;

    #[derive(Clone, Copy, Debug, PartialEq, Eq, PartialOrd, Ord)]
    #[repr(u8)]
    pub enum Color { Black = 0, White = 1, Other = 2 }     // utils.rs:1-6

    /// What `connected_components` returns (lib.rs:42-113): `parent[i]` is the smallest pixel index of `i`'s component.
    #[derive(Clone, Debug)]
    pub struct UnionFind { pub parent: Vec<u32>, pub cluster_sizes: Vec<u32> }
    impl UnionFind {
        pub fn find(&self, idx: usize) -> usize { self.parent[idx] as usize }
        pub fn get_size(&self, idx: usize) -> usize { self.cluster_sizes[idx] as usize }
    }

    pub struct Detector {
        ctx: *mut cb_ctx,
        det_ctx: *mut cb_ctx,                  // detect_tags(): a second context sized for the undecimated decode stages
        width: usize,
        height: usize,
        valid_tags: &'static [usize],
        pub buf: Vec<u8>,                      // Color map
        pub points: Vec<(usize, usize)>,
        pub lines: Vec<(usize, usize, usize, usize)>,
    }
    unsafe impl Send for Detector {}
    impl Detector {
        /// `Detector::new(width, height, valid_tags)` (lib.rs:158).
        pub fn new(width: usize, height: usize, valid_tags: &'static [usize]) -> Self {
            let ctx = unsafe { cb_create(0, 8, 8, 1, 1) };
            assert!(!ctx.is_null(), "chalkydri_b200: {}", last_error(std::ptr::null()));
            Self { ctx, det_ctx: std::ptr::null_mut(), width, height, valid_tags, buf: vec![0; width * height], points: Vec::new(), lines: Vec::new() }
        }
        /// `process_frame(&mut self, input: &[u8])` (lib.rs:265-287): packed RGB; asserts the length like lib.rs:267.  One library
        /// call: the frame is uploaded once, the intermediate maps stay on the device.
        pub fn process_frame(&mut self, input: &[u8]) {
            assert_eq!(input.len(), self.width * self.height * 3);
            const CAP: usize = 1 << 20;
            let mut xy = vec![0i32; CAP * 2];
            let mut ln = vec![0i32; CAP * 4];
            let (mut n, mut m) = (0i64, 0i64);
            let rc = unsafe { cb_cat_process_frame(self.ctx, input.as_ptr(), self.width as c_int, self.height as c_int, self.buf.as_mut_ptr(),
                                                   xy.as_mut_ptr(), CAP as i64, &mut n, ln.as_mut_ptr(), CAP as i64, &mut m) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
            self.points = (0..n as usize).map(|i| (xy[2 * i] as usize, xy[2 * i + 1] as usize)).collect();
            self.lines = (0..m as usize).map(|i| (ln[4 * i] as usize, ln[4 * i + 1] as usize, ln[4 * i + 2] as usize, ln[4 * i + 3] as usize)).collect();
        }
        pub fn calc_otsu(&mut self, input: &[u8]) {
            let rc = unsafe { cb_cat_calc_otsu(self.ctx, input.as_ptr(), self.width as c_int, self.height as c_int, self.buf.as_mut_ptr()) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
        }
        pub fn thresh(&mut self, input: &[u8]) {
            let rc = unsafe { cb_cat_thresh(self.ctx, input.as_ptr(), self.width as c_int, self.height as c_int, self.buf.as_mut_ptr()) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
        }
        pub fn detect_corners(&mut self) {
            const CAP: usize = 1 << 20;
            let mut xy = vec![0i32; CAP * 2];
            let mut n = 0i64;
            let rc = unsafe { cb_cat_detect_corners(self.ctx, self.buf.as_ptr(), self.width as c_int, self.height as c_int, xy.as_mut_ptr(), CAP as i64, &mut n) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
            self.points = (0..(n as usize).min(CAP)).map(|i| (xy[2 * i] as usize, xy[2 * i + 1] as usize)).collect();
        }
        pub fn check_edges(&mut self) {
            const CAP: usize = 1 << 20;
            let xy: Vec<i32> = self.points.iter().flat_map(|&(x, y)| [x as i32, y as i32]).collect();
            let mut ln = vec![0i32; CAP * 4];
            let mut m = 0i64;
            let rc = unsafe { cb_cat_check_edges(self.ctx, self.buf.as_ptr(), self.width as c_int, self.height as c_int, xy.as_ptr(),
                                                 self.points.len() as i64, ln.as_mut_ptr(), CAP as i64, &mut m) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
            self.lines = (0..(m as usize).min(CAP)).map(|i| (ln[4 * i] as usize, ln[4 * i + 1] as usize, ln[4 * i + 2] as usize, ln[4 * i + 3] as usize)).collect();
        }
        /// The decode the reference intends for CAT (book/src/maintenance/apriltags.md:58-60, lib.rs:551-613): CAT's own ternary map
        /// (`thresh`, or `calc_otsu`), then the C library's stages on it, in one library call.  Only `valid_tags` when that list is not empty.
        pub fn detect_tags(&mut self, input: &[u8], use_otsu: bool) -> Vec<cb_detection> {
            assert_eq!(input.len(), self.width * self.height * 3);
            const MAX_DETS: usize = 64;
            if self.det_ctx.is_null() {                       // the decode stages run undecimated: capacity of twice the frame size
                self.det_ctx = unsafe { cb_create(0, 2 * self.width as c_int, 2 * self.height as c_int, 1, MAX_DETS as c_int) };
                assert!(!self.det_ctx.is_null(), "chalkydri_b200: {}", last_error(std::ptr::null()));
                let rc = unsafe { cb_set_family_tag36h11(self.det_ctx, 3) };
                if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.det_ctx)); }
            }
            let mut out: Vec<cb_detection> = vec![unsafe { std::mem::zeroed() }; MAX_DETS];
            let mut n = 0i32;
            let rc = unsafe { cb_cat_detect_tags(self.det_ctx, input.as_ptr(), self.width as c_int, self.height as c_int, use_otsu as c_int,
                                                 out.as_mut_ptr(), &mut n) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.det_ctx)); }
            out.truncate(n as usize);
            if !self.valid_tags.is_empty() { out.retain(|d| self.valid_tags.contains(&(d.id as usize))); }
            out
        }
        /// `connected_components(&self) -> UnionFind` (lib.rs:501).
        pub fn connected_components(&self) -> UnionFind {
            let n = self.width * self.height;
            let mut uf = UnionFind { parent: vec![0; n], cluster_sizes: vec![0; n] };
            let rc = unsafe { cb_cat_connected_components(self.ctx, self.buf.as_ptr(), self.width as c_int, self.height as c_int,
                                                          uf.parent.as_mut_ptr(), uf.cluster_sizes.as_mut_ptr()) };
            if rc != CB_OK { panic!("chalkydri_b200: {}", last_error(self.ctx)); }
            uf
        }
    }
    impl Clone for Detector {
        /// lib.rs:663-667: cloning makes a fresh, empty detector of the same size.
        fn clone(&self) -> Self { Self::new(self.width, self.height, self.valid_tags) }
    }
    impl Drop for Detector { fn drop(&mut self) { unsafe { cb_destroy(self.ctx); if !self.det_ctx.is_null() { cb_destroy(self.det_ctx) } } } }
}

/// Several GPUs of one box from one process: one context + one host thread per GPU inside the library, every GPU's lists land in
/// its slice of the caller's one array (`cb_pool_detect_gray`); no collective.
pub struct DetectorPool { pool: *mut cb_pool, max_dets: usize }
unsafe impl Send for DetectorPool {}
impl DetectorPool {
    pub fn new(devices: &[i32], max_w: i32, max_h: i32, max_batch: i32, max_dets: i32, bits_corrected: usize) -> Result<Self, Error> {
        let pool = unsafe { cb_pool_create(if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() }, devices.len() as c_int, max_w, max_h, max_batch, max_dets) };
        if pool.is_null() { return Err(Error(CB_ERR_CUDA, unsafe { CStr::from_ptr(cb_pool_last_error(std::ptr::null())) }.to_string_lossy().into_owned())); }
        let p = Self { pool, max_dets: max_dets as usize };
        p.check(unsafe { cb_pool_set_family_tag36h11(pool, bits_corrected as c_int) })?;
        Ok(p)
    }
    fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == CB_OK { Ok(()) } else { Err(Error(rc, unsafe { CStr::from_ptr(cb_pool_last_error(self.pool)) }.to_string_lossy().into_owned())) }
    }
    pub fn len(&self) -> usize { unsafe { cb_pool_size(self.pool) as usize } }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn context(&mut self, i: usize) -> *mut cb_ctx { unsafe { cb_pool_context(self.pool, i as c_int) } }
    pub fn detect(&mut self, frames: &[u8], width: i32, height: i32, stride: i32, frame_stride: usize, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), Error> {
        assert!(out.len() >= counts.len() * self.max_dets);
        self.check(unsafe { cb_pool_detect_gray(self.pool, frames.as_ptr(), width, height, stride, frame_stride, counts.len() as c_int, out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
    pub fn timing(&self) -> cb_pool_timing { let mut t = cb_pool_timing::default(); unsafe { cb_pool_get_timing(self.pool, &mut t) }; t }
}
impl Drop for DetectorPool { fn drop(&mut self) { unsafe { cb_pool_destroy(self.pool) } } }

/// Pinned host memory for frames (`cb_host_alloc`): lets the H2D copy of batch k+1 run under the kernels of batch k.
pub struct PinnedBuffer { ptr: *mut u8, len: usize }
unsafe impl Send for PinnedBuffer {}
impl PinnedBuffer {
    pub fn new(len: usize) -> Option<Self> {
        let ptr = unsafe { cb_host_alloc(len) } as *mut u8;
        (!ptr.is_null()).then_some(Self { ptr, len })
    }
    pub fn as_slice(&self) -> &[u8] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [u8] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
}
impl Drop for PinnedBuffer { fn drop(&mut self) { unsafe { cb_host_free(self.ptr as *mut std::os::raw::c_void) } } }

/// Device-resident frames (`cb_device_alloc` + `cb_detect_gray_device`): for producers that already write into GPU memory.
pub struct DeviceFrames { ctx: *mut cb_ctx, ptr: *mut u8, len: usize }
impl DeviceFrames {
    pub fn new(det: &Detector, len: usize) -> Option<Self> {
        let ptr = unsafe { cb_device_alloc(det.ctx, len) } as *mut u8;
        (!ptr.is_null()).then_some(Self { ctx: det.ctx, ptr, len })
    }
    pub fn upload(&mut self, src: &[u8]) -> Result<(), Error> {
        assert!(src.len() <= self.len);
        check(self.ctx, unsafe { cb_memcpy_h2d(self.ctx, self.ptr as *mut _, src.as_ptr() as *const _, src.len()) })
    }
    pub fn download(&self, dst: &mut [u8]) -> Result<(), Error> {
        assert!(dst.len() <= self.len);
        check(self.ctx, unsafe { cb_memcpy_d2h(self.ctx, dst.as_mut_ptr() as *mut _, self.ptr as *const _, dst.len()) })
    }
    #[allow(clippy::too_many_arguments)]
    pub fn detect(&self, det: &mut Detector, width: i32, height: i32, stride: i32, frame_stride: usize, out: &mut [cb_detection], counts: &mut [i32]) -> Result<(), Error> {
        check(det.ctx, unsafe { cb_detect_gray_device(det.ctx, self.ptr, width, height, stride, frame_stride, counts.len() as c_int, out.as_mut_ptr(), counts.as_mut_ptr()) })
    }
}
impl Drop for DeviceFrames { fn drop(&mut self) { unsafe { cb_device_free(self.ctx, self.ptr as *mut _) } } }

/// Pre-processing taps: packed RGB -> gray with CAT's formula (utils.rs:33-46), YUYV -> Y.
pub fn rgb_to_gray(det: &mut Detector, rgb: &[u8], width: i32, height: i32, batch: i32, gray: &mut [u8]) -> Result<(), Error> {
    check(det.ctx, unsafe { cb_rgb_to_gray(det.ctx, rgb.as_ptr(), width, height, batch, gray.as_mut_ptr()) })
}
pub fn yuyv_to_gray(det: &mut Detector, yuyv: &[u8], width: i32, height: i32, batch: i32, gray: &mut [u8]) -> Result<(), Error> {
    check(det.ctx, unsafe { cb_yuyv_to_gray(det.ctx, yuyv.as_ptr(), width, height, batch, gray.as_mut_ptr()) })
}
/// Problems whose arrays already live in device memory.
#[allow(clippy::too_many_arguments)]
pub unsafe fn sqpnp_batch_device(ctx: *mut cb_ctx, tags: *const cb_iso3, bearings: *const f64, n_tags: *const i32, max_tags: i32,
                                 robot_to_cam: *const cb_iso3, gyro: *const f64, sign_change_error: f64, n: i64, out: *mut cb_pose, ok: *mut u8) -> Result<(), Error> {
    check(ctx, unsafe { cb_sqpnp_batch_device(ctx, tags, bearings, n_tags, max_tags, robot_to_cam, gyro, sign_change_error, n, out, ok) })
}
pub fn device_count() -> i32 { unsafe { cb_device_count() } }
pub fn version() -> String { unsafe { CStr::from_ptr(cb_version()) }.to_string_lossy().into_owned() }
