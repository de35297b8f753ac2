"""Host-side mirror of the detector interface the reference uses (Python above the C ABI).

Reference seam (seam 1): `apriltag::DetectorBuilder::default().add_family_bits(family, bits).build()`
and `Detector::detect(&Image) -> Vec<Detection>` with `Detection::{id, hamming, decision_margin, corners, center,
homography}` (/root/reference/crates/apriltags/src/lib.rs:19,258-261,301-314).  Same names, argument meaning and
error behaviour (unknown family / bad bits raise at build time, like the `unwrap()`s at lib.rs:229,261); the batch
entry points are the B200 addition.  All compute happens in libchalkydri_b200.so (CUDA); nothing here falls back to CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import DET_DTYPE, ChalkydriError

FAMILY = "tag36h11"       # crates/apriltags/src/lib.rs:45
MAX_DETECTIONS = 16       # crates/apriltags/src/lib.rs:42 (the reference's own cap; capacity is a parameter here)


@dataclass
class Detection:
    """One `apriltag::Detection`."""
    _rec: np.void

    def id(self) -> int:
        return int(self._rec["id"])

    def hamming(self) -> int:
        return int(self._rec["hamming"])

    def decision_margin(self) -> float:
        return float(self._rec["decision_margin"])

    def corners(self):
        return [[float(x), float(y)] for x, y in self._rec["p"]]

    def center(self):
        return [float(self._rec["c"][0]), float(self._rec["c"][1])]

    def homography(self) -> np.ndarray:
        return np.array(self._rec["H"], np.float64).reshape(3, 3)


class Image:
    """image_u8_t view {buf, width, height, stride} (image_from_cuimage, lib.rs:197-213): borrows the pixels."""

    def __init__(self, buf: np.ndarray, width: int | None = None, height: int | None = None, stride: int | None = None):
        buf = np.asarray(buf, np.uint8)
        if buf.ndim == 2 and width is None:
            height, width = buf.shape
            stride = buf.strides[0]
            if buf.strides[1] != 1:
                buf = np.ascontiguousarray(buf)
                stride = width
        self.buf, self.width, self.height, self.stride = buf, int(width), int(height), int(stride if stride else width)


class DetectorBuilder:
    def __init__(self):
        self._families = []
        self._device = 0
        self._max_size = None
        self._max_batch = 1
        self._max_dets = 64

    @staticmethod
    def default():
        return DetectorBuilder()

    def add_family_bits(self, family: str, bits_corrected: int):
        if family != FAMILY:
            raise ValueError(f"unknown family {family!r}: this build carries tag36h11 only (the reference's FAMILY)")
        self._families.append((family, int(bits_corrected)))
        return self

    # B200 additions (capacity of the device context)
    def device(self, index: int):
        self._device = int(index)
        return self

    def capacity(self, max_width: int, max_height: int, max_batch: int = 1, max_dets_per_frame: int = 64):
        self._max_size = (int(max_width), int(max_height))
        self._max_batch, self._max_dets = int(max_batch), int(max_dets_per_frame)
        return self

    def build(self) -> "Detector":
        if not self._families:
            raise ValueError("no tag family added")
        if self._max_size is None:
            raise ValueError("call capacity(max_width, max_height, ...) before build(): device buffers are sized once")
        return Detector(self._device, self._max_size[0], self._max_size[1], self._max_batch, self._max_dets, self._families[-1][1])


class Detector:
    def __init__(self, device: int, max_width: int, max_height: int, max_batch: int, max_dets: int, bits_corrected: int):
        L = capi.lib()
        self._L = L
        self._ctx = L.cb_create(device, max_width, max_height, max_batch, max_dets)
        if not self._ctx:
            raise ChalkydriError(capi.CB_ERR_CUDA, L.cb_last_error(None).decode())
        self.max_batch, self.max_dets = max_batch, max_dets
        self.device = device
        self._inflight = []          # (frames, batch) of submitted batches: keeps the host arrays alive until collect()
        self._check(L.cb_set_family_tag36h11(self._ctx, bits_corrected))

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.cb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ChalkydriError(rc, self._L.cb_last_error(self._ctx).decode())

    def set_params(self, quad_decimate=2.0, quad_sigma=0.0, refine_edges=1, decode_sharpening=0.25, min_cluster_pixels=5,
                   max_nmaxima=10, critical_rad=float(np.float32(10 * np.pi / 180)), max_line_fit_mse=10.0, min_white_black_diff=5):
        self._check(self._L.cb_set_params(self._ctx, quad_decimate, quad_sigma, refine_edges, decode_sharpening, min_cluster_pixels,
                                          max_nmaxima, critical_rad, max_line_fit_mse, min_white_black_diff))

    # ---- the reference call: one frame in, list of detections out ----
    def detect(self, image) -> list:
        if not isinstance(image, Image):
            image = Image(image)
        frames = image.buf
        out = np.zeros(self.max_dets, DET_DTYPE)
        counts = np.zeros(1, np.int32)
        self._check(self._L.cb_detect_gray(self._ctx, capi.ptr(frames), image.width, image.height, image.stride,
                                           image.stride * image.height, 1, capi.ptr(out), capi.ptr(counts)))
        return [Detection(out[i]) for i in range(int(counts[0]))]

    # ---- batched entry points ----
    def detect_batch(self, frames: np.ndarray, out: np.ndarray | None = None, counts: np.ndarray | None = None):
        """frames [B,H,W] u8 in host memory -> (detections [B,max_dets] structured, counts [B])."""
        if frames.dtype != np.uint8 or frames.ndim != 3 or not frames.flags.c_contiguous:
            raise ValueError("frames must be a C-contiguous [B,H,W] uint8 array")
        B, H, W = frames.shape
        if out is None:
            out = np.zeros((B, self.max_dets), DET_DTYPE)
        if counts is None:
            counts = np.zeros(B, np.int32)
        self._check(self._L.cb_detect_gray(self._ctx, capi.ptr(frames), W, H, W, H * W, B, capi.ptr(out), capi.ptr(counts)))
        return out, counts

    # ---- streaming form: submit batch k+1 before collecting batch k, and its H2D copy runs under batch k's kernels ----
    def submit(self, frames: np.ndarray):
        """Enqueue frames [B,H,W] u8 (B <= max_batch; keep the array alive and unchanged until collect())."""
        if frames.dtype != np.uint8 or frames.ndim != 3 or not frames.flags.c_contiguous:
            raise ValueError("frames must be a C-contiguous [B,H,W] uint8 array")
        B, H, W = frames.shape
        self._check(self._L.cb_detect_gray_submit(self._ctx, capi.ptr(frames), W, H, W, H * W, B))
        self._inflight.append((frames, B))

    def collect(self, out: np.ndarray | None = None, counts: np.ndarray | None = None):
        """Wait for the oldest submitted batch -> (detections [B,max_dets], counts [B]) like detect_batch."""
        B = self._inflight[0][1] if self._inflight else 1          # nothing in flight: the library reports CB_ERR_STATE
        if out is None:
            out = np.zeros((B, self.max_dets), DET_DTYPE)
        if counts is None:
            counts = np.zeros(B, np.int32)
        try:
            self._check(self._L.cb_detect_gray_collect(self._ctx, capi.ptr(out), capi.ptr(counts)))
        finally:
            if self._inflight:
                self._inflight.pop(0)
        return out, counts

    @property
    def pending(self) -> int:
        return int(self._L.cb_detect_gray_pending(self._ctx))

    def detect_batch_device(self, dev_ptr: int, B: int, H: int, W: int, stride: int | None = None, frame_stride: int | None = None,
                            out: np.ndarray | None = None, counts: np.ndarray | None = None):
        stride = stride or W
        frame_stride = frame_stride or stride * H
        if out is None:
            out = np.zeros((B, self.max_dets), DET_DTYPE)
        if counts is None:
            counts = np.zeros(B, np.int32)
        self._check(self._L.cb_detect_gray_device(self._ctx, C.c_void_p(dev_ptr), W, H, stride, frame_stride, B, capi.ptr(out), capi.ptr(counts)))
        return out, counts

    def detect_rgb_batch(self, frames_rgb: np.ndarray):
        B, H, W, ch = frames_rgb.shape
        assert ch == 3 and frames_rgb.dtype == np.uint8 and frames_rgb.flags.c_contiguous
        out = np.zeros((B, self.max_dets), DET_DTYPE)
        counts = np.zeros(B, np.int32)
        self._check(self._L.cb_detect_rgb(self._ctx, capi.ptr(frames_rgb), W, H, B, capi.ptr(out), capi.ptr(counts)))
        return out, counts

    def detect_yuyv_batch(self, frames_yuyv: np.ndarray):
        B, H, W2 = frames_yuyv.shape
        W = W2 // 2
        assert frames_yuyv.dtype == np.uint8 and frames_yuyv.flags.c_contiguous
        out = np.zeros((B, self.max_dets), DET_DTYPE)
        counts = np.zeros(B, np.int32)
        self._check(self._L.cb_detect_yuyv(self._ctx, capi.ptr(frames_yuyv), W, H, B, capi.ptr(out), capi.ptr(counts)))
        return out, counts

    def detect_yuv420_batch(self, frames_yuv: np.ndarray):
        """NV12 / NV21 / I420 / YV12 buffers [B, H*3/2, W] (gst_to_cu.rs:152-188): the Y plane is the gray image."""
        B, H32, W = frames_yuv.shape
        H = H32 * 2 // 3
        assert frames_yuv.dtype == np.uint8 and frames_yuv.flags.c_contiguous and H * 3 // 2 == H32
        out = np.zeros((B, self.max_dets), DET_DTYPE)
        counts = np.zeros(B, np.int32)
        self._check(self._L.cb_detect_yuv420(self._ctx, capi.ptr(frames_yuv), W, H, B, capi.ptr(out), capi.ptr(counts)))
        return out, counts

    # ---- stage taps (parity tests) ----
    def rgb_to_gray(self, frames_rgb: np.ndarray) -> np.ndarray:
        """pre-processing tap: packed RGB [B,H,W,3] -> gray [B,H,W] (utils.rs:43)"""
        frames_rgb = np.ascontiguousarray(frames_rgb, np.uint8)
        B, H, W, _ = frames_rgb.shape
        out = np.empty((B, H, W), np.uint8)
        self._check(self._L.cb_rgb_to_gray(self._ctx, capi.ptr(frames_rgb), W, H, B, capi.ptr(out)))
        return out

    def yuyv_to_gray(self, frames_yuyv: np.ndarray) -> np.ndarray:
        frames_yuyv = np.ascontiguousarray(frames_yuyv, np.uint8)
        B, H, W2 = frames_yuyv.shape
        out = np.empty((B, H, W2 // 2), np.uint8)
        self._check(self._L.cb_yuyv_to_gray(self._ctx, capi.ptr(frames_yuyv), W2 // 2, H, B, capi.ptr(out)))
        return out

    def decimated_size(self, W, H):
        w, h = C.c_int(), C.c_int()
        self._check(self._L.cb_decimated_size(self._ctx, W, H, C.byref(w), C.byref(h)))
        return w.value, h.value

    def frame_flags(self, n: int) -> np.ndarray:
        """Overflow bits per frame of the last completed detection call (cb_frame_flags): a flagged frame reported an empty list."""
        f = np.zeros(n, np.uint32)
        rc = self._L.cb_frame_flags(self._ctx, capi.ptr(f), n)
        if rc < 0:
            self._check(rc)
        return f

    def clusters(self, frames: np.ndarray, cap: int = 1 << 21):
        """Stage tap of gradient_clusters(): (pts [n,4] int16 = x, y, gx, gy as upstream stores them, cluster_of [n], nclusters [B]);
        the selected clusters only, points in upstream's append order."""
        B, H, W = frames.shape
        pts = np.zeros((cap, 4), np.int16); cl = np.zeros(cap, np.int32); ncl = np.zeros(B, np.int32)
        n = C.c_int64()
        self._check(self._L.cb_clusters(self._ctx, capi.ptr(frames), W, H, W, H * W, B, capi.ptr(pts), capi.ptr(cl), cap, C.byref(n), capi.ptr(ncl)))
        if n.value > cap:
            raise ChalkydriError(capi.CB_ERR_OVERFLOW, f"{n.value} points exceed the capacity {cap}")
        return pts[:n.value], cl[:n.value], ncl

    def threshold(self, frames: np.ndarray) -> np.ndarray:
        B, H, W = frames.shape
        w, h = self.decimated_size(W, H)
        out = np.empty((B, h, w), np.uint8)
        self._check(self._L.cb_threshold(self._ctx, capi.ptr(frames), W, H, W, H * W, B, capi.ptr(out)))
        return out

    def labels(self, frames: np.ndarray):
        B, H, W = frames.shape
        w, h = self.decimated_size(W, H)
        lab = np.empty((B, h, w), np.uint32)
        sz = np.empty((B, h, w), np.uint32)
        self._check(self._L.cb_labels(self._ctx, capi.ptr(frames), W, H, W, H * W, B, capi.ptr(lab), capi.ptr(sz)))
        return lab, sz

    def quads(self, frames: np.ndarray, cap: int = 4096):
        B, H, W = frames.shape
        q = np.zeros((B, cap, 4, 2), np.float32)
        counts = np.zeros(B, np.int32)
        npts = C.c_int64()
        self._check(self._L.cb_quads(self._ctx, capi.ptr(frames), W, H, W, H * W, B, capi.ptr(q), cap, capi.ptr(counts), C.byref(npts)))
        return q, counts, npts.value

    def timing(self) -> dict:
        t = capi.Timing()
        self._check(self._L.cb_get_timing(self._ctx, C.byref(t)))
        return t.as_dict()

    # raw context for the solver / CAT wrappers that share it
    @property
    def ctx(self):
        return self._ctx
