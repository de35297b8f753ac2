"""Host-side mirror of the in-house CAT detector (/root/reference/crates/chalkydri-apriltags/src/lib.rs) above the C ABI.

`Detector::new(width, height, valid_tags)` (lib.rs:158), `process_frame(&[u8] packed RGB)` (lib.rs:265; asserts the
length like lib.rs:267), `calc_otsu`, `thresh`, `detect_corners`, `check_edges`, `connected_components` (lib.rs:501).
State mirrors the reference: `buf` (Color map 0/1/2), `points`, `lines`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import ChalkydriError

BLACK, WHITE, OTHER = 0, 1, 2     # utils.rs:1-6


class UnionFind:
    """Result of connected_components(): labels = smallest pixel index of each component, sizes per pixel."""

    def __init__(self, labels: np.ndarray, sizes: np.ndarray):
        self.parent, self.cluster_sizes = labels, sizes

    def find(self, idx: int) -> int:
        return int(self.parent.reshape(-1)[idx])

    def get_size(self, idx: int) -> int:
        return int(self.cluster_sizes.reshape(-1)[idx])


class CatDetector:
    def __init__(self, width: int, height: int, valid_tags=(), device: int = 0):
        self._L = capi.lib()
        self._ctx = self._L.cb_create(device, 8, 8, 1, 1)
        if not self._ctx:
            raise ChalkydriError(capi.CB_ERR_CUDA, self._L.cb_last_error(None).decode())
        self.width, self.height, self.valid_tags = int(width), int(height), tuple(valid_tags)
        self.buf = np.zeros((height, width), np.uint8)      # alloc_zeroed: all Black (lib.rs:167)
        self.points = np.zeros((0, 2), np.int32)
        self.lines = np.zeros((0, 4), np.int32)

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.cb_destroy(self._ctx)
            self._ctx = None
        if getattr(self, "_det_ctx", None):
            self._L.cb_destroy(self._det_ctx)
            self._det_ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ChalkydriError(rc, self._L.cb_last_error(self._ctx).decode())

    def _rgb(self, input_):
        a = np.ascontiguousarray(input_, np.uint8).reshape(-1)
        assert a.size == self.width * self.height * 3, "input must be packed RGB (lib.rs:267)"
        return a

    def calc_otsu(self, input_):
        self._check(self._L.cb_cat_calc_otsu(self._ctx, capi.ptr(self._rgb(input_)), self.width, self.height, capi.ptr(self.buf)))

    def thresh(self, input_):
        self._check(self._L.cb_cat_thresh(self._ctx, capi.ptr(self._rgb(input_)), self.width, self.height, capi.ptr(self.buf)))

    def detect_corners(self, cap: int = 1 << 22):
        xy = np.zeros((cap, 2), np.int32)
        n = C.c_int64()
        self._check(self._L.cb_cat_detect_corners(self._ctx, capi.ptr(self.buf), self.width, self.height, capi.ptr(xy), cap, C.byref(n)))
        if n.value > cap:
            raise ChalkydriError(capi.CB_ERR_OVERFLOW, f"{n.value} corners exceed the capacity {cap}")
        self.points = xy[:n.value].copy()

    def check_edges(self, cap: int = 1 << 22):
        lines = np.zeros((cap, 4), np.int32)
        n = C.c_int64()
        self._check(self._L.cb_cat_check_edges(self._ctx, capi.ptr(self.buf), self.width, self.height, capi.ptr(self.points),
                                               len(self.points), capi.ptr(lines), cap, C.byref(n)))
        if n.value > cap:
            raise ChalkydriError(capi.CB_ERR_OVERFLOW, f"{n.value} lines exceed the capacity {cap}")
        self.lines = lines[:n.value].copy()

    def process_frame(self, input_, cap: int = 1 << 20, want_color: bool = True):
        """lib.rs:265-287: calc_otsu -> reset points / lines -> detect_corners -> check_edges, in ONE library call
        (cb_cat_process_frame): the frame goes up once and the intermediate maps stay on the device."""
        rgb = self._rgb(input_)
        if getattr(self, "_xy_buf", None) is None or len(self._xy_buf) < cap:
            self._xy_buf, self._lines_buf = np.zeros((cap, 2), np.int32), np.zeros((cap, 4), np.int32)
        n, m = C.c_int64(), C.c_int64()
        self._check(self._L.cb_cat_process_frame(self._ctx, capi.ptr(rgb), self.width, self.height, capi.ptr(self.buf) if want_color else None,
                                                 capi.ptr(self._xy_buf), cap, C.byref(n), capi.ptr(self._lines_buf), cap, C.byref(m)))
        self.points = self._xy_buf[:n.value].copy()
        self.lines = self._lines_buf[:m.value].copy()

    def detect_tags(self, input_, use_otsu: bool = False, max_dets: int = 64):
        """The decode the reference intends for CAT (book/src/maintenance/apriltags.md:58-60, lib.rs:551-613): CAT's own ternary map
        (`thresh`, or `calc_otsu`), then the C library's stages on it -- in one library call (cb_cat_detect_tags).  Returns the
        detection records; with `valid_tags` given to the constructor only those ids."""
        rgb = self._rgb(input_)
        if getattr(self, "_det_ctx", None) is None:           # the decode stages run undecimated: capacity of twice the frame size
            self._det_ctx = self._L.cb_create(0, 2 * self.width, 2 * self.height, 1, max_dets)
            if not self._det_ctx:
                raise ChalkydriError(capi.CB_ERR_CUDA, self._L.cb_last_error(None).decode())
            self._det_cap = max_dets
            rc = self._L.cb_set_family_tag36h11(self._det_ctx, 3)
            if rc:
                raise ChalkydriError(rc, self._L.cb_last_error(self._det_ctx).decode())
        out = np.zeros(self._det_cap, capi.DET_DTYPE)
        n = C.c_int32()
        rc = self._L.cb_cat_detect_tags(self._det_ctx, capi.ptr(rgb), self.width, self.height, 1 if use_otsu else 0, capi.ptr(out), C.byref(n))
        if rc:
            raise ChalkydriError(rc, self._L.cb_last_error(self._det_ctx).decode())
        dets = out[:n.value]
        if self.valid_tags:
            dets = dets[np.isin(dets["id"], self.valid_tags)]
        return dets

    def process_frame_stagewise(self, input_):
        """the same through the per-stage entry points (host buffers between the stages); kept for the parity tests"""
        self.calc_otsu(input_)
        self.points = np.zeros((0, 2), np.int32)
        self.lines = np.zeros((0, 4), np.int32)
        self.detect_corners()
        self.check_edges()

    def timing(self) -> dict:
        t = capi.Timing()
        self._check(self._L.cb_get_timing(self._ctx, C.byref(t)))
        return t.as_dict()

    def connected_components(self) -> UnionFind:
        lab = np.empty((self.height, self.width), np.uint32)
        sz = np.empty((self.height, self.width), np.uint32)
        self._check(self._L.cb_cat_connected_components(self._ctx, capi.ptr(self.buf), self.width, self.height, capi.ptr(lab), capi.ptr(sz)))
        return UnionFind(lab, sz)
