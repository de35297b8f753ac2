"""Several GPUs of one box from one process: `DetectorPool` shards a batch of frames over one detector context
per GPU with no collective -- host threads inside libchalkydri_b200.so feed each GPU its contiguous share through the streaming
form of the call and every GPU's lists land in its slice of one output array (cb_pool_detect_gray).

The reference runs one `AprilTags` task per camera in one process (crates/apriltags/src/lib.rs:166-182; three cameras in
chalkydri.ron:2-105); this is the same shape with the cameras' frames spread over GPUs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import DET_DTYPE, ChalkydriError


class DetectorPool:
    def __init__(self, devices=None, max_width=1280, max_height=720, max_batch=256, max_dets=64, bits_corrected=3):
        L = capi.lib()
        self._L = L
        if devices is None:
            arr, n = None, 0
        else:
            arr = np.ascontiguousarray(devices, np.int32)
            n = len(arr)
        self._pool = L.cb_pool_create(capi.ptr(arr), n, max_width, max_height, max_batch, max_dets)
        if not self._pool:
            raise ChalkydriError(capi.CB_ERR_CUDA, L.cb_pool_last_error(None).decode())
        self.max_batch, self.max_dets = max_batch, max_dets
        self._check(L.cb_pool_set_family_tag36h11(self._pool, bits_corrected))

    def _check(self, rc):
        if rc != 0:
            raise ChalkydriError(rc, self._L.cb_pool_last_error(self._pool).decode())

    def __len__(self):
        return int(self._L.cb_pool_size(self._pool))

    def detect_batch(self, frames: np.ndarray, out: np.ndarray | None = None, counts: np.ndarray | None = None):
        """frames [n,H,W] u8 (pinned for full speed) -> (detections [n,max_dets], counts [n]); n is not limited by max_batch."""
        if frames.dtype != np.uint8 or frames.ndim != 3 or not frames.flags.c_contiguous:
            raise ValueError("frames must be a C-contiguous [n,H,W] uint8 array")
        n, H, W = frames.shape
        if out is None:
            out = np.zeros((n, self.max_dets), DET_DTYPE)
        if counts is None:
            counts = np.zeros(n, np.int32)
        self._check(self._L.cb_pool_detect_gray(self._pool, capi.ptr(frames), W, H, W, H * W, n, capi.ptr(out), capi.ptr(counts)))
        return out, counts

    def timing(self) -> dict:
        t = capi.PoolTiming()
        self._check(self._L.cb_pool_get_timing(self._pool, C.byref(t)))
        return {n: getattr(t, n) for n, _ in t._fields_}

    def close(self):
        if getattr(self, "_pool", None):
            self._L.cb_pool_destroy(self._pool)
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
