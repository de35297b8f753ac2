"""ctypes binding of libchalkydri_b200.so (include/chalkydri_b200.h).

This is the only route from Python into the product: every function below is a thin call into the C ABI, which in turn
only launches CUDA kernels.  There is no CPU fallback; loading fails loudly when the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`), and cb_create fails when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchalkydri_b200.so")

CB_OK, CB_ERR_ARG, CB_ERR_CUDA, CB_ERR_UNSUPPORTED, CB_ERR_OVERFLOW, CB_ERR_STATE = 0, -1, -2, -3, -4, -5

DET_DTYPE = np.dtype([("frame", "<i4"), ("id", "<i4"), ("hamming", "<i4"), ("decision_margin", "<f4"),
                      ("H", "<f8", (9,)), ("c", "<f8", (2,)), ("p", "<f8", (4, 2))])
ISO_DTYPE = np.dtype([("t", "<f8", (3,)), ("q", "<f8", (4,))])
POSE_DTYPE = np.dtype([("rot", "<f8", (9,)), ("pos", "<f8", (3,)), ("std_devs", "<f8", (3,))])
# struct VisionMeasurement (crates/whacknet/src/lib.rs:40-66)
VISION_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("rot", "<f8"), ("std_x", "<f8"), ("std_y", "<f8"), ("std_rot", "<f8"), ("ts", "<u8"),
                         ("camera_id", "u1"), ("tag_count", "u1"), ("reserved", "u1", (6,))])
assert DET_DTYPE.itemsize == 168 and ISO_DTYPE.itemsize == 56 and POSE_DTYPE.itemsize == 120 and VISION_DTYPE.itemsize == 64


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "preprocess_ms", "threshold_ms", "ccl_ms", "cluster_ms", "quad_ms",
                                         "decode_ms", "d2h_ms", "total_ms")] + [("threshold_launches", C.c_int32),
                                                                                ("kernel_launches", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PoolTiming(C.Structure):
    _fields_ = [("wall_ms", C.c_float), ("max_device_ms", C.c_float), ("min_device_ms", C.c_float), ("n_devices", C.c_int32)]


# every symbol include/chalkydri_b200.h declares
EXPORTS = ["cb_create", "cb_destroy", "cb_last_error", "cb_set_family_tag36h11", "cb_set_params", "cb_detect_gray",
           "cb_detect_gray_device", "cb_detect_gray_submit", "cb_detect_gray_collect", "cb_detect_gray_pending", "cb_detect_rgb", "cb_detect_yuyv", "cb_detect_yuv420", "cb_rgb_to_gray", "cb_yuyv_to_gray", "cb_decimated_size", "cb_threshold", "cb_labels",
           "cb_quads", "cb_clusters", "cb_get_timing", "cb_frame_flags", "cb_sqpnp_set", "cb_sqpnp_batch", "cb_sqpnp_batch_device",
           "cb_create_solver_camera_transform", "cb_unproject_opencv5", "cb_set_field", "cb_set_camera", "cb_detect_pose_gray", "cb_detect_pose_gray_submit", "cb_detect_pose_gray_collect", "cb_pack_vision_measurements", "cb_cat_calc_otsu", "cb_cat_thresh",
           "cb_cat_detect_corners", "cb_cat_check_edges", "cb_cat_process_frame", "cb_cat_detect_tags", "cb_cat_connected_components", "cb_host_alloc", "cb_host_free",
           "cb_device_alloc", "cb_device_free", "cb_memcpy_h2d", "cb_memcpy_d2h", "cb_device_count", "cb_version",
           "cb_pool_create", "cb_pool_destroy", "cb_pool_last_error", "cb_pool_size", "cb_pool_context", "cb_pool_set_family_tag36h11",
           "cb_pool_detect_gray", "cb_pool_get_timing"]

_lib = None


class ChalkydriError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"chalkydri_b200 error {code}: {text}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
        L.cb_create.restype = vp; L.cb_create.argtypes = [i32] * 5
        L.cb_destroy.restype = None; L.cb_destroy.argtypes = [vp]
        L.cb_last_error.restype = C.c_char_p; L.cb_last_error.argtypes = [vp]
        L.cb_set_family_tag36h11.argtypes = [vp, i32]
        L.cb_set_params.argtypes = [vp, f32, f32, i32, f64, i32, i32, f32, f32, i32]
        L.cb_detect_gray.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp]
        L.cb_detect_gray_device.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp]
        L.cb_detect_gray_submit.argtypes = [vp, vp, i32, i32, i32, sz, i32]
        L.cb_detect_gray_collect.argtypes = [vp, vp, vp]
        L.cb_detect_gray_pending.argtypes = [vp]
        L.cb_detect_rgb.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.cb_detect_yuyv.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.cb_detect_yuv420.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.cb_rgb_to_gray.argtypes = [vp, vp, i32, i32, i32, vp]
        L.cb_yuyv_to_gray.argtypes = [vp, vp, i32, i32, i32, vp]
        L.cb_decimated_size.argtypes = [vp, i32, i32, vp, vp]
        L.cb_threshold.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp]
        L.cb_labels.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp]
        L.cb_quads.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, i32, vp, vp]
        L.cb_get_timing.argtypes = [vp, vp]
        L.cb_sqpnp_set.argtypes = [vp, i32, f64]
        L.cb_sqpnp_batch.argtypes = [vp, vp, vp, vp, i32, vp, vp, f64, i64, vp, vp]
        L.cb_sqpnp_batch_device.argtypes = [vp, vp, vp, vp, i32, vp, vp, f64, i64, vp, vp]
        L.cb_create_solver_camera_transform.argtypes = [f64] * 6 + [vp]
        L.cb_unproject_opencv5.argtypes = [vp, vp, vp, i64, vp, vp]
        L.cb_pack_vision_measurements.argtypes = [vp, vp, vp, vp, C.c_uint8, i32, vp]
        L.cb_set_field.argtypes = [vp, vp, vp, i32]
        L.cb_set_camera.argtypes = [vp, vp, vp]
        L.cb_detect_pose_gray.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, f64, vp, vp, vp, vp, vp]
        L.cb_detect_pose_gray_submit.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, f64]
        L.cb_detect_pose_gray_collect.argtypes = [vp, vp, vp, vp, vp, vp]
        L.cb_cat_calc_otsu.argtypes = [vp, vp, i32, i32, vp]
        L.cb_cat_thresh.argtypes = [vp, vp, i32, i32, vp]
        L.cb_cat_detect_corners.argtypes = [vp, vp, i32, i32, vp, i64, vp]
        L.cb_cat_check_edges.argtypes = [vp, vp, i32, i32, vp, i64, vp, i64, vp]
        L.cb_cat_process_frame.argtypes = [vp, vp, i32, i32, vp, vp, i64, vp, vp, i64, vp]
        L.cb_cat_detect_tags.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.cb_clusters.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp, i64, vp, vp]
        L.cb_frame_flags.argtypes = [vp, vp, i32]
        L.cb_cat_connected_components.argtypes = [vp, vp, i32, i32, vp, vp]
        L.cb_host_alloc.restype = vp; L.cb_host_alloc.argtypes = [sz]
        L.cb_host_free.restype = None; L.cb_host_free.argtypes = [vp]
        L.cb_device_alloc.restype = vp; L.cb_device_alloc.argtypes = [vp, sz]
        L.cb_device_free.restype = None; L.cb_device_free.argtypes = [vp, vp]
        L.cb_memcpy_h2d.argtypes = [vp, vp, vp, sz]
        L.cb_memcpy_d2h.argtypes = [vp, vp, vp, sz]
        L.cb_pool_create.restype = vp; L.cb_pool_create.argtypes = [vp, i32, i32, i32, i32, i32]
        L.cb_pool_destroy.restype = None; L.cb_pool_destroy.argtypes = [vp]
        L.cb_pool_last_error.restype = C.c_char_p; L.cb_pool_last_error.argtypes = [vp]
        L.cb_pool_size.argtypes = [vp]
        L.cb_pool_context.restype = vp; L.cb_pool_context.argtypes = [vp, i32]
        L.cb_pool_set_family_tag36h11.argtypes = [vp, i32]
        L.cb_pool_detect_gray.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp]
        L.cb_pool_get_timing.argtypes = [vp, vp]
        L.cb_device_count.restype = i32
        L.cb_version.restype = C.c_char_p
        _lib = L
    return _lib


def ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def pinned_array(shape, dtype):
    """numpy array backed by page-locked host memory from cb_host_alloc (kept alive by the returned array)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().cb_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("cb_host_alloc failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = p
    return arr


_PINNED: dict = {}


def free_pinned(arr):
    p = _PINNED.pop(arr.__array_interface__["data"][0], None)
    if p:
        lib().cb_host_free(p)
