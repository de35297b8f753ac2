"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg,
never by chalkydri_b200/.  See oracle/oracle.h for what each entry point restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Params(C.Structure):
    _fields_ = [("quad_decimate", C.c_float), ("refine_edges", C.c_int), ("decode_sharpening", C.c_double),
                ("min_cluster_pixels", C.c_int), ("max_nmaxima", C.c_int), ("critical_rad", C.c_float),
                ("max_line_fit_mse", C.c_float), ("min_white_black_diff", C.c_int), ("bits_corrected", C.c_int)]


class Detection(C.Structure):
    _fields_ = [("id", C.c_int32), ("hamming", C.c_int32), ("decision_margin", C.c_float), ("pad_", C.c_float),
                ("H", C.c_double * 9), ("c", C.c_double * 2), ("p", (C.c_double * 2) * 4)]


DET_DTYPE = np.dtype([("id", "<i4"), ("hamming", "<i4"), ("decision_margin", "<f4"), ("pad_", "<f4"),
                      ("H", "<f8", (9,)), ("c", "<f8", (2,)), ("p", "<f8", (4, 2))])
assert DET_DTYPE.itemsize == C.sizeof(Detection) == 168

QUAD_DTYPE = np.dtype([("p", "<f4", (4, 2)), ("reversed_border", "<i4"), ("npoints", "<i4"), ("cluster_id", "<u8")])


class Quad(C.Structure):
    _fields_ = [("p", (C.c_float * 2) * 4), ("reversed_border", C.c_int32), ("npoints", C.c_int32), ("cluster_id", C.c_uint64)]


assert QUAD_DTYPE.itemsize == C.sizeof(Quad)


class Taps(C.Structure):
    _fields_ = [("thresh", C.c_void_p), ("labels", C.c_void_p), ("comp_size", C.c_void_p), ("quads", C.c_void_p),
                ("quads_cap", C.c_int32), ("nquads", C.c_int32), ("nclusters", C.c_int32), ("npoints", C.c_int64),
                ("w", C.c_int32), ("h", C.c_int32), ("pts", C.c_void_p), ("pts_cluster", C.c_void_p), ("pts_cap", C.c_int64)]


class Iso3(C.Structure):
    _fields_ = [("t", C.c_double * 3), ("q", C.c_double * 4)]


ISO_DTYPE = np.dtype([("t", "<f8", (3,)), ("q", "<f8", (4,))])
POSE_DTYPE = np.dtype([("rot", "<f8", (9,)), ("pos", "<f8", (3,)), ("std_devs", "<f8", (3,))])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("apriltag_oracle.cpp", "sqpnp_oracle.cpp", "cat_oracle.cpp", "oracle.h",
                                             "tag36h11_codes.inc", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_tag36h11_codes.restype = C.POINTER(C.c_uint64)
        L.orc_detect.restype = C.c_int
        L.orc_detect_batch.restype = C.c_int
        L.orc_sqpnp_solve_robot_pose.restype = C.c_int
        L.orc_cat_grayscale.restype = C.c_uint8
        L.orc_cat_detect_corners.restype = C.c_int64
        L.orc_cat_check_edges.restype = C.c_int64
        L.orc_unproject_opencv5.restype = C.c_int
    return _LIB


def default_params(**over) -> Params:
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def tag36h11_codes() -> np.ndarray:
    n = C.c_int()
    ptr = lib().orc_tag36h11_codes(C.byref(n))
    return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def decimated_size(W, H, f=2.0):
    w, h = C.c_int(), C.c_int()
    lib().orc_decimated_size(W, H, C.c_float(f), C.byref(w), C.byref(h))
    return w.value, h.value


def threshold(gray: np.ndarray, prm: Params | None = None) -> np.ndarray:
    prm = prm or default_params()
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    w, h = decimated_size(W, H, prm.quad_decimate)
    out = np.empty((h, w), np.uint8)
    lib().orc_threshold(_ptr(gray), W, H, W, C.byref(prm), _ptr(out))
    return out


def detect(gray: np.ndarray, prm: Params | None = None, cap: int = 256, taps: bool = False, pts_cap: int = 0):
    """Returns detections (structured array); with taps=True also a dict of intermediates."""
    prm = prm or default_params()
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros(cap, DET_DTYPE)
    if not taps:
        n = lib().orc_detect(_ptr(gray), W, H, W, C.byref(prm), _ptr(out), cap, None)
        if n < 0:
            raise RuntimeError(f"orc_detect error {n}")
        return out[:n]
    w, h = decimated_size(W, H, prm.quad_decimate)
    t = Taps()
    thr = np.empty((h, w), np.uint8)
    labels = np.empty((h, w), np.uint32)
    csize = np.empty((h, w), np.uint32)
    quads = np.zeros(4096, QUAD_DTYPE)
    t.thresh, t.labels, t.comp_size, t.quads, t.quads_cap = _ptr(thr), _ptr(labels), _ptr(csize), _ptr(quads), 4096
    pts = pcl = None
    if pts_cap:
        pts = np.zeros((pts_cap, 4), np.int16)
        pcl = np.zeros(pts_cap, np.uint64)
        t.pts, t.pts_cluster, t.pts_cap = _ptr(pts), _ptr(pcl), pts_cap
    n = lib().orc_detect(_ptr(gray), W, H, W, C.byref(prm), _ptr(out), cap, C.byref(t))
    if n < 0:
        raise RuntimeError(f"orc_detect error {n}")
    d = {"thresh": thr, "labels": labels, "comp_size": csize, "quads": quads[:min(t.nquads, 4096)], "nquads": t.nquads,
         "nclusters": t.nclusters, "npoints": t.npoints, "w": t.w, "h": t.h}
    if pts_cap:
        k = min(int(t.npoints), pts_cap)
        d["pts"], d["pts_cluster"] = pts[:k], pcl[:k]
    return out[:n], d


def detect_with_map(gray: np.ndarray, tmap: np.ndarray, prm: Params | None = None, cap: int = 256):
    """Upstream's pipeline from connected_components() on, fed with an external ternary map (0 / 127 / 255) of the decimated frame."""
    prm = prm or default_params()
    gray = np.ascontiguousarray(gray, np.uint8)
    tmap = np.ascontiguousarray(tmap, np.uint8)
    H, W = gray.shape
    w, h = decimated_size(W, H, prm.quad_decimate)
    assert tmap.shape == (h, w)
    out = np.zeros(cap, DET_DTYPE)
    L = lib()
    L.orc_detect_with_map.restype = C.c_int
    n = L.orc_detect_with_map(_ptr(gray), W, H, W, _ptr(tmap), C.byref(prm), _ptr(out), cap)
    if n < 0:
        raise RuntimeError(f"orc_detect_with_map error {n}")
    return out[:n]


def detect_batch(frames: np.ndarray, prm: Params | None = None, cap: int = 64, nthreads: int = 1):
    prm = prm or default_params()
    frames = np.ascontiguousarray(frames, np.uint8)
    B, H, W = frames.shape
    out = np.zeros((B, cap), DET_DTYPE)
    counts = np.zeros(B, np.int32)
    rc = lib().orc_detect_batch(_ptr(frames), W, H, W, C.c_int64(H * W), B, C.byref(prm), _ptr(out), cap, _ptr(counts), nthreads)
    if rc < 0:
        raise RuntimeError(f"orc_detect_batch error {rc}")
    return out, counts


# ---------------- SQPnP ----------------
def sqpnp_solve_robot_pose(tags: np.ndarray, bearings: np.ndarray, robot_to_cam: np.ndarray, gyro: float,
                           sign_change_error: float = 600.0, max_iter: int = 15, tol_sq: float = 1e-16):
    tags = np.ascontiguousarray(tags, ISO_DTYPE)
    bearings = np.ascontiguousarray(bearings, np.float64).reshape(-1, 3)
    r2c = np.ascontiguousarray(robot_to_cam, ISO_DTYPE).reshape(1)
    out = np.zeros(1, POSE_DTYPE)
    ok = lib().orc_sqpnp_solve_robot_pose(_ptr(tags), len(tags), _ptr(bearings), len(bearings), _ptr(r2c), C.c_double(gyro),
                                          C.c_double(sign_change_error), max_iter, C.c_double(tol_sq), _ptr(out))
    return (out[0] if ok else None)


def sqpnp_batch(tags, bearings, n_tags, robot_to_cam, gyro, sign_change_error=600.0, nthreads=1):
    tags = np.ascontiguousarray(tags, ISO_DTYPE)
    n, max_tags = tags.shape
    bearings = np.ascontiguousarray(bearings, np.float64).reshape(n, max_tags * 4, 3)
    n_tags = np.ascontiguousarray(n_tags, np.int32)
    gyro = np.ascontiguousarray(gyro, np.float64)
    r2c = np.ascontiguousarray(robot_to_cam, ISO_DTYPE).reshape(1)
    out = np.zeros(n, POSE_DTYPE)
    ok = np.zeros(n, np.uint8)
    lib().orc_sqpnp_batch(_ptr(tags), _ptr(bearings), _ptr(n_tags), max_tags, _ptr(r2c), _ptr(gyro),
                          C.c_double(sign_change_error), C.c_int64(n), _ptr(out), _ptr(ok), nthreads)
    return out, ok


def create_solver_camera_transform(fwd, left, up, roll_deg, pitch_deg, yaw_deg):
    out = np.zeros(1, ISO_DTYPE)
    lib().orc_create_solver_camera_transform(C.c_double(fwd), C.c_double(left), C.c_double(up), C.c_double(roll_deg),
                                             C.c_double(pitch_deg), C.c_double(yaw_deg), _ptr(out))
    return out[0]


def sqpnp_omega(pts3d, bearings):
    pts3d = np.ascontiguousarray(pts3d, np.float64)
    bearings = np.ascontiguousarray(bearings, np.float64)
    om, qi, qrt = np.zeros(81), np.zeros(9), np.zeros(27)
    lib().orc_sqpnp_omega(_ptr(pts3d), _ptr(bearings), len(pts3d), _ptr(om), _ptr(qi), _ptr(qrt))
    return om.reshape(9, 9).T, qi.reshape(3, 3).T, qrt.reshape(3, 9).T


def sym_eigen9(a):
    a = np.ascontiguousarray(np.asarray(a, np.float64).T)
    d, v = np.zeros(9), np.zeros(81)
    lib().orc_sym_eigen9(_ptr(a), _ptr(d), _ptr(v))
    return d, v.reshape(9, 9).T


def nearest_so3(m):
    a = np.ascontiguousarray(np.asarray(m, np.float64).T).reshape(9)
    out = np.zeros(9)
    lib().orc_nearest_so3(_ptr(a), _ptr(out))
    return out.reshape(3, 3).T


def unproject_opencv5(params9, u, v):
    p = np.ascontiguousarray(params9, np.float64)
    out = np.zeros(3)
    ok = lib().orc_unproject_opencv5(_ptr(p), C.c_double(u), C.c_double(v), _ptr(out))
    return out if ok else None


# ---------------- CAT ----------------
def cat_grayscale(r, g, b):
    return int(lib().orc_cat_grayscale(C.c_uint8(r), C.c_uint8(g), C.c_uint8(b)))


def cat_calc_otsu(rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_cat_calc_otsu(_ptr(rgb), w, h, _ptr(out))
    return out


def cat_thresh(rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_cat_thresh(_ptr(rgb), w, h, _ptr(out))
    return out


def cat_detect_corners(color, cap=1 << 20):
    color = np.ascontiguousarray(color, np.uint8)
    h, w = color.shape
    xy = np.zeros((cap, 2), np.int32)
    n = lib().orc_cat_detect_corners(_ptr(color), w, h, _ptr(xy), C.c_int64(cap))
    return xy[:min(n, cap)], int(n)


def cat_check_edges(color, xy, cap=1 << 20):
    color = np.ascontiguousarray(color, np.uint8)
    xy = np.ascontiguousarray(xy, np.int32)
    h, w = color.shape
    lines = np.zeros((cap, 4), np.int32)
    n = lib().orc_cat_check_edges(_ptr(color), w, h, _ptr(xy), C.c_int64(len(xy)), _ptr(lines), C.c_int64(cap))
    return lines[:min(n, cap)], int(n)


def cat_connected_components(color):
    color = np.ascontiguousarray(color, np.uint8)
    h, w = color.shape
    labels = np.empty((h, w), np.uint32)
    sizes = np.empty((h, w), np.uint32)
    lib().orc_cat_connected_components(_ptr(color), w, h, _ptr(labels), _ptr(sizes))
    return labels, sizes
