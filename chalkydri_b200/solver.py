"""Host-side mirror of `chalkydri_sqpnp::SqPnP` (/root/reference/crates/chalkydri_sqpnp/src/lib.rs) above the C ABI.

Same names, argument meaning and failure convention (`None`) as the reference: `SqPnP.new()`, the const builders
`max_iter` / `tolerance` (lib.rs:214-222), `solve_robot_pose` (lib.rs:297-304) and the associated function
`create_solver_camera_transform` (lib.rs:430-461).  `solve_robot_pose_batch` is the B200 addition (one warp per problem).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import ISO_DTYPE, POSE_DTYPE, ChalkydriError

SIGN_FLIP_CONST = 600.0    # crates/apriltags/src/lib.rs:6
MAX_TAGS = 32            # csrc/sqpnp.cuh SQ_MAX_TAGS


def iso(t, q) -> np.ndarray:
    out = np.zeros((), ISO_DTYPE)
    out["t"], out["q"] = t, q
    return out


class SqPnP:
    def __init__(self, device: int = 0, ctx=None):
        self._L = capi.lib()
        self._own = ctx is None
        self._ctx = ctx if ctx is not None else self._L.cb_create(device, 8, 8, 1, 1)
        if not self._ctx:
            raise ChalkydriError(capi.CB_ERR_CUDA, self._L.cb_last_error(None).decode())
        self._max_iter, self._tol = 15, 1e-8      # lib.rs:203-204 (tol_sq = 1e-16)
        self._apply()

    @staticmethod
    def new(device: int = 0):
        return SqPnP(device)

    def _check(self, rc):
        if rc != 0:
            raise ChalkydriError(rc, self._L.cb_last_error(self._ctx).decode())

    def _apply(self):
        self._check(self._L.cb_sqpnp_set(self._ctx, self._max_iter, self._tol))

    def max_iter(self, max_iter: int):
        self._max_iter = int(max_iter)
        self._apply()
        return self

    def tolerance(self, tol: float):
        self._tol = float(tol)
        self._apply()
        return self

    def close(self):
        if self._own and getattr(self, "_ctx", None):
            self._L.cb_destroy(self._ctx)
        self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def create_solver_camera_transform(fwd_m, left_m, up_m, roll_deg, pitch_deg, yaw_deg) -> np.ndarray:
        out = np.zeros(1, ISO_DTYPE)
        rc = capi.lib().cb_create_solver_camera_transform(fwd_m, left_m, up_m, roll_deg, pitch_deg, yaw_deg, capi.ptr(out))
        if rc:
            raise ChalkydriError(rc, "create_solver_camera_transform")
        return out[0]

    def solve_robot_pose(self, points_isometry, points_2d, robot_to_cam, gyro: float, sign_change_error: float):
        """-> (rot 3x3, position [3], std_devs [3]) or None, like Option<(Rot3, Vec3, Vec3)>."""
        tags = np.ascontiguousarray(points_isometry, ISO_DTYPE).reshape(-1)
        p2 = np.ascontiguousarray(points_2d, np.float64).reshape(-1, 3)
        n = len(tags)
        if n > MAX_TAGS:        # the reference has no cap; silently answering None would read as "no pose"
            raise ValueError(f"{n} tags: the solver kernels hold at most {MAX_TAGS} tags per problem")
        if n * 4 < 3 or n * 4 != len(p2):     # lib.rs:255-257
            return None
        out, ok = self.solve_robot_pose_batch(tags[None, :], p2[None, :, :], np.array([n], np.int32), robot_to_cam,
                                              np.array([gyro], np.float64), sign_change_error)
        if not ok[0]:
            return None
        return out[0]["rot"].reshape(3, 3).T.copy(), out[0]["pos"].copy(), out[0]["std_devs"].copy()

    def solve_robot_pose_batch(self, tags, bearings, n_tags, robot_to_cam, gyro, sign_change_error: float = SIGN_FLIP_CONST):
        """tags [N,max_tags] iso, bearings [N,max_tags*4,3], n_tags [N], gyro [N] -> (poses [N], ok [N])."""
        tags = np.ascontiguousarray(tags, ISO_DTYPE)
        N, max_tags = tags.shape
        bearings = np.ascontiguousarray(bearings, np.float64).reshape(N, max_tags * 4, 3)
        n_tags = np.ascontiguousarray(n_tags, np.int32)
        gyro = np.ascontiguousarray(gyro, np.float64)
        r2c = np.ascontiguousarray(robot_to_cam, ISO_DTYPE).reshape(1)
        out = np.zeros(N, POSE_DTYPE)
        ok = np.zeros(N, np.uint8)
        self._check(self._L.cb_sqpnp_batch(self._ctx, capi.ptr(tags), capi.ptr(bearings), capi.ptr(n_tags), max_tags, capi.ptr(r2c),
                                           capi.ptr(gyro), float(sign_change_error), N, capi.ptr(out), capi.ptr(ok)))
        return out, ok

    def timing(self) -> dict:
        t = capi.Timing()
        self._check(self._L.cb_get_timing(self._ctx, C.byref(t)))
        return t.as_dict()

    def unproject(self, params9, pixels):
        """GenericModel::unproject for OpenCVModel5 (crates/apriltags/src/lib.rs:316-321): -> (bearings [n,3], ok [n])."""
        prm = np.ascontiguousarray(params9, np.float64)
        px = np.ascontiguousarray(pixels, np.float64).reshape(-1, 2)
        out = np.zeros((len(px), 3), np.float64)
        ok = np.zeros(len(px), np.uint8)
        self._check(self._L.cb_unproject_opencv5(self._ctx, capi.ptr(prm), capi.ptr(px), len(px), capi.ptr(out), capi.ptr(ok)))
        return out, ok


def euler_angles(rot: np.ndarray):
    """nalgebra Rotation3::euler_angles (roll, pitch, yaw), used at crates/apriltags/src/lib.rs:343."""
    if abs(rot[2, 0]) < 1.0:
        pitch = -np.arcsin(rot[2, 0])
        c = np.cos(pitch)
        return (np.arctan2(rot[2, 1] / c, rot[2, 2] / c), pitch, np.arctan2(rot[1, 0] / c, rot[0, 0] / c))
    if rot[2, 0] <= -1.0:
        return (np.arctan2(rot[0, 1], rot[0, 2]), np.pi / 2, 0.0)
    return (np.arctan2(-rot[0, 1], -rot[0, 2]), -np.pi / 2, 0.0)
