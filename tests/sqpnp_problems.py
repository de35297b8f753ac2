"""Seeded SQPnP problems: tag poses from field.json, random robot pose in front of the tag(s),
exact projection + corner noise, gyro = true yaw + N(0, 2 deg)."""
import numpy as np

from chalkydri_b200 import field
from chalkydri_b200.capi import ISO_DTYPE

S = 0.1651 / 2
CORNERS = np.array([[0, -S, -S], [0, S, -S], [0, S, S], [0, -S, S]])


def qmat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = w * w + x * x - y * y - z * z; m[..., 0, 1] = 2 * (x * y - w * z); m[..., 0, 2] = 2 * (w * y + x * z)
    m[..., 1, 0] = 2 * (w * z + x * y); m[..., 1, 1] = w * w - x * x + y * y - z * z; m[..., 1, 2] = 2 * (y * z - w * x)
    m[..., 2, 0] = 2 * (x * z - w * y); m[..., 2, 1] = 2 * (w * x + y * z); m[..., 2, 2] = w * w - x * x - y * y + z * z
    return m


def camera_iso():
    """robot_to_cam as create_solver_camera_transform(0.2, 0.1, 0.5, 0, -10, 15) would give (fixed literal, no product code)."""
    from chalkydri_b200.solver import SqPnP
    return SqPnP.create_solver_camera_transform(0.2, 0.1, 0.5, 0.0, -10.0, 15.0)


def make_problems(n, seed, two_tag_frac=0.1, noise_px=0.25, max_tags=2, fast=False):
    rng = np.random.default_rng(seed)
    layout = field.load()
    ids = sorted(layout)
    T = np.array([layout[i]["t"] for i in ids])
    Q = np.array([layout[i]["q"] for i in ids])
    Rm = qmat(Q)
    r2c = camera_iso()
    R_r2c, t_r2c = qmat(r2c["q"]), r2c["t"]
    k = rng.integers(0, len(ids), n)
    two = rng.random(n) < two_tag_frac
    # second tag: nearest other tag facing the same way
    nearest = np.empty(len(ids), int)
    for i in range(len(ids)):
        d = np.linalg.norm(T - T[i], axis=1) + 1e3 * (np.einsum("ij,j->i", Rm[:, :, 0], Rm[i, :, 0]) < 0.9) + 1e6 * (np.arange(len(ids)) == i)
        nearest[i] = int(np.argmin(d))
    k2 = nearest[k]
    normal = Rm[k][:, :, 0]
    side = np.cross(np.array([0, 0, 1.0]), normal)
    d = rng.uniform(0.5, 8.0, n)
    lat = rng.uniform(-0.4, 0.4, n) * d
    off = normal * d[:, None] + side * lat[:, None]
    pos = T[k] + off
    pos[:, 2] = 0.0
    yaw = np.arctan2(-off[:, 1], -off[:, 0]) + rng.uniform(-0.3, 0.3, n)
    cy, sy = np.cos(yaw), np.sin(yaw)
    Rr = np.zeros((n, 3, 3)); Rr[:, 0, 0] = cy; Rr[:, 0, 1] = -sy; Rr[:, 1, 0] = sy; Rr[:, 1, 1] = cy; Rr[:, 2, 2] = 1
    tags = np.zeros((n, max_tags), ISO_DTYPE)
    bearings = np.zeros((n, max_tags * 4, 3))
    n_tags = np.where(two, 2, 1).astype(np.int32)
    f = 900.0
    for slot, kk in enumerate((k, k2)):
        tags["t"][:, slot] = T[kk]
        tags["q"][:, slot] = Q[kk]
        pw = np.einsum("nij,cj->nci", Rm[kk], CORNERS) + T[kk][:, None, :]
        pr = np.einsum("nji,ncj->nci", Rr, pw - pos[:, None, :])          # world -> robot
        pc = np.einsum("ij,ncj->nci", R_r2c, pr) + t_r2c                   # robot -> camera
        z = pc[:, :, 2:3]
        b = pc / np.where(np.abs(z) < 1e-9, 1e-9, z)
        b[:, :, :2] += rng.normal(0, noise_px / f, (n, 4, 2)) if noise_px > 0 else 0.0
        b[:, :, 2] = 1.0
        bad = (z[:, :, 0] <= 0.05).any(1)
        if slot == 0:
            b[bad] = np.array([0.0, 0.0, -1.0])     # degenerate: every point behind the camera -> solver returns None
        else:
            n_tags[bad & two] = 1
        bearings[:, slot * 4:(slot + 1) * 4] = b
    gyro = yaw + rng.normal(0, np.deg2rad(2.0), n) * (0 if noise_px == 0 else 1)
    return tags, bearings, n_tags, r2c, gyro, {"pos": pos, "yaw": yaw}
