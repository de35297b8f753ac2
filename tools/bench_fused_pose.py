"""End-to-end frames/s of the fused call cb_detect_pose_gray (host frames in -> detection lists + robot pose per frame out)
on the c2 workload, next to cb_detect_gray alone."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi, field
from chalkydri_b200.capi import DET_DTYPE, ISO_DTYPE, POSE_DTYPE
from chalkydri_b200.detector import DetectorBuilder
from chalkydri_b200.solver import SqPnP, SIGN_FLIP_CONST

W, H, B = 1456, 1088, 256
frames, _ = synth.render_batch(W, H, B, 8, seed=0x5EED + 2, unique=16, edge_px=(40.0, 200.0))
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 64).build()
L = capi.lib()
tags = field.load()
ids = np.array(sorted(tags), np.int32)
poses_f = np.array([tags[int(i)] for i in ids], ISO_DTYPE)
assert L.cb_set_field(det.ctx, capi.ptr(ids), capi.ptr(poses_f), len(ids)) == 0
r2c = np.ascontiguousarray(np.array(SqPnP.create_solver_camera_transform(0.2, -0.1, 0.5, 0.0, -10.0, 15.0), ISO_DTYPE))
cam = np.array(synth.scaled_calib(W, H), np.float64)
assert L.cb_set_camera(det.ctx, capi.ptr(cam), capi.ptr(r2c)) == 0
h = capi.pinned_array(frames.shape, np.uint8); h[...] = frames
out = capi.pinned_array((B, 64), DET_DTYPE); counts = capi.pinned_array((B,), np.int32)
poses = capi.pinned_array((B,), POSE_DTYPE); ok = capi.pinned_array((B,), np.uint8); nt = capi.pinned_array((B,), np.int32)
gyro = np.full(B, 0.3)
def fused():
    rc = L.cb_detect_pose_gray(det.ctx, capi.ptr(h), W, H, W, W * H, B, capi.ptr(gyro), SIGN_FLIP_CONST, capi.ptr(out), capi.ptr(counts),
                               capi.ptr(poses), capi.ptr(ok), capi.ptr(nt))
    assert rc == 0, L.cb_last_error(det.ctx)
def plain():
    det.detect_batch(h, out=out, counts=counts)
def stream(n, pose):
    """submit batch k+1, then collect batch k (two in flight); every copy inside the caller's timed region"""
    def sub():
        if pose:
            rc = L.cb_detect_pose_gray_submit(det.ctx, capi.ptr(h), W, H, W, W * H, B, capi.ptr(gyro), SIGN_FLIP_CONST)
        else:
            rc = L.cb_detect_gray_submit(det.ctx, capi.ptr(h), W, H, W, W * H, B)
        assert rc == 0, L.cb_last_error(det.ctx)
    sub()
    for k in range(n):
        if k + 1 < n:
            sub()
        if pose:
            rc = L.cb_detect_pose_gray_collect(det.ctx, capi.ptr(out), capi.ptr(counts), capi.ptr(poses), capi.ptr(ok), capi.ptr(nt))
        else:
            rc = L.cb_detect_gray_collect(det.ctx, capi.ptr(out), capi.ptr(counts))
        assert rc == 0, L.cb_last_error(det.ctx)
res = {}
N = 10
for name, fn in (("detect_only", plain), ("detect_pose_fused", fused)):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(N): fn()
    res[name] = B * N / (time.perf_counter() - t0)
fused()
blocking_poses, blocking_ok = poses.copy(), ok.copy()
for name, pose in (("detect_only_streaming", False), ("detect_pose_fused_streaming", True)):
    stream(3, pose)
    t0 = time.perf_counter()
    stream(N, pose)
    res[name] = B * N / (time.perf_counter() - t0)
res["streaming_poses_identical_to_blocking_call"] = bool((ok == blocking_ok).all() and poses[ok > 0].tobytes() == blocking_poses[blocking_ok > 0].tobytes())
res.update(workload="c2: 256 x 1456x1088, 8 tags", unit="frames/s e2e (pinned host frames in, lists [+ poses] out)",
           poses_solved_per_step=int(ok.sum()), tags_used_per_step=int(nt.sum()), detections_per_step=int(counts.sum()))
print(json.dumps(res))
