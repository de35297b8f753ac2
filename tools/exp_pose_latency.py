"""Single-frame latency of the task-shaped calls: detection alone, the fused detect -> pose call, and one SQPnP problem."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi, field
from chalkydri_b200.pipeline import AprilTags
from chalkydri_b200.solver import SqPnP
from tests.sqpnp_problems import make_problems

W, H = 1280, 720
frames, _ = synth.render_batch(W, H, 2, 4, seed=0x5EED + 1, edge_px=(60.0, 150.0))


class Comm:
    def gyro_angle(self): return 0.1
    def publish(self, *a): pass


def p50(fn, n=200, warm=20):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.percentile(np.array(ts) * 1e3, 50))


import json
calib = synth.scaled_calib(W, H)
keys = ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")
config = {"family": "tag36h11", "bits_corrected": 3, "cam_id": 7,
          "robot_to_cam": json.dumps({"x": 0.2, "y": 0.1, "z": 0.5, "roll": 0.0, "pitch": -10.0, "yaw": 15.0}),
          "calib": json.dumps({"OpenCVModel5": dict(zip(keys, calib))})}
task = AprilTags.new(config, Comm(), max_width=W, max_height=H, max_batch=1)
if task is not None:
    pin = capi.pinned_array((1, H, W), np.uint8); pin[0] = frames[0]
    print("process_batch (fused detect->pose, B=1) p50 ms", round(p50(lambda: task.process_batch(1_000_000, [999_000], pin)), 3))
    print("detect alone p50 ms", round(p50(lambda: task.detector.detect_batch(pin)), 3))
tags, bearings, n_tags, r2c, gyro, _ = make_problems(4, 1, 0.5, 0.25)
s = SqPnP.new()
print("one SQPnP problem through cb_sqpnp_batch p50 ms", round(p50(lambda: s.solve_robot_pose_batch(tags[:1], bearings[:1], n_tags[:1], r2c, gyro[:1], 600.0)), 3))

