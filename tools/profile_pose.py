import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np
from chalkydri_b200 import synth, capi
from chalkydri_b200.pipeline import AprilTags
W, H = 1280, 720
frames, _ = synth.render_batch(W, H, 1, 4, seed=0x5EED + 1, edge_px=(60.0, 150.0))
class Comm:
    def gyro_angle(self): return 0.1
    def publish(self, *a): pass
calib = synth.scaled_calib(W, H)
keys = ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")
config = {"family": "tag36h11", "bits_corrected": 3, "cam_id": 7, "robot_to_cam": json.dumps({"x": 0.2, "y": 0.1, "z": 0.5, "roll": 0.0, "pitch": -10.0, "yaw": 15.0}), "calib": json.dumps({"OpenCVModel5": dict(zip(keys, calib))})}
task = AprilTags.new(config, Comm(), max_width=W, max_height=H, max_batch=1)
pin = capi.pinned_array((1, H, W), np.uint8); pin[0] = frames[0]
for _ in range(3):
    r = task.process_batch(1_000_000, [999_000], pin)
print(r, task.last_batch[4])
