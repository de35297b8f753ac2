"""One pass of both pre-processing kernels on a 64-frame batch of 1456x1088 frames (for ncu captures of rgb_to_gray / yuyv_to_gray)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200.detector import DetectorBuilder
W, H, B = 1456, 1088, 64
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 16).build()
rng = np.random.default_rng(0)
rgb = np.repeat(rng.integers(100, 140, (4, H, W, 3), dtype=np.uint8), B // 4, axis=0)
yuyv = np.repeat(rng.integers(100, 140, (4, H, W * 2), dtype=np.uint8), B // 4, axis=0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    det.detect_rgb_batch(rgb)
    det.detect_yuyv_batch(yuyv)
print("ok", det.timing()["preprocess_ms"])
det.close()
