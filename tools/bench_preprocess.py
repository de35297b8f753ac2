"""Pre-processing kernels (row P1 + camera formats) against the HBM roofline: packed RGB -> gray (utils.rs:43) and YUYV -> gray
on a 256-frame batch of 1456x1088 frames.  Algorithmic bytes per frame: RGB 4*W*H (3 read + 1 written), YUYV 3*W*H.
Times are the library's CUDA events around the conversion launch (cb_timing.preprocess_ms), frames already in HBM."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import capi
from chalkydri_b200.detector import DetectorBuilder

W, H, B = 1456, 1088, 256
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
peak = 6533.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 16).build()
rng = np.random.default_rng(0)
res = {"workload": f"{B} x {W}x{H}", "peak_gbs": peak}
for name, bpp, fn in (("rgb_to_gray_kernel", 3, det.rgb_to_gray), ("yuyv_to_gray_kernel", 2, det.yuyv_to_gray)):
    shape = (B, H, W, 3) if bpp == 3 else (B, H, W * 2)
    h = capi.pinned_array(shape, np.uint8)
    h[:8] = rng.integers(0, 256, (8,) + shape[1:], dtype=np.uint8)
    for b in range(8, B):
        h[b] = h[b % 8]
    ms = []
    for i in range(reps + 2):
        fn(h)                                                     # the stage tap: H2D, ONE conversion launch, D2H
        if i >= 2:
            ms.append(det.timing()["preprocess_ms"])
    t = float(np.median(ms))
    algo = (bpp + 1) * W * H * B
    res[name] = {"ms_per_launch": t, "algorithmic_bytes_per_launch": algo, "achieved_gbs": algo / (t * 1e-3) / 1e9, "frac": algo / (t * 1e-3) / 1e9 / peak}
    capi.free_pinned(h)
print(json.dumps(res))
det.close()
