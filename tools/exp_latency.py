"""Single-frame latency breakdown (c2 or c1 frame) through the host-buffer entry point."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
W, H, TAGS, SEED, EDGE = (1456, 1088, 8, 0x5EED + 2, (40.0, 200.0)) if wl == "c2" else (1280, 720, 4, 0x5EED + 1, (60.0, 150.0))
frames, _ = synth.render_batch(W, H, 4, TAGS, seed=SEED, edge_px=EDGE)
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, 1, 64).build()
pin = capi.pinned_array(frames[:1].shape, np.uint8); pin[:] = frames[:1]
for _ in range(20): det.detect_batch(pin)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); out, counts = det.detect_batch(pin); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("wall ms p50 %.3f p90 %.3f min %.3f" % (np.percentile(ts, 50), np.percentile(ts, 90), ts.min()))
print(det.timing())
det.close()
