"""Single-frame latency breakdown (c2 frame) through the host-buffer entry point."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder
frames, _ = synth.render_batch(1456, 1088, 4, 8, seed=0x5EED + 2, edge_px=(40.0, 200.0))
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(1456, 1088, 1, 64).build()
pin = capi.pinned_array(frames[:1].shape, np.uint8); pin[:] = frames[:1]
for _ in range(20): det.detect_batch(pin)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); out, counts = det.detect_batch(pin); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("wall ms p50 %.3f p90 %.3f min %.3f" % (np.percentile(ts, 50), np.percentile(ts, 90), ts.min()))
print(det.timing())
det.close()
