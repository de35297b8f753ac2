"""Small fixed workload for ncu captures: B frames of the c2 configuration through the device-resident entry point."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frames, _ = synth.render_batch(1456, 1088, B, 8, seed=0x5EED + 2, unique=min(B, 8), edge_px=(40.0, 200.0))
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(1456, 1088, B, 64).build()
L = capi.lib()
d = L.cb_device_alloc(det.ctx, frames.nbytes)
L.cb_memcpy_h2d(det.ctx, d, capi.ptr(frames), frames.nbytes)
for _ in range(reps):
    out, counts = det.detect_batch_device(d, B, 1088, 1456)
    print(det.timing(), int(counts.sum()))
L.cb_device_free(det.ctx, d)
det.close()
