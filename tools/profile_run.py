"""Small fixed workload for ncu captures: B frames of the c2 configuration through the device-resident entry point."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"           # c2: 1456x1088, 8 tags; c1: 1280x720, 4 tags (the bench's headline workload)
W, H, TAGS, SEED, EDGE = (1456, 1088, 8, 0x5EED + 2, (40.0, 200.0)) if wl == "c2" else (1280, 720, 4, 0x5EED + 1, (60.0, 150.0))
frames, _ = synth.render_batch(W, H, B, TAGS, seed=SEED, unique=min(B, 8), edge_px=EDGE)
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 64).build()
L = capi.lib()
d = L.cb_device_alloc(det.ctx, frames.nbytes)
L.cb_memcpy_h2d(det.ctx, d, capi.ptr(frames), frames.nbytes)
for _ in range(reps):
    out, counts = det.detect_batch_device(d, B, H, W)
    print(det.timing(), int(counts.sum()))
L.cb_device_free(det.ctx, d)
det.close()
