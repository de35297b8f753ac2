"""Experiment: do two independent pipelines (two contexts / streams) on one GPU beat one pipeline over the whole batch?"""
import sys, os, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder

B = 256
nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames, _ = synth.render_batch(1456, 1088, B, 8, seed=0x5EED + 2, unique=16, edge_px=(40.0, 200.0))
L = capi.lib()
per = B // nctx
dets = [DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(1456, 1088, per, 64).build() for _ in range(nctx)]
dptr = []
for i, d in enumerate(dets):
    p = L.cb_device_alloc(d.ctx, frames[i * per:(i + 1) * per].nbytes)
    L.cb_memcpy_h2d(d.ctx, p, capi.ptr(np.ascontiguousarray(frames[i * per:(i + 1) * per])), frames[i * per:(i + 1) * per].nbytes)
    dptr.append(p)

def run(i, reps, out):
    for _ in range(reps):
        o, c = dets[i].detect_batch_device(dptr[i], per, 1088, 1456)
    out[i] = int(c.sum())

for reps in (2, 5):
    res = [0] * nctx
    ts = [threading.Thread(target=run, args=(i, reps, res)) for i in range(nctx)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.perf_counter() - t0
    print(f"nctx={nctx} reps={reps}: {dt / reps * 1e3:.2f} ms per {B} frames -> {B * reps / dt:.0f} frames/s, dets {sum(res)}")
