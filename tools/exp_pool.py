"""The c4 stream through the single-process pool alone (no other rank on the box): python tools/exp_pool.py [ngpu]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import capi
from chalkydri_b200.detector import DET_DTYPE
from chalkydri_b200.pool import DetectorPool
from tools.bench_c4_stream import unique_frames, W, H, BATCH, CAP, UNIQUE, TOTAL
ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2
uniq, ntags = unique_frames()
hp = capi.pinned_array((TOTAL, H, W), np.uint8)
for i in range(TOTAL):
    hp[i] = uniq[i % UNIQUE]
out = np.zeros((TOTAL, CAP), DET_DTYPE); counts = np.zeros(TOTAL, np.int32)
pool = DetectorPool(list(range(ngpu)), W, H, BATCH, CAP)
pool.detect_batch(hp, out=out, counts=counts)
for _ in range(3):
    t0 = time.perf_counter(); pool.detect_batch(hp, out=out, counts=counts); dt = time.perf_counter() - t0
    print(ngpu, "GPUs:", round(TOTAL / dt), "frames/s", pool.timing())
pool.close()
