"""Measure pinned host->device copy bandwidth on this box (bounds the end-to-end frames/s of the HOST-buffer entry point)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import capi
from chalkydri_b200.detector import DetectorBuilder
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(640, 480, 2, 16).build()
L = capi.lib()
for mb in (25, 100, 405):
    n = mb << 20
    h = capi.pinned_array((n,), np.uint8)
    h[:] = 1
    d = L.cb_device_alloc(det.ctx, n)
    L.cb_memcpy_h2d(det.ctx, d, capi.ptr(h), n)
    t0 = time.perf_counter()
    for _ in range(5):
        L.cb_memcpy_h2d(det.ctx, d, capi.ptr(h), n)
    dt = (time.perf_counter() - t0) / 5
    print(f"{mb} MiB pinned H2D: {dt*1e3:.2f} ms  {n/dt/1e9:.1f} GB/s")
    L.cb_device_free(det.ctx, d)
det.close()
