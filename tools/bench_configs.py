"""Throughput of the other BASELINE.json detector configs on one GPU (bench.py measures configs[1]):
c1 1280x720 / 4 tags, c3 4608x2592 / 40 tags incl. 10 small ones, c4 1280x800 / 6 tags (the per-GPU share of the 4096-frame stream).
Prints one JSON line per config: device-resident and end-to-end frames/s, single-frame p50 latency, detections vs ground truth."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder, DET_DTYPE

CONFIGS = [
    ("c1", 1280, 720, 4, 256, 8, dict(edge_px=(60.0, 150.0))),
    ("c3", 4608, 2592, 40, 32, 4, dict(edge_px=(40.0, 300.0), small_tags=10)),
    ("c4", 1280, 800, 6, 512, 8, dict(edge_px=(40.0, 160.0))),
]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
L = capi.lib()
for name, W, H, tags, B, unique, kw in CONFIGS:
    frames, truths = synth.render_batch(W, H, B, tags, seed=0x5EED + int(name[1]), unique=unique, **kw)
    det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 64).build()
    h = capi.pinned_array(frames.shape, np.uint8); h[...] = frames
    out = capi.pinned_array((B, 64), DET_DTYPE); counts = capi.pinned_array((B,), np.int32)
    d = L.cb_device_alloc(det.ctx, frames.nbytes)
    assert d and L.cb_memcpy_h2d(det.ctx, d, capi.ptr(h), frames.nbytes) == 0
    for _ in range(3):
        det.detect_batch_device(d, B, H, W, out=out, counts=counts)
    ms = 0.0
    for _ in range(steps):
        det.detect_batch_device(d, B, H, W, out=out, counts=counts)
        ms += det.timing()["total_ms"]
    ndet = int(counts.sum())
    for _ in range(2):
        det.detect_batch(h, out=out, counts=counts)
    t0 = time.perf_counter()
    for _ in range(steps):
        det.detect_batch(h, out=out, counts=counts)
    wall = time.perf_counter() - t0
    lat = []
    for i in range(30):
        t0 = time.perf_counter(); det.detect_batch(h[:1], out=out[:1], counts=counts[:1]); lat.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"config": name, "workload": f"{B} x {W}x{H}, {tags} tags", "value": B * steps / (ms / 1e3), "e2e": B * steps / wall,
                      "unit": "frames/s", "p50_frame_latency_ms": float(np.median(lat[5:])), "detections_per_step": ndet,
                      "ground_truth_tags_per_step": int(sum(len(t["ids"]) for t in truths)), "h2d_bytes_per_step": int(frames.nbytes)}))
    L.cb_device_free(det.ctx, d)
    det.close()
