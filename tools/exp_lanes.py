"""Several detector contexts ("lanes") on ONE GPU, each fed by its own host thread: do the kernels of one lane fill the
issue slots the other leaves idle?   python tools/exp_lanes.py [c1|c2] [steps]

Device arm: the 256-frame batch resident in HBM, split evenly over the lanes (wall clock over `steps` rounds, all lanes
joined every round).  Host arm: the same through the pool with the GPU listed `lanes` times (pinned host frames)."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from chalkydri_b200 import capi  # noqa: E402
from chalkydri_b200.detector import DetectorBuilder, DET_DTYPE  # noqa: E402
from chalkydri_b200.pool import DetectorPool  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
BATCH = bench.BATCH
W, H = bench.WORKLOADS[wl]["W"], bench.WORKLOADS[wl]["H"]
frames, _ = bench.make_frames(wl, 0)
L = capi.lib()
hp = capi.pinned_array(frames.shape, np.uint8)
hp[...] = frames


def device_lanes(lanes):
    per = BATCH // lanes
    dets = [DetectorBuilder.default().add_family_bits("tag36h11", 3).device(0).capacity(W, H, per, 64).build() for _ in range(lanes)]
    d_frames = L.cb_device_alloc(dets[0].ctx, frames.nbytes)
    assert L.cb_memcpy_h2d(dets[0].ctx, d_frames, capi.ptr(hp), frames.nbytes) == 0
    outs = [capi.pinned_array((per, 64), DET_DTYPE) for _ in range(lanes)]
    cnts = [capi.pinned_array((per,), np.int32) for _ in range(lanes)]

    def run(i, n):
        for _ in range(n):
            dets[i].detect_batch_device(d_frames + i * per * W * H, per, H, W, out=outs[i], counts=cnts[i])

    def rounds(n):
        ts = [threading.Thread(target=run, args=(i, n)) for i in range(lanes)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    rounds(3)
    t0 = time.perf_counter()
    rounds(steps)
    dt = time.perf_counter() - t0
    ndet = int(sum(c.sum() for c in cnts))
    print(f"{wl} device arm, {lanes} lane(s) x {per} frames: {BATCH * steps / dt:9.0f} frames/s  ({dt / steps * 1e3:.2f} ms per 256 frames, {ndet} detections)", flush=True)
    L.cb_device_free(dets[0].ctx, d_frames)
    for d in dets:
        d.close()


def host_lanes(lanes):
    per = BATCH // (2 * lanes)
    out = np.zeros((BATCH, 64), DET_DTYPE)
    counts = np.zeros(BATCH, np.int32)
    pool = DetectorPool([0] * lanes, W, H, per, 64)
    for _ in range(2):
        pool.detect_batch(hp, out=out, counts=counts)
    t0 = time.perf_counter()
    for _ in range(steps):
        pool.detect_batch(hp, out=out, counts=counts)
    dt = time.perf_counter() - t0
    print(f"{wl} host frames through the pool, {lanes} lane(s), batches of {per}: {BATCH * steps / dt:9.0f} frames/s  ({int(counts.sum())} detections)", flush=True)
    pool.close()


for lanes in (1, 2, 4):
    device_lanes(lanes)
for lanes in (1, 2, 4):
    host_lanes(lanes)
