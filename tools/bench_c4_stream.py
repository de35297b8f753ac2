"""BASELINE.json configs[3]: a multi-camera stream of 4096 1280x800 frames sharded over N B200 GPUs, detection lists gathered to
the host of rank 0.  Strong scaling: the 4096 frames are fixed, rank r takes the contiguous chunk shard_range(4096, r, N)
(no collective on the data path), feeds it in batches of 256 through the streaming form of the detector call (pinned host
frames, batch k+1 submitted before batch k is collected) and the fixed-size detection records are gathered once at the end.
The timed region covers every H2D copy, every kernel, every list read-back and the gather; time = max over ranks.

  python tools/bench_c4_stream.py                       (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_c4_stream.py
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from chalkydri_b200 import synth, capi
from chalkydri_b200.detector import DetectorBuilder, DET_DTYPE
from chalkydri_b200.sharding import shard_range, gather_detections, stream_shard

TOTAL, W, H, BATCH, CAP, UNIQUE = 4096, 1280, 800, 256, 16, 8
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

# frame i of the stream is unique frame i % UNIQUE ("cameras" take turns); 0..6 tags per frame
uniq, truths = [], []
for u in range(UNIQUE):
    f, t = synth.render_frame(W, H, u % 7, seed=0x5EED + 4 + 31 * u, edge_px=(40.0, 160.0))
    uniq.append(f); truths.append(len(t["ids"]))
lo, hi = shard_range(TOTAL, rank, world)
n = hi - lo
h = capi.pinned_array((n, H, W), np.uint8)
for i in range(n):
    h[i] = uniq[(lo + i) % UNIQUE]
out = np.zeros((n, CAP), DET_DTYPE); counts = np.zeros(n, np.int32)
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).device(local).capacity(W, H, BATCH, CAP).build()


def one_pass():
    stream_shard(det, h, BATCH, out=out, counts=counts)
    return gather_detections(out, counts, lo, TOTAL, dist if world > 1 else None, device=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


one_pass()
walls = []
for _ in range(reps):
    barrier()
    t0 = time.perf_counter()
    g_out, g_counts = one_pass()
    barrier()
    walls.append(time.perf_counter() - t0)
wall = float(np.median(walls))
if world > 1:
    t_ = torch.tensor([wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    wall = float(t_.item())
if rank == 0:
    want = sum(truths[i % UNIQUE] for i in range(TOTAL))
    frames_in_order = bool((g_out["frame"][g_counts > 0, 0] == np.nonzero(g_counts > 0)[0]).all())
    print(json.dumps({"config": "c4", "workload": f"{TOTAL} x {W}x{H} frames, 0-6 tags each, sharded over {world} GPU(s), lists gathered to rank 0",
                      "scaling": "strong", "n_gpus": world, "value": TOTAL / wall, "unit": "frames/s", "ms_per_pass": wall * 1e3,
                      "passes": reps, "detections": int(g_counts.sum()), "ground_truth_tags": int(want), "gathered_in_frame_order": frames_in_order,
                      "h2d_bytes_per_pass": TOTAL * W * H, "gathered_bytes_per_pass": int(g_out.nbytes + g_counts.nbytes),
                      "api": "cb_detect_gray_submit / cb_detect_gray_collect per rank + one dist.gather of the records"}), flush=True)
det.close()
if world > 1:
    dist.destroy_process_group()
