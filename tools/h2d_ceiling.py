"""Host -> device copy ceiling of the box with N ranks copying at once (VERDICT r1 task 6): plain pinned cudaMemcpyAsync of the
bench's per-step payload on every rank simultaneously, no kernels.  bench.py's end-to-end arm is bounded by this number.

  python tools/h2d_ceiling.py                 (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_ceiling.py

Per rank: 236 MB (256 x 1280x720, the headline batch) and 405 MB (256 x 1456x1088) pinned buffers, 20 copies each, with and
without binding the rank's host thread (and therefore its pinned allocation, first-touched by this thread) to the CPUs NVML
reports as local to the GPU.  Reports GB/s per rank (min / max over ranks) and aggregate."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allr(x, op):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=op)
    return float(t.item())


def measure(nbytes, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(7)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    mine = nbytes * reps / (time.perf_counter() - t0) / 1e9
    barrier()
    wall = allr(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
    return {"per_rank_gbs_min": allr(mine, dist.ReduceOp.MIN if world > 1 else None), "per_rank_gbs_max": allr(mine, dist.ReduceOp.MAX if world > 1 else None),
            "aggregate_gbs": nbytes * reps * world / wall / 1e9}


res = {"n_gpus": world, "cpus": os.cpu_count()}
for label, pin in (("default_affinity", False), ("gpu_local_affinity", True)):
    note = ""
    if pin:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from bench import set_gpu_local_affinity             # the same binding bench.py's ranks use
        note = set_gpu_local_affinity(local)
    res[label] = {"affinity": note, "c1_236MB": measure(256 * 1280 * 720), "c2_405MB": measure(256 * 1456 * 1088)}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
