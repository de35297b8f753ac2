"""Multi-GPU sharding of the frame batch: independent frames, contiguous chunks per rank, no collective on the
data path.  The only exchange is the gather of the fixed-size detection records to rank 0 (`gather_detections`).

One process per GPU; `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is plumbing for the gather only.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous chunk [lo, hi) of rank `rank`: frames[g*B/N : (g+1)*B/N] with the remainder spread over the first ranks."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def camera_to_rank(camera: int, world: int) -> int:
    """Multi-camera stream (BASELINE configs[3]): camera c -> GPU c mod N."""
    return camera % world


def stream_shard(detector, frames: np.ndarray, batch: int, out: np.ndarray | None = None, counts: np.ndarray | None = None):
    """One rank's share of a frame stream through the streaming form of the detector call: `frames` [n,H,W] is cut into batches of
    `batch` frames, batch k+1 is submitted before batch k is collected (two in flight), and the records come back with `frame`
    = index inside `frames` (the library numbers frames inside a batch).  `detector` needs submit(frames) / collect(out=, counts=)
    / max_dets -- chalkydri_b200.detector.Detector, or a stand-in in the CPU tests."""
    from .capi import DET_DTYPE
    n = len(frames)
    if out is None:
        out = np.zeros((n, detector.max_dets), DET_DTYPE)
    if counts is None:
        counts = np.zeros(n, np.int32)
    starts = list(range(0, n, batch))
    if not starts:
        return out, counts
    detector.submit(frames[starts[0]:starts[0] + batch])
    for k, s in enumerate(starts):
        if k + 1 < len(starts):
            detector.submit(frames[starts[k + 1]:starts[k + 1] + batch])
        detector.collect(out=out[s:s + batch], counts=counts[s:s + batch])
        e = min(s + batch, n)
        out["frame"][s:e] = np.arange(s, e, dtype=np.int32)[:, None]      # batch-local -> shard-local frame index (every record of the row)
    return out, counts


class SharedDetections:
    """ONE host array for the whole job that every rank of the box maps (POSIX shared memory): `out` [n_total, cap] records and
    `counts` [n_total].  Rank r's streaming calls write its lists straight into out[lo:hi] / counts[lo:hi] -- "each GPU's D2H
    into its slice of one host array"  for the one-process-per-GPU launch: no collective, no re-upload of records,
    no padding.  The creating rank calls unlink() after everyone closed."""

    def __init__(self, name: str, n_total: int, cap: int, create: bool):
        from multiprocessing import shared_memory
        from .capi import DET_DTYPE
        rec = DET_DTYPE.itemsize
        self.nbytes = n_total * cap * rec + n_total * 4
        if create:
            try:
                shared_memory.SharedMemory(name=name).unlink()           # a stale segment of a crashed run
            except FileNotFoundError:
                pass
            self._shm = shared_memory.SharedMemory(name=name, create=True, size=self.nbytes)
        else:
            self._shm = shared_memory.SharedMemory(name=name)
        self._owner = create
        self.out = np.ndarray((n_total, cap), DET_DTYPE, buffer=self._shm.buf, offset=0)
        self.counts = np.ndarray((n_total,), np.int32, buffer=self._shm.buf, offset=n_total * cap * rec)

    def close(self):
        self.out = self.counts = None
        self._shm.close()

    def unlink(self):
        if self._owner:
            self._shm.unlink()


def stream_shard_into(detector, frames: np.ndarray, batch: int, shared: SharedDetections, lo: int):
    """stream_shard for rank-local `frames` = job frames [lo, lo + len(frames)): lists go straight into the shared array's slice,
    `frame` becomes the job-wide index."""
    hi = lo + len(frames)
    stream_shard(detector, frames, batch, out=shared.out[lo:hi], counts=shared.counts[lo:hi])
    shared.out["frame"][lo:hi] += lo


def gather_detections(local_out: np.ndarray, local_counts: np.ndarray, lo: int, n_total: int, dist=None, device=None):
    """Gather every rank's [n_local, cap] detection records + counts into rank 0's [n_total, cap] array, ordered by frame index,
    through torch.distributed (any backend).  This is the generic route -- across nodes, or when the ranks do not share a host;
    on one box the ranks write into a SharedDetections array instead and nothing is gathered at all.

    Returns (out, counts) on rank 0 and (None, None) elsewhere.  With dist=None (single process) it is the identity."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = local_out.copy()
        out["frame"] += 0
        return out, local_counts.copy()
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = local_out.shape[1]
    rec = local_out.dtype.itemsize
    max_local = -(-n_total // world)
    pad_out = np.zeros((max_local, cap), local_out.dtype)
    pad_out[:len(local_out)] = local_out
    pad_cnt = np.full(max_local, -1, np.int32)
    pad_cnt[:len(local_counts)] = local_counts
    dev = device if device is not None else "cpu"
    t_out = torch.from_numpy(pad_out.view(np.uint8).reshape(max_local, cap * rec)).to(dev)
    t_cnt = torch.from_numpy(pad_cnt).to(dev)
    t_lo = torch.tensor([lo, len(local_out)], dtype=torch.int64, device=dev)
    if rank == 0:
        g_out = [torch.empty_like(t_out) for _ in range(world)]
        g_cnt = [torch.empty_like(t_cnt) for _ in range(world)]
        g_lo = [torch.empty_like(t_lo) for _ in range(world)]
    else:
        g_out = g_cnt = g_lo = None
    dist.gather(t_out, g_out, dst=0)
    dist.gather(t_cnt, g_cnt, dst=0)
    dist.gather(t_lo, g_lo, dst=0)
    if rank != 0:
        return None, None
    out = np.zeros((n_total, cap), local_out.dtype)
    counts = np.zeros(n_total, np.int32)
    for r in range(world):
        l, n = (int(v) for v in g_lo[r].cpu())
        o = g_out[r].cpu().numpy().reshape(max_local, cap * rec).view(local_out.dtype).reshape(max_local, cap)[:n].copy()
        o["frame"] += l                   # frame index inside the whole job
        out[l:l + n] = o
        counts[l:l + n] = g_cnt[r].cpu().numpy()[:n]
    return out, counts
