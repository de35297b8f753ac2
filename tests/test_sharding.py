"""N>1 path on CPU: world_size-2 gloo processes shard a frame batch and gather detection records to rank 0."""
import os
import socket

import numpy as np
import pytest

from chalkydri_b200.sharding import camera_to_rank, shard_range


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 256, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [camera_to_rank(c, 4) for c in range(6)] == [0, 1, 2, 3, 0, 1]


def _worker(rank, world, port, n_total, cap, tmp):
    import torch.distributed as dist
    from chalkydri_b200.capi import DET_DTYPE
    from chalkydri_b200.sharding import gather_detections, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    # each rank fabricates the records its frames would produce: frame f has f % 3 detections with id = f*10 + k
    out = np.zeros((hi - lo, cap), DET_DTYPE)
    counts = np.zeros(hi - lo, np.int32)
    for i, f in enumerate(range(lo, hi)):
        counts[i] = f % 3
        for k in range(counts[i]):
            out[i, k]["frame"] = i                 # local index, like the C ABI reports it
            out[i, k]["id"] = f * 10 + k
            out[i, k]["p"] = f + 0.5
    g_out, g_counts = gather_detections(out, counts, lo, n_total, dist)
    if rank == 0:
        np.save(os.path.join(tmp, "out.npy"), g_out)
        np.save(os.path.join(tmp, "counts.npy"), g_counts)
    else:
        assert g_out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total, cap = 11, 4
    mp.spawn(_worker, args=(2, port, n_total, cap, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "out.npy")
    counts = np.load(tmp_path / "counts.npy")
    assert counts.tolist() == [f % 3 for f in range(n_total)]
    for f in range(n_total):
        for k in range(counts[f]):
            assert out[f, k]["frame"] == f and out[f, k]["id"] == f * 10 + k and out[f, k]["p"][0, 0] == f + 0.5


class _StandInDetector:
    """submit / collect with the library's queue rules (two in flight, oldest first, frame = index inside the batch); frame f
    "contains" pixel-value-coded detections: count = value of its first pixel, ids derived from its second."""
    max_dets = 4

    def __init__(self):
        self.q = []

    def submit(self, frames):
        assert len(self.q) < 2, "more than two batches in flight"
        self.q.append(frames)

    def collect(self, out, counts):
        frames = self.q.pop(0)
        assert len(out) == len(frames) == len(counts)
        for b, f in enumerate(frames):
            counts[b] = int(f[0, 0])
            for k in range(counts[b]):
                out[b, k]["frame"] = b
                out[b, k]["id"] = int(f[0, 1]) * 10 + k


def _stream_worker(rank, world, port, n_total, batch, tmp):
    import torch.distributed as dist
    from chalkydri_b200.sharding import gather_detections, shard_range, stream_shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    frames = np.zeros((hi - lo, 2, 2), np.uint8)
    for i in range(lo, hi):
        frames[i - lo, 0, 0] = i % 4                     # detections in frame i
        frames[i - lo, 0, 1] = i % 25                    # id base
    out, counts = stream_shard(_StandInDetector(), frames, batch)
    g_out, g_counts = gather_detections(out, counts, lo, n_total, dist)
    if rank == 0:
        np.save(os.path.join(tmp, "s_out.npy"), g_out)
        np.save(os.path.join(tmp, "s_counts.npy"), g_counts)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stream_two_ranks_gloo(tmp_path):
    """BASELINE configs[3] on the CPU: a 23-frame stream over two ranks in batches of 5 (ragged last batches), the streaming call
    per rank, one gather -- every record ends up under its job-wide frame index."""
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total = 23
    mp.spawn(_stream_worker, args=(2, port, n_total, 5, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "s_out.npy")
    counts = np.load(tmp_path / "s_counts.npy")
    assert counts.tolist() == [f % 4 for f in range(n_total)]
    for f in range(n_total):
        for k in range(counts[f]):
            assert out[f, k]["frame"] == f and out[f, k]["id"] == (f % 25) * 10 + k


def _shared_worker(rank, world, port, n_total, batch, name, tmp):
    import torch.distributed as dist
    from chalkydri_b200.sharding import SharedDetections, shard_range, stream_shard_into
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    shared = SharedDetections(name, n_total, _StandInDetector.max_dets, create=True) if rank == 0 else None
    if rank == 0:
        shared.counts[:] = -1
    dist.barrier()
    if rank != 0:
        shared = SharedDetections(name, n_total, _StandInDetector.max_dets, create=False)
    lo, hi = shard_range(n_total, rank, world)
    frames = np.zeros((hi - lo, 2, 2), np.uint8)
    for i in range(lo, hi):
        frames[i - lo, 0, 0] = i % 4
        frames[i - lo, 0, 1] = i % 25
    stream_shard_into(_StandInDetector(), frames, batch, shared, lo)
    dist.barrier()                                        # the only synchronisation: every slice is in the array
    if rank == 0:
        np.save(os.path.join(tmp, "sh_out.npy"), shared.out.copy())
        np.save(os.path.join(tmp, "sh_counts.npy"), shared.counts.copy())
    shared.close()
    dist.barrier()
    if rank == 0:
        shared.unlink()
    dist.destroy_process_group()


def test_sharded_stream_into_one_shared_host_array(tmp_path):
    """The one-box form of BASELINE configs[3]: two ranks write their lists straight into their slices of ONE host array (POSIX
    shared memory); no gather, no collective on the data path (the barriers only order create / fill / read)."""
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total = 23
    mp.spawn(_shared_worker, args=(2, port, n_total, 5, f"cb_test_{os.getpid()}", str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "sh_out.npy")
    counts = np.load(tmp_path / "sh_counts.npy")
    assert counts.tolist() == [f % 4 for f in range(n_total)]
    for f in range(n_total):
        for k in range(counts[f]):
            assert out[f, k]["frame"] == f and out[f, k]["id"] == (f % 25) * 10 + k
