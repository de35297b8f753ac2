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
