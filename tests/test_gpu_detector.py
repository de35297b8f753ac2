"""GPU parity tests of the detector: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): threshold bitmap, component partition, tag ids and hamming BIT-EXACT; corners within
1e-3 px; decision margin within 1e-3 relative.  Full-size cases are checked through size-independent properties."""
import os

import numpy as np
import pytest

from chalkydri_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CORNER_TOL = 1e-3       # px (north_star)
MARGIN_RTOL = 1e-3


def make_detector(W, H, B=1, dets=128):
    from chalkydri_b200.detector import DetectorBuilder
    return DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, dets).build()


def canon_quads(q):
    out = []
    for c in q:
        c = np.asarray(c, np.float64).reshape(4, 2)
        k = min(range(4), key=lambda i: (c[i, 0], c[i, 1]))
        out.append(tuple(np.roll(c, -k, 0).reshape(-1)))
    return sorted(out)


def assert_same_detections(got, ref):
    assert got["id"].tolist() == ref["id"].tolist()
    assert got["hamming"].tolist() == ref["hamming"].tolist()
    if len(got):
        assert np.abs(got["p"] - ref["p"]).max() < CORNER_TOL
        assert np.abs(got["c"] - ref["c"]).max() < CORNER_TOL
        assert np.allclose(got["decision_margin"], ref["decision_margin"], rtol=MARGIN_RTOL, atol=1e-3)
        assert np.allclose(got["H"], ref["H"], rtol=1e-4, atol=2e-3)


CASES = [  # W, H, tags, seed, edge range
    (1280, 720, 4, 1, (60, 150)),      # c1
    (1456, 1088, 8, 2, (40, 200)),     # c2 frame
    (1280, 800, 6, 3, (40, 160)),      # c4 frame
    (642, 486, 3, 4, (40, 100)),       # decimates to 321x243: partial tiles, unaligned rows -> generic kernels
    (1000, 750, 5, 5, (40, 120)),      # w = 500 (multiple of 4) but stride not a multiple of 16
]


@pytest.mark.parametrize("W,H,tags,seed,edge", CASES)
def test_stages_bit_exact(oracle, W, H, tags, seed, edge):
    frames, _ = synth.render_batch(W, H, 2, tags, seed=seed, edge_px=edge)
    det = make_detector(W, H, 2)
    thr = det.threshold(frames)
    lab, sz = det.labels(frames)
    q, qc, _ = det.quads(frames)
    out, counts = det.detect_batch(frames)
    for b in range(2):
        ref, taps = oracle.detect(frames[b], taps=True)
        assert (thr[b] == taps["thresh"]).all(), "threshold bitmap differs"
        assert (lab[b] == taps["labels"]).all(), "component partition differs"
        assert (sz[b] == taps["comp_size"]).all()
        assert qc[b] == taps["nquads"]
        gq, oq = canon_quads(q[b, :qc[b]]), canon_quads(taps["quads"]["p"])
        assert np.abs(np.array(gq) - np.array(oq)).max() < 1e-4 if gq else True
        assert_same_detections(out[b, :counts[b]], ref)
    det.close()


def _axis_aligned_frame(W=1280, H=720):
    """Pixel-replicated upright tags (every edge axis aligned: whole runs of points share one slope key), plus thin
    symmetric outlines whose border-polarity sum cancels to ~0 (the sign then depends on the summation order)."""
    img = np.full((H, W), 150, np.uint8)
    x = 40
    for k, (tag_id, cell) in enumerate([(3, 8), (17, 12), (101, 6), (250, 16), (586, 10), (42, 5)]):
        pat = np.kron(synth.tag_pattern(tag_id), np.ones((cell, cell), np.uint8))
        n = pat.shape[0]
        y = 30 + (k % 2) * 330
        img[y:y + n, x:x + n] = np.where(pat > 0, 230, 25)
        x += n + 30
    for k in range(6):                                      # outlines 2 / 4 px thick, several sizes
        x0, y0, sz, th = 60 + k * 180, 560, 60 + 14 * k, 2 + 2 * (k % 2)
        img[y0:y0 + sz, x0:x0 + sz] = 30
        img[y0 + th:y0 + sz - th, x0 + th:x0 + sz - th] = 150
    img[700:704, 100:1100] = 20                            # long thin lines
    img[100:620, 1240:1243] = 235
    return img


def test_axis_aligned_tags_slope_ties_and_thin_outlines(oracle):
    """ptsort()'s order among equal slope keys and the sign of near-cancelling polarity sums must match upstream exactly."""
    frame = _axis_aligned_frame()
    det = make_detector(1280, 720, 2)
    frames = np.stack([frame, np.ascontiguousarray(frame[::-1, ::-1])])       # second frame: 180 degree turn, other tie patterns
    q, qc, _ = det.quads(frames)
    out, counts = det.detect_batch(frames)
    for b in range(2):
        ref, taps = oracle.detect(frames[b], taps=True)
        assert qc[b] == taps["nquads"] and qc[b] >= 6
        gq, oq = canon_quads(q[b, :qc[b]]), canon_quads(taps["quads"]["p"])
        assert np.abs(np.array(gq) - np.array(oq)).max() < 1e-4
        assert_same_detections(out[b, :counts[b]], ref)
        assert len(ref) >= 5
    det.close()


def test_huge_cluster_uses_global_work_arrays(oracle):
    """A cluster above 8192 points (here the outline of a 4000x2300 rectangle, ~12.6 k points, below upstream's
    3(2w+2h) cap) no longer fits the largest shared-memory tier: both sorts run out of the global scratch arrays."""
    H, W = 2592, 4608
    rng = np.random.default_rng(7)
    img = np.full((H, W), 170, np.uint8)
    img[140:2440, 300:4300] = 40
    img[400:700, 800:1100] = 200                           # some structure inside, plus a tag so the decode stage has work
    pat = np.kron(synth.tag_pattern(5), np.ones((24, 24), np.uint8))
    img[1000:1240, 2000:2240] = np.where(pat > 0, 235, 20)
    img = np.clip(img.astype(np.int16) + rng.integers(-1, 2, img.shape), 0, 255).astype(np.uint8)
    det = make_detector(W, H, 1, 64)
    q, qc, _ = det.quads(img[None])
    out, counts = det.detect_batch(img[None])
    ref, taps = oracle.detect(img, taps=True, pts_cap=8_000_000)
    _, sizes = np.unique(taps["pts_cluster"][:taps["npoints"]], return_counts=True)
    assert sizes.max() > 8192, "the fixture no longer produces a cluster above the shared-memory tiers"
    assert qc[0] == taps["nquads"] and qc[0] >= 2
    gq, oq = canon_quads(q[0, :qc[0]]), canon_quads(taps["quads"]["p"])
    assert np.abs(np.array(gq) - np.array(oq)).max() < 1e-4
    assert_same_detections(out[0, :counts[0]], ref)
    assert 5 in ref["id"].tolist()
    det.close()


def test_shared_reciprocal_division_is_the_compilers_division(tmp_path):
    """fit_line() divides five moments by one weight through cb::DivBy; it must equal a / b bit for bit (2e10 pairs)."""
    import subprocess, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "divby_check")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false", "-O3", "-I",
                           os.path.join(root, "chalkydri_b200", "csrc"), "-o", exe, os.path.join(root, "tests", "cuda", "divby_check.cu")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_randomised_frames_match_oracle(oracle):
    """tests/fuzz_parity.py: odd sizes, random shapes / lines / noise / replicated tags; every stage compared with the oracle."""
    import subprocess, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_parity.py"), "16", "2024"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_small_capacity_context_large_batch(oracle):
    """A context with room for 8 frames fed 72: the pipelined host path must never put more than half its capacity in flight."""
    frames, _ = synth.render_batch(640, 480, 72, 2, seed=5, unique=6, edge_px=(50, 110))
    det = make_detector(640, 480, 8)
    out, counts = det.detect_batch(frames)
    for b in (0, 1, 5, 7, 8, 30, 71):
        assert_same_detections(out[b, :counts[b]], oracle.detect(frames[b]))
    assert counts.sum() >= 72
    det.close()


def test_cluster_tile_table_overflow_path(oracle, monkeypatch):
    """CB_TILE_PROBES=1 makes every hash collision in the per-tile cluster table take the overflow path (global table)."""
    monkeypatch.setenv("CB_TILE_PROBES", "1")
    frames, _ = synth.render_batch(1456, 1088, 2, 8, seed=2, edge_px=(40, 200))
    det = make_detector(1456, 1088, 2)
    q, qc, _ = det.quads(frames)
    for b in range(2):
        _, taps = oracle.detect(frames[b], taps=True)
        assert qc[b] == taps["nquads"]
        gq, oq = canon_quads(q[b, :qc[b]]), canon_quads(taps["quads"]["p"])
        assert np.abs(np.array(gq) - np.array(oq)).max() < 1e-4
    det.close()


def test_strided_frame_with_partial_tiles(oracle):
    """stride != width and w, h not multiples of 4: vector loads + upstream's remainder rule (last full tile, never 127)."""
    from chalkydri_b200.detector import Image
    full, _ = synth.render_frame(1456, 1088, 8, seed=12, edge_px=(40, 200))
    W, H = 1450, 1084                                     # decimates to 725 x 542
    crop = np.ascontiguousarray(full[:H, :W])
    det = make_detector(1456, 1088, 1)
    got = det.detect(Image(full, W, H, 1456))
    ref = oracle.detect(crop)
    assert [d.id() for d in got] == ref["id"].tolist() and [d.hamming() for d in got] == ref["hamming"].tolist()
    for d, r in zip(got, ref):
        assert np.abs(np.array(d.corners()) - r["p"]).max() < CORNER_TOL
    # threshold tap on a buffer whose stride is a multiple of 16 but whose decimated size leaves partial tiles
    padded = np.zeros((H, 1456), np.uint8)
    padded[:, :W] = crop
    import ctypes as C
    from chalkydri_b200 import capi
    out = np.empty((542, 725), np.uint8)
    rc = det._L.cb_threshold(det.ctx, capi.ptr(padded), W, H, 1456, 1456 * H, 1, capi.ptr(out))
    assert rc == 0
    assert (out == oracle.threshold(crop)).all()
    det.close()


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
def test_every_threshold_kernel_shape_bit_exact(oracle, monkeypatch, cfg):
    """The five shapes of the tensor-map threshold kernel (tiles per lane x ring depth x store form, api.cu kThrVariants) and a
    few segment heights, on widths that give one strip, several strips, an odd number of tiles (strip tail stored directly next to the
    bulk stores) and partial tiles: identical to the oracle's threshold map byte for byte."""
    monkeypatch.setenv("CB_THR_CFG", str(cfg))
    for (W, H), ysegs in (((1280, 720), 7), ((1456, 1088), 0), ((3024, 808), 3), ((1448, 240), 2), ((1450, 1084), 5)):
        monkeypatch.setenv("CB_THR_YSEGS", str(ysegs)) if ysegs else monkeypatch.delenv("CB_THR_YSEGS", raising=False)
        stride = (W + 15) // 16 * 16
        rng = np.random.default_rng(W + cfg)
        frames = np.zeros((2, H, stride), np.uint8)
        base, _ = synth.render_batch(stride, H, 2, 3, seed=W, edge_px=(40, 120))
        frames[:] = base
        frames[1, :, : stride // 2] = rng.integers(0, 256, (H, stride // 2), dtype=np.uint8)      # noise: every tile a different threshold
        det = make_detector(stride, H, 2)           # a fresh context: the plan cache is per context, the hooks are read per plan
        w, h = (W + 1) // 2, (H + 1) // 2
        out = np.empty((2, h, w), np.uint8)
        from chalkydri_b200 import capi
        rc = det._L.cb_threshold(det.ctx, capi.ptr(frames), W, H, stride, stride * H, 2, capi.ptr(out))
        assert rc == 0, det._L.cb_last_error(det.ctx)
        for b in range(2):
            assert (out[b] == oracle.threshold(np.ascontiguousarray(frames[b, :, :W]))).all(), (cfg, W, H, ysegs, b)
        det.close()


def test_threshold_shape_timing_keeps_the_map(oracle, monkeypatch):
    """CB_THR_TUNE=1 times the kernel shapes once per geometry on large batches (threshold_plan): whatever shape wins, the map is the
    oracle's."""
    monkeypatch.setenv("CB_THR_TUNE", "1")
    monkeypatch.delenv("CB_THR_CFG", raising=False)
    monkeypatch.delenv("CB_THR_YSEGS", raising=False)
    frames, _ = synth.render_batch(1280, 720, 48, 4, seed=77, unique=3, edge_px=(60, 150))        # 44 MB: above the timing threshold
    det = make_detector(1280, 720, 48)
    thr = det.threshold(frames)
    for b in (0, 1, 2, 47):
        assert (thr[b] == oracle.threshold(frames[b])).all()
    det.close()


def test_graph_replay_of_the_single_frame_pipeline(oracle):
    """From its second call with one geometry on, a small batch runs as ONE captured CUDA graph (api.cu detect_device_chunk): every
    replay, on different frame contents, gives the oracle's detections; changing a parameter drops the graphs and takes effect."""
    frames, _ = synth.render_batch(1280, 720, 5, 4, seed=21, unique=5, edge_px=(60, 150))
    det = make_detector(1280, 720, 1)
    for k in (0, 1, 2, 3, 4, 1):
        out, counts = det.detect_batch(frames[k][None])
        assert_same_detections(out[0, :counts[0]], oracle.detect(frames[k]))
    t = det.timing()
    assert t["total_ms"] > 0 and t["kernel_launches"] >= 20            # the graph keeps the launch count it was captured with
    # bits_corrected is a kernel parameter baked into the graph: a tag with two flipped bits decodes at 3, not at 1
    frame, truth = synth.render_frame(1280, 720, 3, seed=5, edge_px=(80, 150), bit_errors=2)
    for _ in range(3):
        out, counts = det.detect_batch(frame[None])
    ref3 = oracle.detect(frame)
    assert_same_detections(out[0, :counts[0]], ref3)
    assert det._L.cb_set_family_tag36h11(det.ctx, 1) == 0
    for _ in range(3):
        out, counts = det.detect_batch(frame[None])
    ref1 = oracle.detect(frame, oracle.default_params(bits_corrected=1))
    assert_same_detections(out[0, :counts[0]], ref1)
    assert len(ref1) < len(ref3)
    det.close()


@pytest.mark.parametrize("W,H,tags,seed,edge", [(1280, 720, 4, 1, (60, 150)), (642, 486, 3, 4, (40, 100)), (1456, 1088, 8, 2, (40, 200))])
def test_gradient_clusters_bit_exact(oracle, W, H, tags, seed, edge):
    """Row A4 directly (cb_clusters): every cluster fit_quad() can accept holds exactly upstream's boundary points -- as a set against
    the oracle's point dump, and in upstream's append order (scan order y, x, probe) inside the cluster -- and no such cluster is
    missing.  Cluster ids are not compared (union-find representatives are an implementation detail); point sets are."""
    frames, _ = synth.render_batch(W, H, 2, tags, seed=seed, edge_px=edge)
    frames[1, : H // 2] = np.random.default_rng(seed).integers(0, 256, (H // 2, W), dtype=np.uint8)      # noise: thousands of small clusters
    det = make_detector(W, H, 2)
    pts, cl, ncl = det.clusters(frames)
    w, h = (W + 1) // 2, (H + 1) // 2
    lim = 3 * (2 * w + 2 * h)
    base = 0
    for b in range(2):
        _, taps = oracle.detect(frames[b], taps=True, pts_cap=4_000_000)
        ids, first, counts = np.unique(taps["pts_cluster"], return_index=True, return_counts=True)      # (the dump is grouped by cluster)
        want = {frozenset(map(tuple, taps["pts"][f:f + c].tolist())) for f, c in zip(first, counts) if 24 <= c <= lim}
        got = set()
        for c in range(base, base + int(ncl[b])):
            p = pts[cl == c]
            got.add(frozenset(map(tuple, p.tolist())))
            assert len(p) >= 24 and len(set(map(tuple, p.tolist()))) == len(p), "a point twice in one cluster"
            rows = (p[:, 1].astype(np.int64) - (p[:, 3] != 0)) // 2                  # pixel row of the probing pixel
            assert (np.diff(rows) >= 0).all(), "the points of a cluster are not in scan order"
        base += int(ncl[b])
        assert len(got) == ncl[b] == len(want) and got == want
    assert base == cl.max() + 1
    det.close()


def test_overflowing_frame_is_isolated(oracle, monkeypatch):
    """A frame that overflows a per-frame device table (here: the candidate-quad list, shrunk through a test hook) reports an empty
    list and its flag; the other frames of the batch are complete and the call succeeds -- in the blocking call, the chunked blocking
    call and the streaming form."""
    monkeypatch.setenv("CB_TEST_QUADS_PER_FRAME", "600")
    from chalkydri_b200 import capi
    W, H, B = 1280, 720, 72                                   # 72 frames: the blocking call takes its pipelined (chunked) path
    frames, _ = synth.render_batch(W, H, B, 4, seed=9, unique=4, edge_px=(60, 150))
    noisy = (5, 40, 71)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in noisy:                                           # a field of 16-pixel dark squares: more than a thousand candidate quads
        frames[b] = np.where(((xx + 3 * b) % 28 < 16) & ((yy + b) % 28 < 16), 30, 220).astype(np.uint8)
    det = make_detector(W, H, B)
    refs = {b: oracle.detect(frames[b]) for b in (0, 1, 6, 39, 70)}

    def check(out, counts):
        flags = det.frame_flags(B)
        assert [b for b in range(B) if flags[b]] == list(noisy) and all(flags[b] & 8 for b in noisy)
        assert all(counts[b] == 0 for b in noisy)
        for b, ref in refs.items():
            assert_same_detections(out[b, :counts[b]], ref)

    check(*det.detect_batch(frames))                          # pipelined blocking call
    o8, c8 = det.detect_batch(frames[:8])                     # simple path
    f8 = det.frame_flags(8)
    assert f8[5] & 8 and c8[5] == 0 and not f8[:5].any() and not f8[6:].any()
    assert_same_detections(o8[0, :c8[0]], refs[0])
    pin = capi.pinned_array(frames.shape, np.uint8); pin[:] = frames
    det.submit(pin)
    check(*det.collect())                                     # streaming form
    capi.free_pinned(pin)
    det.close()


def test_c3_full_resolution_small_tags(oracle):
    frame, truth = synth.render_frame(4608, 2592, 40, seed=4, edge_px=(40, 300), small_tags=10)
    det = make_detector(4608, 2592, 1, 256)
    out, counts = det.detect_batch(frame[None])
    ref = oracle.detect(frame)
    assert_same_detections(out[0, :counts[0]], ref)
    assert counts[0] >= 28
    det.close()


def test_golden_fixture(oracle):
    g = np.load(os.path.join(GOLD, "detector_c1.npz"))
    det = make_detector(1280, 720, 1)
    out, counts = det.detect_batch(g["frame"][None])
    d = out[0, :counts[0]]
    assert d["id"].tolist() == g["ids"].tolist() and d["hamming"].tolist() == g["hamming"].tolist()
    assert np.abs(d["p"] - g["corners"]).max() < CORNER_TOL
    thr = det.threshold(g["frame"][None])[0]
    assert (np.packbits(thr == 255) == g["thresh_white_bits"]).all() and (np.packbits(thr == 0) == g["thresh_black_bits"]).all()
    q, qc, _ = det.quads(g["frame"][None])
    assert qc[0] == int(g["nquads"])
    assert np.abs(np.array(canon_quads(q[0, :qc[0]])) - np.array(canon_quads(g["quads"]))).max() < 1e-4
    det.close()


def test_edge_cases(oracle):
    det = make_detector(640, 480, 3)
    flat = np.full((480, 640), 128, np.uint8)
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    checker = ((np.add.outer(np.arange(480) // 16, np.arange(640) // 16) % 2) * 200 + 20).astype(np.uint8)
    frames = np.stack([flat, noise, checker])
    thr = det.threshold(frames)
    lab, sz = det.labels(frames)
    out, counts = det.detect_batch(frames)
    for b in range(3):
        ref, taps = oracle.detect(frames[b], taps=True)
        assert (thr[b] == taps["thresh"]).all() and (lab[b] == taps["labels"]).all() and (sz[b] == taps["comp_size"]).all()
        assert_same_detections(out[b, :counts[b]], ref)
    assert counts[0] == 0 and (thr[0] == 127).all()
    det.close()


def test_batch_64_c2_matches_oracle(oracle):
    frames, truths = synth.render_batch(1456, 1088, 64, 8, seed=11, unique=8, edge_px=(40, 200))
    det = make_detector(1456, 1088, 64, 64)
    out, counts = det.detect_batch(frames)
    ref, rc = oracle.detect_batch(frames, cap=64, nthreads=os.cpu_count() or 1)
    assert counts.tolist() == rc.tolist()
    for b in range(64):
        assert_same_detections(out[b, :counts[b]], ref[b, :rc[b]])
    # every rendered tag of 40+ px edge is found
    found = sum(int(np.isin(t["ids"], out[b, :counts[b]]["id"]).sum()) for b, t in enumerate(truths))
    assert found >= 0.97 * sum(len(t["ids"]) for t in truths)
    det.close()


def test_full_size_properties_c2():
    """256 x 1456x1088 (BASELINE configs[1]): chunking, determinism, frame-order independence."""
    frames, truths = synth.render_batch(1456, 1088, 256, 8, seed=21, unique=4, edge_px=(40, 200))
    det = make_detector(1456, 1088, 96, 64)          # max_batch 96 < 256 exercises the internal chunking
    out, counts = det.detect_batch(frames)
    out2, counts2 = det.detect_batch(frames)
    assert counts.tolist() == counts2.tolist() and out.tobytes() == out2.tobytes(), "not deterministic"
    perm = np.random.default_rng(0).permutation(256)
    outp, countsp = det.detect_batch(np.ascontiguousarray(frames[perm]))
    for i, b in enumerate(perm):
        assert countsp[i] == counts[b]
        a, c = outp[i, :countsp[i]].copy(), out[b, :counts[b]].copy()
        a["frame"] = 0; c["frame"] = 0
        assert a.tobytes() == c.tobytes(), "result depends on the position inside the batch"
    assert (out["frame"][np.arange(256)[:, None].repeat(64, 1) < 0].size == 0)
    for b in range(256):
        assert (out[b, :counts[b]]["frame"] == b).all()
        ids = out[b, :counts[b]]["id"]
        assert (np.diff(ids) >= 0).all(), "detections must be sorted by id"
        assert np.isin(truths[b]["ids"], ids).mean() >= 0.75
    det.close()


def test_rgb_and_yuyv_inputs(oracle):
    gray, _ = synth.render_frame(1280, 720, 4, seed=8, edge_px=(60, 150))
    det = make_detector(1280, 720, 2)
    rgb = synth.gray_to_rgb(gray, seed=1)
    g_ref = np.array([oracle.cat_grayscale(int(r), int(g), int(b)) for r, g, b in rgb.reshape(-1, 3)[:4096]], np.uint8)
    out, counts = det.detect_rgb_batch(rgb[None])
    # oracle on the oracle's own gray conversion of the same RGB frame
    lut = {}
    flat = rgb.reshape(-1, 3)
    keys = flat[:, 0].astype(np.uint32) << 16 | flat[:, 1].astype(np.uint32) << 8 | flat[:, 2]
    uk, inv = np.unique(keys, return_inverse=True)
    gv = np.array([oracle.cat_grayscale(int(k >> 16), int((k >> 8) & 255), int(k & 255)) for k in uk], np.uint8)
    gray2 = gv[inv].reshape(gray.shape)
    assert (gray2.reshape(-1)[:4096] == g_ref).all()
    assert_same_detections(out[0, :counts[0]], oracle.detect(gray2))
    yuyv = np.empty((720, 1280 * 2), np.uint8)
    yuyv[:, 0::2] = gray
    yuyv[:, 1::2] = 128
    out, counts = det.detect_yuyv_batch(yuyv[None])
    assert_same_detections(out[0, :counts[0]], oracle.detect(gray))
    det.close()


def test_reference_call_shape(oracle):
    """The reference's own call: one frame in, Vec<Detection> out (crates/apriltags/src/lib.rs:301-314)."""
    from chalkydri_b200.detector import Image
    gray, truth = synth.render_frame(1280, 720, 4, seed=9, edge_px=(60, 150))
    det = make_detector(1280, 720, 1)
    dets = det.detect(Image(gray))
    assert sorted(d.id() for d in dets) == sorted(truth["ids"].tolist())
    for d in dets:
        k = truth["ids"].tolist().index(d.id())
        assert np.abs(np.array(d.corners()) - truth["corners"][k]).max() < 0.5
        assert d.hamming() == 0 and d.decision_margin() > 20 and d.homography().shape == (3, 3)
    det.close()


def test_error_behaviour():
    from chalkydri_b200.capi import ChalkydriError
    det = make_detector(640, 480, 1)
    with pytest.raises(ChalkydriError):
        det.detect_batch(np.zeros((1, 600, 800), np.uint8))          # larger than the context
    with pytest.raises(ChalkydriError):
        det.set_params(quad_sigma=0.8)                                 # not implemented: loud, not silent
    with pytest.raises(ChalkydriError):
        det._check(det._L.cb_set_family_tag36h11(det.ctx, 4))
    det.close()


def test_streaming_submit_collect_matches_the_synchronous_call(oracle):
    """cb_detect_gray_submit / _collect: batches of different sizes, two in flight, results identical to cb_detect_gray
    (bit for bit: same kernels, same chunk rule apart from the chunk sizes) and to the oracle on sampled frames."""
    from chalkydri_b200 import capi
    det = make_detector(640, 480, 8)
    batches = []
    for i, n in enumerate((8, 5, 1, 8, 3)):
        f, _ = synth.render_batch(640, 480, n, 2, seed=20 + i, unique=min(n, 4), edge_px=(50, 110))
        h = capi.pinned_array(f.shape, np.uint8)
        h[...] = f
        batches.append(h)
    want = [tuple(a.copy() for a in det.detect_batch(b)) for b in batches]
    got = []
    det.submit(batches[0])
    for k in range(len(batches)):
        if k + 1 < len(batches):
            det.submit(batches[k + 1])
            assert det.pending == 2
        got.append(det.collect())
    assert det.pending == 0
    for (o, c), (wo, wc), b in zip(got, want, batches):
        assert c.tolist() == wc.tolist()
        for i in range(len(c)):
            assert o[i, :c[i]].tobytes() == wo[i, :c[i]].tobytes()
        assert_same_detections(o[0, :c[0]], oracle.detect(np.asarray(b[0])))
    det.detect_batch(batches[1])                  # the synchronous call works again once the queue is empty
    det.close()
    for b in batches:
        capi.free_pinned(b)


def test_streaming_queue_rules():
    from chalkydri_b200.capi import ChalkydriError, CB_ERR_STATE
    det = make_detector(640, 480, 4)
    f = np.full((2, 480, 640), 128, np.uint8)
    with pytest.raises(ChalkydriError) as e:
        det.collect()                                                   # nothing submitted
    assert e.value.code == CB_ERR_STATE
    with pytest.raises(ChalkydriError):
        det.submit(np.zeros((5, 480, 640), np.uint8))                   # more than max_batch
    det.submit(f)
    det.submit(f)
    with pytest.raises(ChalkydriError) as e:
        det.submit(f)                                                   # a third batch in flight
    assert e.value.code == CB_ERR_STATE
    with pytest.raises(ChalkydriError) as e:
        det.detect_batch(f)                                             # synchronous call while batches are in flight
    assert e.value.code == CB_ERR_STATE
    with pytest.raises(ChalkydriError):
        det.threshold(f)
    _, c = det.collect()
    assert c.tolist() == [0, 0]
    det.collect()
    assert det.pending == 0
    det.close()
    det2 = make_detector(640, 480, 4)              # destroying a context with a batch still in flight is safe
    det2.submit(f)
    det2.close()


@pytest.mark.parametrize("W,H", [(1000, 750), (642, 486)])
def test_rgb_and_yuyv_batches_with_ragged_and_unaligned_planes(oracle, W, H):
    """The batched pre-processing launch: 1000x750 frames end in a partial 512-pixel chunk; 642x486 frames make the second and
    third frame's RGB / YUYV planes start off a 16-byte boundary (scalar path).  Gray must equal utils.rs:43 bit for bit, which
    the detections on the oracle's gray conversion of the same frames confirm."""
    from tests.test_oracle_cat import gray_np
    B = 3
    grays, _ = synth.render_batch(W, H, B, 3, seed=50, edge_px=(50, 110))
    det = make_detector(W, H, B)
    rgb = np.stack([synth.gray_to_rgb(grays[b], seed=b) for b in range(B)])
    out, counts = det.detect_rgb_batch(rgb)
    g2 = gray_np(rgb)
    assert g2[0, 0, 0] == oracle.cat_grayscale(*(int(v) for v in rgb[0, 0, 0]))
    for b in range(B):
        assert_same_detections(out[b, :counts[b]], oracle.detect(g2[b]))
    assert counts.sum() >= 6
    yuyv = np.empty((B, H, W * 2), np.uint8)
    yuyv[:, :, 0::2] = grays
    yuyv[:, :, 1::2] = 77
    out, counts = det.detect_yuyv_batch(yuyv)
    for b in range(B):
        assert_same_detections(out[b, :counts[b]], oracle.detect(grays[b]))
    det.close()


def test_rgb_to_gray_every_colour_bit_exact():
    """All 2^24 RGB triples through rgb_to_gray_kernel's vector path (TMA-staged chunks, conversion-free arithmetic) against the
    exact-FMA numpy form of utils.rs:43 that tests/test_oracle_cat.py pins to the oracle; then ragged and unaligned planes
    (scalar path) and YUYV."""
    from tests.test_oracle_cat import gray_np
    det = make_detector(640, 480, 1)
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8).reshape(1, 4096, 4096, 3)
    got = det.rgb_to_gray(rgb)
    for lo in range(0, 4096, 512):                                    # in slabs: the float64 emulation is memory hungry
        assert (got[0, lo:lo + 512] == gray_np(rgb[0, lo:lo + 512])).all()
    rng = np.random.default_rng(3)
    small = rng.integers(0, 256, (3, 486, 642, 3), dtype=np.uint8)    # plane size 936036 B: frames 1 and 2 start unaligned, ragged tails
    assert (det.rgb_to_gray(small) == gray_np(small)).all()
    yuyv = rng.integers(0, 256, (3, 486, 642 * 2), dtype=np.uint8)
    assert (det.yuyv_to_gray(yuyv) == yuyv[:, :, 0::2]).all()
    yuyv = rng.integers(0, 256, (2, 750, 2000), dtype=np.uint8)       # aligned planes with a partial last chunk
    assert (det.yuyv_to_gray(yuyv) == yuyv[:, :, 0::2]).all()
    det.close()


# ---- round 2: error-corrected decode, family bits, reconcile, decimation factors, 4:2:0 buffers ------------------------------

def make_detector_bits(W, H, bits, B=1, dets=128):
    from chalkydri_b200.detector import DetectorBuilder
    return DetectorBuilder.default().add_family_bits("tag36h11", bits).capacity(W, H, B, dets).build()


@pytest.mark.parametrize("bits", [0, 1, 2, 3])
def test_hamming_decode_and_family_bits(oracle, bits):
    """add_family_bits(family, bits) for bits 0..3 (the reference configures 3, falls back to 1: crates/apriltags/src/lib.rs:230,
    279-282) on tags carrying 0..4 inverted data bits: ids and hamming bit-exact against the oracle, and against what the
    rendering says must decode (k <= bits -> hamming k, otherwise nothing)."""
    from tests import frames as fr
    frames, truths = [], []
    for k in range(5):
        im, t = synth.render_frame(1280, 720, 4, seed=31, edge_px=(70, 150), bit_errors=k)
        frames.append(im); truths.append(t)
    im, t = fr.bit_error_frame()
    frames.append(im); truths.append(t)
    frames = np.stack(frames)
    det = make_detector_bits(1280, 720, bits, B=len(frames))
    out, counts = det.detect_batch(frames)
    for b, t in enumerate(truths):
        got = out[b, :counts[b]]
        assert_same_detections(got, oracle.detect(frames[b], oracle.default_params(bits_corrected=bits)))
        want = {int(i): int(h) for i, h in zip(t["ids"], t["hamming"]) if h <= bits}
        assert {int(d["id"]): int(d["hamming"]) for d in got} == want
    det.close()


def test_bit_error_golden_fixture():
    """tests/golden/detector_biterr.npz (hamming 0..3 in one frame) with no oracle in the loop, plus OpenCV's decode of the same
    frame with maxCorrectionBits = k (tests/golden/detector_cv2_pin.npz): the same set of tags."""
    g = np.load(os.path.join(GOLD, "detector_biterr.npz"))
    pin = np.load(os.path.join(GOLD, "detector_cv2_pin.npz"))
    for bits in range(4):
        det = make_detector_bits(1280, 720, bits)
        out, counts = det.detect_batch(g["frame"][None])
        d = out[0, :counts[0]]
        assert d["id"].tolist() == g[f"ids_b{bits}"].tolist() == pin[f"biterr_ids_k{bits}"].tolist()
        assert d["hamming"].tolist() == g[f"hamming_b{bits}"].tolist()
        if len(d):
            assert np.abs(d["p"] - g[f"corners_b{bits}"]).max() < CORNER_TOL
            assert np.allclose(d["decision_margin"], g[f"margin_b{bits}"], rtol=MARGIN_RTOL, atol=1e-3)
        det.close()


def test_reconcile_overlapping_duplicates(oracle):
    """Two detections of one id with overlapping polygons (a small copy of a tag inside one of its own white cells): upstream's
    swap-remove reconcile keeps the lower hamming, then the higher decision margin.  All five cases in one batch."""
    from tests import frames as fr
    frames = np.stack([fr.nested_same_id_frame(**kw) for kw, _ in fr.RECONCILE_CASES] +
                      [fr.nested_same_id_frame(**dict(kw, small_id=9)) for kw, _ in fr.RECONCILE_CASES])
    det = make_detector(1456, 1088, len(frames))
    out, counts = det.detect_batch(frames)
    n = len(fr.RECONCILE_CASES)
    for b, (kw, want) in enumerate(fr.RECONCILE_CASES):
        got = out[b, :counts[b]]
        assert_same_detections(got, oracle.detect(frames[b]))
        assert got["id"].tolist() == [7]
        if want is not None:
            assert (np.ptp(got[0]["p"][:, 0]) > 400) == (want == "big")
        both = out[n + b, :counts[n + b]]
        assert both["id"].tolist() == [7, 9]                     # distinct ids: both survive
        assert_same_detections(both, oracle.detect(frames[n + b]))
    det.close()


@pytest.mark.parametrize("f,W,H", [(1.0, 640, 480), (3.0, 640, 480), (3.0, 1000, 750), (1.0, 322, 246), (4.0, 1280, 720)])
def test_other_decimation_factors(oracle, f, W, H):
    """quad_decimate 1, 3, 4 (cb_set_params; the reference leaves it at 2): every stage against the oracle.  Factor 1 works on
    the frame itself (corners are not rescaled) and needs a context created with twice the frame size."""
    frames, _ = synth.render_batch(W, H, 2, 3, seed=7, edge_px=(60, 110) if W < 1000 else (80, 200))
    det = make_detector(2 * W, 2 * H, 2) if f == 1.0 else make_detector(W, H, 2)
    det.set_params(quad_decimate=f)
    prm = oracle.default_params(quad_decimate=f)
    thr = det.threshold(frames)
    lab, sz = det.labels(frames)
    q, qc, _ = det.quads(frames)
    out, counts = det.detect_batch(frames)
    for b in range(2):
        ref, taps = oracle.detect(frames[b], prm, taps=True)
        assert thr[b].shape == taps["thresh"].shape and (thr[b] == taps["thresh"]).all()
        assert (lab[b] == taps["labels"]).all() and (sz[b] == taps["comp_size"]).all()
        assert qc[b] == taps["nquads"]
        gq, oq = canon_quads(q[b, :qc[b]]), canon_quads(taps["quads"]["p"])
        assert np.abs(np.array(gq) - np.array(oq)).max() < 1e-4 if gq else True
        assert_same_detections(out[b, :counts[b]], ref)
        assert counts[b] >= (2 if W >= 640 else 1)
    det.close()


def test_decimation_argument_rules():
    from chalkydri_b200.capi import ChalkydriError, CB_ERR_UNSUPPORTED, CB_ERR_ARG
    det = make_detector(640, 480, 1)
    for bad in (1.5, 0.5, 0.0, 17.0):
        with pytest.raises(ChalkydriError) as e:
            det.set_params(quad_decimate=bad)
        assert e.value.code == CB_ERR_UNSUPPORTED
    det.set_params(quad_decimate=1.0)
    with pytest.raises(ChalkydriError) as e:                          # the context's buffers hold a 320x240 working image
        det.detect_batch(np.zeros((1, 480, 640), np.uint8))
    assert e.value.code == CB_ERR_ARG and "twice the frame size" in str(e.value)
    out, counts = det.detect_batch(np.full((1, 240, 320), 128, np.uint8))
    assert counts[0] == 0
    det.close()


def test_quads_against_opencv_apriltag_port():
    """INDEPENDENT PIN, no oracle in the loop: the CUDA detector at quad_decimate = 1 / refine_edges = 0 against the corners
    OpenCV's own port of the UMich quad detector (aruco CORNER_REFINE_APRILTAG) found on the same frames, committed in
    tests/golden/detector_cv2_pin.npz (measured <= 0.05 px, asserted < 0.08 px); and the default path within 0.3 px."""
    pin = np.load(os.path.join(GOLD, "detector_cv2_pin.npz"))
    det1 = make_detector(2560, 1440, 1)
    det1.set_params(quad_decimate=1.0, refine_edges=0)
    det2 = make_detector(1280, 720, 1)

    def quad_dist(c, p):
        return min(np.abs(np.roll(cc, s, 0) - p).max() for cc in (c, c[::-1]) for s in range(4))

    for name in ("c1", "s2", "s3"):
        im, _ = synth.render_frame(1280, 720, 4, seed=int(pin[f"{name}_seed"]), edge_px=(60, 150))
        for det, tol in ((det1, 0.08), (det2, 0.3)):
            out, counts = det.detect_batch(im[None])
            d = out[0, :counts[0]]
            assert d["id"].tolist() == pin[f"{name}_ids"].tolist()
            for rec, c in zip(d, pin[f"{name}_corners"]):
                assert quad_dist(c, rec["p"]) < tol
    det1.close(); det2.close()


def test_yuv420_buffers(oracle):
    """NV12 / I420 camera buffers (gst_to_cu.rs:152-188): the Y plane is the gray image, frames are width*height*3/2 apart and
    the chroma bytes behind each Y plane are never looked at.  Blocking call, chunk-pipelined call (batch >= 64) and the
    streaming form with a frame stride."""
    from tests import frames as fr
    from chalkydri_b200 import capi
    grays, _ = synth.render_batch(640, 480, 70, 2, seed=5, unique=5, edge_px=(50, 110))
    yuv = fr.yuv420_from_gray(grays, seed=1)
    det = make_detector(640, 480, 16)
    out, counts = det.detect_yuv420_batch(yuv[:3])
    for b in range(3):
        assert_same_detections(out[b, :counts[b]], oracle.detect(grays[b]))
    ref_out, ref_counts = det.detect_batch(grays)
    out, counts = det.detect_yuv420_batch(yuv)                      # 70 frames through a 16-frame context: pipelined chunks
    assert counts.tolist() == ref_counts.tolist() and counts.sum() >= 100
    for b in range(70):
        assert out[b, :counts[b]].tobytes() == ref_out[b, :counts[b]].tobytes()
    # streaming form on the same buffers: frame_stride = 1.5 * W * H
    h = capi.pinned_array(yuv[:16].shape, np.uint8)
    h[...] = yuv[:16]
    det._check(det._L.cb_detect_gray_submit(det.ctx, capi.ptr(h), 640, 480, 640, 640 * 480 * 3 // 2, 16))
    o2 = np.zeros((16, det.max_dets), out.dtype); c2 = np.zeros(16, np.int32)
    det._check(det._L.cb_detect_gray_collect(det.ctx, capi.ptr(o2), capi.ptr(c2)))
    assert c2.tolist() == ref_counts[:16].tolist()
    for b in range(16):
        assert o2[b, :c2[b]].tobytes() == ref_out[b, :c2[b]].tobytes()
    capi.free_pinned(h)
    det.close()


def test_pool_shards_frames_over_contexts_without_a_collective(oracle):
    """cb_pool_detect_gray (single process): one context + one host thread per GPU, contiguous shares, every share's
    lists written into its slice of the caller's one array.  A box with one GPU lists it twice -- two contexts, two threads, the
    same code path as two GPUs.  37 frames in batches of 8: ragged shares (19 + 18) and ragged last batches."""
    from chalkydri_b200 import capi
    from chalkydri_b200.pool import DetectorPool
    frames, _ = synth.render_batch(640, 480, 37, 2, seed=61, unique=6, edge_px=(50, 110))
    h = capi.pinned_array(frames.shape, np.uint8)
    h[...] = frames
    det = make_detector(640, 480, 8, 16)
    want, wc = det.detect_batch(frames)
    det.close()
    for devices in ([0], [0, 0], [0, 0, 0]):
        pool = DetectorPool(devices, 640, 480, 8, 16)
        assert len(pool) == len(devices)
        for _ in range(2):                                     # a second call reuses the contexts
            out, counts = pool.detect_batch(h)
            assert counts.tolist() == wc.tolist() and counts.sum() >= 37
            for b in range(37):
                assert out[b, :counts[b]].tobytes() == want[b, :wc[b]].tobytes()
                assert (out[b, :counts[b]]["frame"] == b).all()
        t = pool.timing()
        assert t["n_devices"] == len(devices) and t["wall_ms"] > 0
        pool.close()
    assert_same_detections(want[5, :wc[5]], oracle.detect(frames[5]))
    capi.free_pinned(h)


def test_pool_reports_errors_per_gpu():
    from chalkydri_b200.capi import ChalkydriError
    from chalkydri_b200.pool import DetectorPool
    pool = DetectorPool([0, 0], 640, 480, 4, 16)
    with pytest.raises(ChalkydriError) as e:
        pool.detect_batch(np.zeros((6, 600, 800), np.uint8))      # larger than the contexts
    assert "GPU 0" in str(e.value)
    out, counts = pool.detect_batch(np.full((6, 480, 640), 128, np.uint8))      # the queues were drained: the pool still works
    assert counts.tolist() == [0] * 6
    pool.close()
    with pytest.raises(ChalkydriError):
        DetectorPool([99], 640, 480, 4, 16)
