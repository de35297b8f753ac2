"""CPU tests of the detector oracle (oracle/apriltag_oracle.cpp).

The reference holds no golden vectors for this path (parity unpinned), so the oracle is pinned on
 (1) known-answer entries of AprilTag-3's tag36h11 table, family invariants (587 codes, min distance 11 over rotations),
 (2) ground truth of the synthetic generator (ids exact, corners < 0.5 px),
 (3) cv2.aruco (independent implementation): same ids, corners within 1 px,
 (4) the committed fixture tests/golden/detector_c1.npz (regression pin of the oracle itself).
"""
import os

import numpy as np
import pytest

from chalkydri_b200 import synth
from chalkydri_b200.tagfamily import TAG36H11_CODES

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rot90(w):
    return ((w << 9) | (w >> 27)) & ((1 << 36) - 1)


def test_code_table_known_answers(oracle):
    codes = oracle.tag36h11_codes()
    assert len(codes) == 587
    assert [int(c) for c in codes[:3]] == [0x0000000d7e00984b, 0x0000000dda664ca7, 0x0000000dc4a1c821]
    assert [int(c) for c in codes] == TAG36H11_CODES


def test_code_table_min_distance_11(oracle):
    codes = [int(c) for c in oracle.tag36h11_codes()]
    allrot = []
    for c in codes:
        r = c
        for _ in range(4):
            allrot.append(r)
            r = rot90(r)
    a = np.array(allrot, np.uint64)
    base = np.array(codes, np.uint64)
    x = base[:, None] ^ a[None, :]
    pc = np.zeros(x.shape, np.int32)
    for k in range(36):
        pc += ((x >> np.uint64(k)) & np.uint64(1)).astype(np.int32)
    for i in range(587):
        pc[i, 4 * i] = 99      # identity
    assert pc.min() == 11


def match_truth(dets, truth, tol):
    assert sorted(dets["id"].tolist()) == sorted(truth["ids"].tolist())
    for d in dets:
        k = truth["ids"].tolist().index(int(d["id"]))
        err = np.abs(d["p"] - truth["corners"][k]).max()
        assert err < tol, (int(d["id"]), err)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c1_against_ground_truth(oracle, seed):
    im, truth = synth.render_frame(1280, 720, 4, seed=seed, edge_px=(60, 150))
    dets = oracle.detect(im)
    assert (dets["hamming"] == 0).all()
    match_truth(dets, truth, 0.5)


def test_c1_against_cv2_aruco(oracle):
    cv2 = pytest.importorskip("cv2")
    im, truth = synth.render_frame(1280, 720, 4, seed=1, edge_px=(60, 150))
    dets = oracle.detect(im)
    prm = cv2.aruco.DetectorParameters()
    prm.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_APRILTAG       # OpenCV's own port of the UMich quad detector
    det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11), prm)
    corners, ids, _ = det.detectMarkers(im)
    assert sorted(ids.ravel().tolist()) == sorted(dets["id"].tolist())
    for c, i in zip(corners, ids.ravel()):
        d = dets[dets["id"].tolist().index(int(i))]
        c = c.reshape(4, 2)                # the AprilTag path reports AprilTag's pixel convention (no half-pixel shift)
        # same quadrilateral up to cyclic order / direction; 0.3 px covers refine_edges, which OpenCV's port does not run
        # (test_quads_against_opencv_apriltag_port compares like with like at 0.08 px)
        best = min(np.abs(np.roll(cc, s, 0) - d["p"]).max() for cc in (c, c[::-1]) for s in range(4))
        assert best < 0.3, (int(i), best)


def test_threshold_values_and_remainder(oracle):
    rng = np.random.default_rng(0)
    im = rng.integers(0, 256, (37, 50), dtype=np.uint8)      # decimates to 25 x 19: partial tiles on both edges
    im[:16, :16] = 100                                         # flat block -> 127
    thr = oracle.threshold(im)
    assert thr.shape == (19, 25)
    assert set(np.unique(thr)) <= {0, 127, 255}
    assert (thr[:4, :4] == 127).all()
    assert (thr[:, 24] != 127).all() and (thr[16:, :] != 127).all()    # remainder pixels are never 127


def test_empty_frame_has_no_detections(oracle):
    im = np.full((720, 1280), 128, np.uint8)
    dets, taps = oracle.detect(im, taps=True)
    assert len(dets) == 0 and taps["npoints"] == 0 and (taps["thresh"] == 127).all()


def test_partition_label_is_min_index(oracle):
    im, _ = synth.render_frame(320, 240, 1, seed=5, edge_px=(60, 80))
    _, taps = oracle.detect(im, taps=True)
    lab = taps["labels"].ravel()
    idx = np.arange(lab.size)
    assert (lab <= idx).all()
    assert (lab[lab] == lab).all()
    # sizes agree with the label histogram
    cnt = np.bincount(lab, minlength=lab.size)
    assert (taps["comp_size"].ravel() == cnt[lab]).all()


def test_golden_fixture_c1(oracle):
    """Regression pin: tests/golden/detector_c1.npz was produced by tests/golden/make_golden.py."""
    g = np.load(os.path.join(GOLD, "detector_c1.npz"))
    im, _ = synth.render_frame(1280, 720, 4, seed=int(g["seed"]), edge_px=(60, 150))
    assert (im == g["frame"]).all(), "generator drifted"
    dets, taps = oracle.detect(im, taps=True)
    assert dets["id"].tolist() == g["ids"].tolist()
    assert dets["hamming"].tolist() == g["hamming"].tolist()
    assert np.abs(dets["p"] - g["corners"]).max() < 1e-9
    assert (np.packbits(taps["thresh"] == 255) == g["thresh_white_bits"]).all()
    assert (np.packbits(taps["thresh"] == 0) == g["thresh_black_bits"]).all()
    assert int(taps["npoints"]) == int(g["npoints"]) and int(taps["nquads"]) == int(g["nquads"])
    assert np.abs(taps["quads"]["p"] - g["quads"]).max() < 1e-6         # candidate quads (float, decimated coordinates)
    assert np.allclose(dets["decision_margin"], g["margin"], rtol=1e-6) and np.allclose(dets["H"], g["H"], rtol=1e-9, atol=1e-9)


def test_batch_threads_match_single(oracle):
    frames, _ = synth.render_batch(640, 480, 4, 2, seed=3, edge_px=(50, 90))
    out1, c1 = oracle.detect_batch(frames, nthreads=1)
    out4, c4 = oracle.detect_batch(frames, nthreads=4)
    assert (c1 == c4).all() and c1.sum() >= 6
    assert out1.tobytes() == out4.tobytes()


def threshold_np(im, min_white_black_diff=5):
    """A1 + A2 written a second time, vectorised: point decimation by 2, 4x4 tile extrema, 3x3 tile-neighbourhood
    dilate / erode, per-pixel binarisation, and upstream's remainder rule (last full tile, never 127)."""
    d = im[::2, ::2].astype(np.int32)
    h, w = d.shape
    th, tw = h // 4, w // 4
    t = d[:th * 4, :tw * 4].reshape(th, 4, tw, 4)
    tmax, tmin = t.max((1, 3)), t.min((1, 3))
    pmax = np.pad(tmax, 1, constant_values=-1)
    pmin = np.pad(tmin, 1, constant_values=1 << 20)
    nmax = np.max([pmax[1 + dy:1 + dy + th, 1 + dx:1 + dx + tw] for dy in (-1, 0, 1) for dx in (-1, 0, 1)], axis=0)
    nmin = np.min([pmin[1 + dy:1 + dy + th, 1 + dx:1 + dx + tw] for dy in (-1, 0, 1) for dx in (-1, 0, 1)], axis=0)
    ty = np.minimum(np.arange(h) // 4, th - 1)[:, None]
    tx = np.minimum(np.arange(w) // 4, tw - 1)[None, :]
    mx, mn = nmax[ty, tx], nmin[ty, tx]
    out = np.where(d > mn + (mx - mn) // 2, 255, 0).astype(np.uint8)
    full = (np.arange(h) < th * 4)[:, None] & (np.arange(w) < tw * 4)[None, :]
    out[full & (mx - mn < min_white_black_diff)] = 127
    return out


@pytest.mark.parametrize("W,H,seed", [(640, 480, 1), (322, 246, 2), (100, 74, 3)])
def test_threshold_against_numpy_restatement(oracle, W, H, seed):
    """bit-exact stage: frames with tags (structure) and a noise frame with partial tiles on both edges"""
    im, _ = synth.render_frame(W, H, 2, seed=seed, edge_px=(30, 60)) if W > 200 else (np.random.default_rng(seed).integers(0, 256, (H, W), dtype=np.uint8), None)
    thr = oracle.threshold(im)
    assert (thr == threshold_np(im)).all()


@pytest.mark.parametrize("seed", [4, 5])
def test_partition_against_scipy_graph_components(oracle, seed):
    """A3 as a graph problem: the links upstream's scan makes (x in 1..w-2: left, up; white also up-left / up-right), handed to
    scipy.sparse.csgraph.connected_components; the oracle's min-index labels must describe the same partition -- including
    upstream's behaviour at the unscanned border columns."""
    sparse = pytest.importorskip("scipy.sparse")
    from scipy.sparse.csgraph import connected_components
    im, _ = synth.render_frame(320, 240, 2, seed=seed, edge_px=(40, 80))
    _, taps = oracle.detect(im, taps=True)
    t = taps["thresh"]
    h, w = t.shape
    idx = np.arange(h * w).reshape(h, w)
    src, dst = [], []

    def link(a, b, cond):
        src.append(idx[a][cond]); dst.append(idx[b][cond])

    xs = slice(1, w - 1)
    good = t != 127
    link((slice(None), xs), (slice(None), slice(0, w - 2)), (t[:, xs] == t[:, 0:w - 2]) & good[:, xs])               # left
    link((slice(1, None), xs), (slice(0, h - 1), xs), (t[1:, xs] == t[:-1, xs]) & good[1:, xs])                          # up
    white = t == 255
    link((slice(1, None), xs), (slice(0, h - 1), slice(0, w - 2)), white[1:, xs] & white[:-1, 0:w - 2])                  # up-left
    # up-right: upstream skips it when (x, y-1) == (x+1, y-1), counting on the left-link of pixel (x+1, y-1) -- which does not
    # exist for x+1 = w-1 (never scanned).  So this one guard is part of the semantics: last-column pixels stay on their own
    # unless the pixel to their left differs.  (The guards on the up and up-left links are redundant everywhere.)
    link((slice(1, None), xs), (slice(0, h - 1), slice(2, w)), white[1:, xs] & white[:-1, 2:w] & (t[:-1, xs] != t[:-1, 2:w]))
    s, d = np.concatenate(src), np.concatenate(dst)
    g = sparse.coo_matrix((np.ones(len(s), np.int8), (s, d)), shape=(h * w, h * w))
    ncomp, comp = connected_components(g, directed=False)
    first = np.full(ncomp, h * w, np.int64)
    np.minimum.at(first, comp, np.arange(h * w))
    want = first[comp]
    lab = taps["labels"].ravel().astype(np.int64)
    assert (lab[good.ravel()] == want[good.ravel()]).all()
    assert (taps["comp_size"].ravel()[good.ravel()] == np.bincount(comp, minlength=ncomp)[comp][good.ravel()]).all()


@pytest.mark.parametrize("seed", [4, 5])
def test_gradient_cluster_points_against_numpy_count(oracle, seed):
    """A4 counted a second way, vectorised: boundary points for the probes (1,0), (0,1), (-1,1), (1,1) between components of at
    least 25 pixels with v0 + v1 == 255, the (-1,1) probe suppressed when the previous pixel's (1,1) probe connected
    (`connected_last`); clusters = distinct unordered component pairs."""
    im, _ = synth.render_frame(320, 240, 2, seed=seed, edge_px=(40, 80))
    _, taps = oracle.detect(im, taps=True)
    t = taps["thresh"].astype(np.int32)
    h, w = t.shape
    big = (taps["comp_size"] >= 25) & (t != 127)

    def conn(dx, dy):
        c = np.zeros((h, w), bool)
        ys, xs = slice(1, h - 1), slice(1, w - 1)
        nb = (slice(1 + dy, h - 1 + dy), slice(1 + dx, w - 1 + dx))
        c[ys, xs] = big[ys, xs] & big[nb] & (t[ys, xs] + t[nb] == 255)
        return c

    c10, c01, cm11, c11 = conn(1, 0), conn(0, 1), conn(-1, 1), conn(1, 1)
    last = np.zeros((h, w), bool)
    last[:, 1:] = c11[:, :-1]
    probes = ((c10, (1, 0)), (c01, (0, 1)), (cm11 & ~last, (-1, 1)), (c11, (1, 1)))
    assert sum(int(c.sum()) for c, _ in probes) == int(taps["npoints"]) > 1000
    lab = taps["labels"].astype(np.int64)
    keys = []
    for c, (dx, dy) in probes:
        ys, xs = np.nonzero(c)
        a, b = lab[ys, xs], lab[ys + dy, xs + dx]
        keys.append(np.minimum(a, b) * (h * w) + np.maximum(a, b))
    assert len(np.unique(np.concatenate(keys))) == int(taps["nclusters"])


def test_detection_record_conventions(oracle):
    """apriltag_detection_t as upstream fills it: p[i] = H (-1,1), (1,1), (1,-1), (-1,-1) and c = H (0,0) (homography_project on
    the rotated homography), counter-clockwise corner winding in image coordinates (y down), hamming 0 and a positive decision
    margin on clean renderings, records sorted by id."""
    im, truth = synth.render_frame(1280, 720, 4, seed=1, edge_px=(60, 150))
    dets = oracle.detect(im)
    assert sorted(dets["id"].tolist()) == sorted(truth["ids"]) == dets["id"].tolist()
    for d in dets:
        H = d["H"].reshape(3, 3)

        def project(x, y):
            v = H @ np.array([x, y, 1.0])
            return v[:2] / v[2]

        pts = np.array([project(-1, 1), project(1, 1), project(1, -1), project(-1, -1)])
        assert np.abs(pts - d["p"]).max() < 1e-9 and np.abs(project(0, 0) - d["c"]).max() < 1e-9
        x, y = d["p"][:, 0], d["p"][:, 1]
        area2 = float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))
        assert area2 < 0                                   # upstream's winding: negative shoelace sum with y pointing down
        assert d["hamming"] == 0 and d["decision_margin"] > 20


# ---- round 2: error-corrected decode, family bits, reconcile, decimation factors, and an independent quad pin ----------

@pytest.mark.parametrize("bits", [0, 1, 2, 3])
def test_hamming_decode_and_family_bits(oracle, bits):
    """quick_decode with add_family_bits(family, bits) (crates/apriltags/src/lib.rs:229-230 bits 3, :279-282 bits 1): a tag with k
    inverted data bits decodes with hamming k when k <= bits and is rejected otherwise (tag36h11's minimum distance 11 keeps the
    radius-3 balls of all 587 x 4 rotated codes disjoint, so the answer is unique)."""
    from tests import frames as fr
    for k in range(5):
        im, truth = synth.render_frame(1280, 720, 4, seed=31, edge_px=(70, 150), bit_errors=k)
        dets = oracle.detect(im, oracle.default_params(bits_corrected=bits))
        if k <= bits:
            assert sorted(dets["id"].tolist()) == sorted(truth["ids"].tolist()) and (dets["hamming"] == k).all()
            match_truth(dets, truth, 0.5)
        else:
            assert len(dets) == 0
    im, truth = fr.bit_error_frame()
    dets = oracle.detect(im, oracle.default_params(bits_corrected=bits))
    want = {int(i): int(h) for i, h in zip(truth["ids"], truth["hamming"]) if h <= bits}
    assert {int(d["id"]): int(d["hamming"]) for d in dets} == want


def test_family_bits_out_of_range(oracle):
    im = np.full((64, 64), 128, np.uint8)
    for bits in (-1, 4):
        with pytest.raises(RuntimeError):
            oracle.detect(im, oracle.default_params(bits_corrected=bits))


def test_bit_error_golden_fixture_and_cv2_decode(oracle):
    """tests/golden/detector_biterr.npz (regression pin) and detector_cv2_pin.npz (OpenCV's decoder with maxCorrectionBits = k
    accepts exactly the tags bits_corrected = k accepts)."""
    g = np.load(os.path.join(GOLD, "detector_biterr.npz"))
    pin = np.load(os.path.join(GOLD, "detector_cv2_pin.npz"))
    from tests import frames as fr
    im, _ = fr.bit_error_frame()
    assert (im == g["frame"]).all(), "generator drifted"
    for bits in range(4):
        dets = oracle.detect(im, oracle.default_params(bits_corrected=bits))
        assert dets["id"].tolist() == g[f"ids_b{bits}"].tolist() and dets["hamming"].tolist() == g[f"hamming_b{bits}"].tolist()
        assert np.abs(dets["p"] - g[f"corners_b{bits}"]).max() < 1e-9 if len(dets) else True
        assert sorted(dets["id"].tolist()) == pin[f"biterr_ids_k{bits}"].tolist()
    assert g["hamming_b3"].max() == 3


def test_reconcile_overlapping_duplicates(oracle):
    """Two detections of one id with overlapping polygons: upstream keeps the lower hamming, then the higher decision margin."""
    from tests import frames as fr
    for kw, want in fr.RECONCILE_CASES:
        both = oracle.detect(fr.nested_same_id_frame(**dict(kw, small_id=9)))          # distinct ids: nothing to reconcile
        assert both["id"].tolist() == [7, 9]
        big, small = both[0], both[1]
        dets = oracle.detect(fr.nested_same_id_frame(**kw))
        assert dets["id"].tolist() == [7], kw
        width = float(np.ptp(dets[0]["p"][:, 0]))
        if want is None:              # equal hamming, margins of the same order: whichever survives must be one of the two, intact
            want = "big" if width > 400 else "small"
        assert (width > 400) == (want == "big"), (kw, width)
        ref = big if want == "big" else small
        assert dets[0]["hamming"] == ref["hamming"] and np.abs(dets[0]["p"] - ref["p"]).max() < 1e-9
        if want == "big":
            assert dets[0]["decision_margin"] == big["decision_margin"]


@pytest.mark.parametrize("f", [1.0, 3.0])
def test_other_decimation_factors(oracle, f):
    """quad_decimate 1 (detector works on the frame itself, corners not rescaled) and 3 (point sampling every third pixel,
    min_tag_width = max(3, 8 / f))."""
    im, truth = synth.render_frame(640, 480, 3, seed=7, edge_px=(60, 110))
    dets, taps = oracle.detect(im, oracle.default_params(quad_decimate=f), taps=True)
    assert taps["thresh"].shape == ((480, 640) if f == 1.0 else (160, 214))
    match_truth(dets, truth, 0.5)
    with pytest.raises(RuntimeError):
        oracle.detect(im, oracle.default_params(quad_decimate=1.5))


@pytest.mark.parametrize("name", ["c1", "s2", "s3"])
def test_quads_against_opencv_apriltag_port(oracle, name):
    """INDEPENDENT PIN of rows A2-A5 + A7.  OpenCV's aruco module contains its own port of the UMich quad detector
    (CORNER_REFINE_APRILTAG); on the same frame, un-decimated and without refine_edges, this restatement must find the same
    quadrilaterals: measured <= 0.05 px, asserted < 0.08 px (the previous round compared with CORNER_REFINE_SUBPIX at 1 px).
    Checked against the committed OpenCV output, and against OpenCV itself when it is importable."""
    pin = np.load(os.path.join(GOLD, "detector_cv2_pin.npz"))
    im, _ = synth.render_frame(1280, 720, 4, seed=int(pin[f"{name}_seed"]), edge_px=(60, 150))
    dets = oracle.detect(im, oracle.default_params(quad_decimate=1.0, refine_edges=0))
    assert dets["id"].tolist() == pin[f"{name}_ids"].tolist()

    def quad_dist(c, p):
        return min(np.abs(np.roll(cc, s, 0) - p).max() for cc in (c, c[::-1]) for s in range(4))

    for d, c in zip(dets, pin[f"{name}_corners"]):
        assert quad_dist(c, d["p"]) < 0.08
    # the full default path (decimate 2 + refine_edges on the full-resolution frame) lands within 0.3 px of the same corners
    full = oracle.detect(im)
    for d, c in zip(full, pin[f"{name}_corners"]):
        assert quad_dist(c, d["p"]) < 0.3
    try:
        import cv2
    except ImportError:
        return
    prm = cv2.aruco.DetectorParameters()
    prm.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_APRILTAG
    corners, ids, _ = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11), prm).detectMarkers(im)
    for c, i in zip(corners, ids.ravel()):
        assert quad_dist(c.reshape(4, 2), dets[dets["id"].tolist().index(int(i))]["p"]) < 0.08
