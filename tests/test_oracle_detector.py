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
    prm.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_SUBPIX
    det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11), prm)
    corners, ids, _ = det.detectMarkers(im)
    assert sorted(ids.ravel().tolist()) == sorted(dets["id"].tolist())
    for c, i in zip(corners, ids.ravel()):
        d = dets[dets["id"].tolist().index(int(i))]
        c = c.reshape(4, 2) + 0.5          # OpenCV reports pixel-centre coordinates
        # same quadrilateral up to cyclic order / direction
        best = min(np.abs(np.roll(cc, s, 0) - d["p"]).max() for cc in (c, c[::-1]) for s in range(4))
        assert best < 1.0, (int(i), best)


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


def test_batch_threads_match_single(oracle):
    frames, _ = synth.render_batch(640, 480, 4, 2, seed=3, edge_px=(50, 90))
    out1, c1 = oracle.detect_batch(frames, nthreads=1)
    out4, c4 = oracle.detect_batch(frames, nthreads=4)
    assert (c1 == c4).all() and c1.sum() >= 6
    assert out1.tobytes() == out4.tobytes()
