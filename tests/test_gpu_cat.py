"""GPU parity tests of the CAT stages vs the oracle restatement of crates/chalkydri-apriltags (all bit-exact)."""
import numpy as np
import pytest

from chalkydri_b200 import synth

pytestmark = pytest.mark.gpu


def rgb_frame(w, h, seed, tags=2):
    gray, _ = synth.render_frame(w, h, tags, seed=seed, edge_px=(40, 90))
    return synth.gray_to_rgb(gray, seed=seed)


@pytest.mark.parametrize("w,h,seed", [(320, 240, 1), (703, 905, 2), (101, 67, 3)])
def test_cat_stages(oracle, w, h, seed):
    from chalkydri_b200.cat import CatDetector
    rgb = rgb_frame(w, h, seed)
    d = CatDetector(w, h, ())
    d.calc_otsu(rgb)
    ref = oracle.cat_calc_otsu(rgb)
    assert (d.buf == ref).all(), "calc_otsu colour map differs"
    d.detect_corners()
    rxy, rn = oracle.cat_detect_corners(ref)
    assert len(d.points) == rn and (d.points == rxy).all()
    if rn > 3000:                       # keep the O(P^2) edge test bounded
        d.points = d.points[:3000]; rxy = rxy[:3000]
    d.check_edges()
    rl, rln = oracle.cat_check_edges(ref, rxy)
    assert len(d.lines) == rln and (d.lines == rl).all()
    uf = d.connected_components()
    lab, sz = oracle.cat_connected_components(ref)
    assert (uf.parent == lab).all() and (uf.cluster_sizes == sz).all()
    d.thresh(rgb)
    assert (d.buf == oracle.cat_thresh(rgb)).all()
    d.close()


def test_process_frame_and_assert(oracle):
    from chalkydri_b200.cat import CatDetector
    rgb = rgb_frame(320, 240, 7)
    d = CatDetector(320, 240, ())
    d.process_frame(rgb.reshape(-1))
    ref = oracle.cat_calc_otsu(rgb)
    rxy, _ = oracle.cat_detect_corners(ref)
    rl, _ = oracle.cat_check_edges(ref, rxy)
    assert (d.points == rxy).all() and (d.lines == rl).all()
    with pytest.raises(AssertionError):
        d.process_frame(rgb.reshape(-1)[:-3])       # wrong length panics in the reference (lib.rs:267)
    d.close()


def test_grayscale_exhaustive_sample(oracle):
    """utils.rs:43 on a dense RGB sample: the fused-multiply-add form is reproduced bit for bit."""
    from chalkydri_b200.cat import CatDetector
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    d = CatDetector(64, 64, ())
    d.thresh(rgb)
    g = np.array([oracle.cat_grayscale(int(r), int(gg), int(b)) for r, gg, b in rgb.reshape(-1, 3)]).reshape(64, 64)
    want = np.where(g < 60, 0, np.where(g > 160, 1, 2))
    assert (d.buf == want).all()
    d.close()


def test_golden_fixture_cat_on_gpu():
    """The CUDA CAT stages against the committed vectors tests/golden/cat_96x72.npz (no oracle in the loop), bit for bit."""
    import os
    from chalkydri_b200.cat import CatDetector
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cat_96x72.npz"))
    d = CatDetector(96, 72, ())
    d.calc_otsu(g["rgb"])
    assert (d.buf == g["otsu"]).all()
    d.detect_corners()
    assert len(d.points) == len(g["corners"]) and (d.points == g["corners"]).all()
    d.check_edges()
    assert len(d.lines) == len(g["lines"]) and (d.lines == g["lines"]).all()
    uf = d.connected_components()
    assert (uf.parent == g["labels"]).all() and (uf.cluster_sizes == g["sizes"]).all()
    d.thresh(g["rgb"])
    assert (d.buf == g["thresh"]).all()
    d.close()


def test_last_column_reads_wrap_to_the_next_row_on_gpu(oracle):
    """x = w-3 (inside the reference's loop, lib.rs:293): px(x+3, y-+3) is column 0 of the next row (utils.rs:27-29)."""
    from chalkydri_b200.cat import CatDetector
    from tests.test_oracle_cat import corners_py
    w, h = 40, 30
    base = np.full((h, w), 1, np.uint8)
    base[10:20, w - 3:] = 0
    with_corner = base.copy()
    with_corner[14, 0] = 0                             # px(w-3+3, 10+3) = row 14, column 0: the wrapped far sample turns black
    ring_other = with_corner.copy()
    ring_other[8, 0] = 2                               # px(w-3+3, 10-3) = row 8, column 0 is Other: ring not all good
    d = CatDetector(w, h, ())
    found = []
    for c in (base, with_corner, ring_other):
        d.buf[...] = c
        d.detect_corners()
        rxy, rn = oracle.cat_detect_corners(c)
        got = [tuple(p) for p in d.points.tolist()]
        assert got == [tuple(p) for p in rxy[:rn].tolist()] == corners_py(c)
        found.append((w - 3, 10) in got)
    assert found == [False, True, False]
    d.close()


def high_contrast_rgb(w, h, seed, tags):
    """tags whose black cells fall below CAT's fixed 60 and whose white cells above its 160 (lib.rs:319-334)"""
    gray, truth = synth.render_frame(w, h, tags, seed=seed, edge_px=(60, 140))
    g = np.clip((gray.astype(np.float32) - 128.0) * 3.0 + 128.0, 0, 255).astype(np.uint8)
    return synth.gray_to_rgb(g, seed=seed), truth


@pytest.mark.parametrize("w,h,seed,use_otsu", [(640, 480, 7, False), (960, 540, 3, False), (703, 905, 5, False), (640, 480, 7, True)])
def test_cat_decode_matches_the_oracle(oracle, w, h, seed, use_otsu):
    """CAT's decode intent (book/src/maintenance/apriltags.md:58-60, lib.rs:551-613) in one call, cb_cat_detect_tags: CAT's own ternary
    map, then the C library's stages.  Oracle: CAT gray (utils.rs:43) and colour map from the CAT restatement, mapped to 0 / 255 /
    127 and handed to upstream's pipeline in place of its threshold() (orc_detect_with_map, quad_decimate = 1)."""
    from chalkydri_b200.cat import CatDetector
    rgb, truth = high_contrast_rgb(w, h, seed, 4)
    d = CatDetector(w, h, ())
    got = d.detect_tags(rgb, use_otsu=use_otsu)
    col = oracle.cat_calc_otsu(rgb) if use_otsu else oracle.cat_thresh(rgb)
    # CAT's gray plane (utils.rs:43: two fused multiply-adds in f32, truncating cast) vectorised; a sample of it is checked against
    # the oracle's scalar restatement (the device plane itself is covered by test_rgb_to_gray_every_colour_bit_exact)
    flat = rgb.reshape(-1, 3).astype(np.float32)
    v = np.float32(0.33)
    cgray = np.floor((flat[:, 0].astype(np.float64) * float(v) + (flat[:, 1].astype(np.float64) * float(v) + (flat[:, 2] * v).astype(np.float64))).astype(np.float32)).astype(np.uint8)
    step = max(1, w * h // 4000)
    assert (cgray[::step] == np.array([oracle.cat_grayscale(int(r), int(g), int(b)) for r, g, b in rgb.reshape(-1, 3)[::step]], np.uint8)).all()
    tmap = np.array([0, 255, 127], np.uint8)[col]
    ref = oracle.detect_with_map(cgray.reshape(h, w), tmap, oracle.default_params(quad_decimate=1.0))
    assert got["id"].tolist() == ref["id"].tolist() and got["hamming"].tolist() == ref["hamming"].tolist()
    if len(ref):
        assert np.abs(got["p"] - ref["p"]).max() < 1e-3
    if not use_otsu:
        assert len(ref) >= 2 and set(ref["id"].tolist()) <= set(int(i) for i in truth["ids"])     # the path does decode the rendered tags
    # valid_tags filters the list like the constructor argument says
    if len(ref):
        keep = (int(ref["id"][0]),)
        d2 = CatDetector(w, h, keep)
        assert set(d2.detect_tags(rgb, use_otsu=use_otsu)["id"].tolist()) == set(keep)
        d2.close()
    d.close()
