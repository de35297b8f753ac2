"""CPU pins of the CAT restatement (oracle/cat_oracle.cpp; reference crates/chalkydri-apriltags/src/{lib,utils}.rs).

The reference has no test or fixture for CAT (PARITY UNPINNED).  Each stage of the C++ restatement is checked
here against a second, independently written restatement in numpy / plain Python on small seeded images; the statrs R-8
quantile is pinned against numpy's `median_unbiased` (Hyndman-Fan type 8)."""
import numpy as np
import pytest

BLACK, WHITE, OTHER = 0, 1, 2          # utils.rs:1-6


def gray_np(rgb):
    """utils.rs:43 with exact FMA emulation: the f64 value of a*b+c is exact for these operand widths, so one rounding to
    f32 is the fused result."""
    k = np.float64(np.float32(0.33))
    r, g, b = (rgb[..., i].astype(np.float64) for i in range(3))
    t0 = (b * k).astype(np.float32).astype(np.float64)             # (b as f32) * 0.33 : rounded product
    t1 = (g * k + t0).astype(np.float32).astype(np.float64)         # fma(g, 0.33, t0)
    t2 = (r * k + t1).astype(np.float32)                            # fma(r, 0.33, t1)
    return np.clip(np.floor(t2), 0, 255).astype(np.uint8)           # `as u8`: truncate, saturate


def test_grayscale_formula(oracle):
    rng = np.random.default_rng(0)
    tri = np.concatenate([rng.integers(0, 256, (20000, 3)), np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [100, 100, 100], [101, 101, 101]])])
    want = gray_np(tri.astype(np.uint8))
    got = np.array([oracle.cat_grayscale(int(r), int(g), int(b)) for r, g, b in tri], np.uint8)
    assert (got == want).all()
    assert oracle.cat_grayscale(255, 255, 255) == 252                # 3 * 255 * 0.33 = 252.45: never reaches 253


def rgb_image(w, h, seed):
    """blocks of flat colour (exercises the flat-window rule), gradients and noise (exercises the quartile rule)"""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.uint8)
    bs = 8
    for by in range(0, h, bs):
        for bx in range(0, w, bs):
            img[by:by + bs, bx:bx + bs] = rng.choice([10, 40, 100, 150, 200, 250])
    noisy = rng.random((h, w)) < 0.4
    img[noisy] = np.clip(img[noisy].astype(int) + rng.integers(-30, 31, (int(noisy.sum()), 3)), 0, 255).astype(np.uint8)
    return img


def statrs_quantile(sorted_vals, tau):
    """statrs 0.18.0 `Data::quantile` (the published R-8 code path): h = (n + 1/3) tau + 1/3, a + (h - floor h)(b - a), in
    exactly this order of double operations (Python floats are the same IEEE doubles)."""
    n = len(sorted_vals)
    h = (float(n) + 1.0 / 3.0) * tau + 1.0 / 3.0
    hf = int(h)
    if hf <= 0 or tau == 0.0:
        return float(sorted_vals[0])
    if hf >= n:
        return float(sorted_vals[-1])
    a, b = float(sorted_vals[hf - 1]), float(sorted_vals[hf])
    return a + (h - float(hf)) * (b - a)


def otsu_py(rgb):
    """calc_otsu (lib.rs:191-259), second restatement.  Returns the map with statrs' operation order, and the mask of pixels
    where numpy's own type-8 estimator (`median_unbiased`) gives a different class -- only possible where an interpolated
    quartile sits within an ulp of an integer, because `as u8` truncates."""
    h, w, _ = rgb.shape
    g = gray_np(rgb).astype(np.float64)
    out = np.empty((h, w), np.uint8)
    numpy_differs = np.zeros((h, w), bool)
    worst = 0.0

    def classify(p, uq, lq):
        return WHITE if p >= np.floor(uq) else (BLACK if p <= np.floor(lq) else OTHER)

    for y in range(h):
        for x in range(w):
            win = np.sort(g[max(y - 2, 0):min(y + 2, h - 1) + 1, max(x - 2, 0):min(x + 2, w - 1) + 1].reshape(-1))
            p = g[y, x]
            if y > 0 and x > 0 and win[-1] - win[0] < 5.0:
                m = np.median(win)
                out[y, x] = BLACK if m < 60.0 else (WHITE if m > 160.0 else OTHER)
                continue
            uq, lq = statrs_quantile(win, 0.75), statrs_quantile(win, 0.25)
            nuq, nlq = np.quantile(win, [0.75, 0.25], method="median_unbiased")
            worst = max(worst, abs(uq - nuq), abs(lq - nlq))
            out[y, x] = classify(p, uq, lq)
            numpy_differs[y, x] = classify(p, nuq, nlq) != out[y, x]
    return out, numpy_differs, worst


@pytest.mark.parametrize("w,h,seed", [(41, 29, 1), (64, 48, 2)])
def test_calc_otsu_against_numpy_quantiles(oracle, w, h, seed):
    rgb = rgb_image(w, h, seed)
    got = oracle.cat_calc_otsu(rgb)
    want, numpy_differs, worst = otsu_py(rgb)
    assert (got == want).all(), f"{int((got != want).sum())} pixels differ, first at {np.argwhere(got != want)[:3].tolist()}"
    assert worst < 1e-9                          # statrs' R-8 formula IS numpy's median_unbiased, up to rounding ...
    assert numpy_differs.mean() < 0.02           # ... which only matters where truncation meets an integer-valued quartile
    assert set(np.unique(got)) == {BLACK, WHITE, OTHER}


def test_thresh_fixed_levels(oracle):
    rgb = rgb_image(50, 30, 3)
    g = gray_np(rgb)
    want = np.where(g < 60, BLACK, np.where(g > 160, WHITE, OTHER)).astype(np.uint8)        # lib.rs:319-334
    assert (oracle.cat_thresh(rgb) == want).all()


def ternary_map(w, h, seed):
    """black squares on white with some Other speckle: plenty of FAST-like corners"""
    rng = np.random.default_rng(seed)
    c = np.full((h, w), WHITE, np.uint8)
    for _ in range(6):
        x0, y0 = int(rng.integers(4, w - 24)), int(rng.integers(4, h - 24))
        s = int(rng.integers(8, 20))
        c[y0:y0 + s, x0:x0 + s] = BLACK
    c[rng.random((h, w)) < 0.02] = OTHER
    return c


def corners_py(c):
    """detect_corners / process_pixel (lib.rs:291-309,345-400): x outer, y inner.  px() is an unchecked LINEAR index
    (utils.rs:27-29): in the last scanned column x = w-3 the (x+3, .) samples are column 0 of the next row; only samples beyond
    the w*h buffer (undefined behaviour in the reference) make a pixel unevaluable."""
    h, w = c.shape
    flat = c.reshape(-1)
    px = lambda x, y: y * w + x                           # noqa: E731
    out = []
    for x in range(3, w - 3 + 1):
        for y in range(3, h - 3 + 1):
            if px(x + 3, y + 3) >= w * h:
                continue
            if flat[px(x, y)] != BLACK:
                continue
            d = [flat[px(x - 1, y - 1)], flat[px(x + 1, y - 1)], flat[px(x - 1, y + 1)], flat[px(x + 1, y + 1)]]
            if not (sum(int(v == BLACK) for v in d) & 1):
                continue
            f = [flat[px(x + 3, y - 3)], flat[px(x + 3, y + 3)], flat[px(x - 3, y + 3)], flat[px(x - 3, y - 3)]]
            if all(v != OTHER for v in f) and (sum(int(v == BLACK) for v in f) & 1):
                out.append((x, y))
    return out


def test_last_column_reads_wrap_to_the_next_row(oracle):
    """x = w-3 is inside the reference's loop (3..=width-3): its right-hand ring samples are px(x+3, y-+3) = column 0 of rows
    y-2 / y+4.  A corner there is found exactly when those wrapped samples say so."""
    w, h = 40, 30
    c = np.full((h, w), WHITE, np.uint8)
    c[10:20, w - 3:] = BLACK                               # a black block whose top-left corner is pixel (w-3, 10)
    x, y = w - 3, 10
    as_list = lambda r: [tuple(p) for p in r[0][:r[1]].tolist()]      # noqa: E731
    # the four far samples are White, White (wrapped: c[y-2, 0], c[y+4, 0]), White, White: even parity, no corner
    assert (x, y) not in corners_py(c) and as_list(oracle.cat_detect_corners(c)) == corners_py(c)
    c2 = c.copy()
    c2[y + 4, 0] = BLACK                                   # the wrapped "bottom right" sample px(x+3, y+3) turns black: a corner
    assert (x, y) in corners_py(c2)
    assert as_list(oracle.cat_detect_corners(c2)) == corners_py(c2)
    c3 = c2.copy()
    c3[y - 2, 0] = OTHER                                   # the wrapped "top right" sample px(x+3, y-3) is Other: ring not all good
    assert (x, y) not in corners_py(c3) and as_list(oracle.cat_detect_corners(c3)) == corners_py(c3)


def test_detect_corners_order_and_predicate(oracle):
    c = ternary_map(96, 72, 4)
    xy, n = oracle.cat_detect_corners(c)
    want = corners_py(c)
    assert n == len(want) and n > 8
    assert [tuple(p) for p in xy.tolist()] == want


def edges_py(c, pts):
    """check_edges / check_edge (lib.rs:409-499): all ordered pairs, second iterator reversed; samples outside the image
    count as Other (the reference's unchecked usize arithmetic is undefined there)."""
    h, w = c.shape

    def at(x, y):
        return int(c[y, x]) if 0 <= x < w and 0 <= y < h else OTHER

    lines = []
    for x1, y1 in pts:
        for x2, y2 in reversed(pts):
            mx, my = (x1 + x2) // 2, (y1 + y2) // 2
            xd, yd = abs(x1 - x2), abs(y1 - y2)
            vert = x1 == x2 or xd < yd
            hori = y1 == y2 or yd < xd
            ax, ay, bx, by = (mx + x1) // 2, (my + y1) // 2, (mx + x2) // 2, (my + y2) // 2
            if vert:          # lib.rs:428-445: right = +5, left = -5 in x
                r1, r2, l1, l2 = at(ax + 5, ay), at(bx + 5, by), at(ax - 5, ay), at(bx - 5, by)
                if OTHER not in (r1, r2, l1, l2) and ((l1 == BLACK) ^ (r2 == BLACK)) and ((l2 == BLACK) ^ (r1 == BLACK)) and l1 == l2:
                    lines.append((x1, y1, x2, y2))
            if hori:          # lib.rs:447-473: top = -5, bottom = +5 in y
                t1, t2, b1, b2 = at(ax, ay - 5), at(bx, by - 5), at(ax, ay + 5), at(bx, by + 5)
                if OTHER not in (t1, t2, b1, b2) and ((t1 == BLACK) ^ (b2 == BLACK)) and ((t2 == BLACK) ^ (b1 == BLACK)) and t1 == t2:
                    lines.append((x1, y1, x2, y2))
    return lines


def test_check_edges_pairs_and_order(oracle):
    c = ternary_map(96, 72, 5)
    xy, n = oracle.cat_detect_corners(c)
    pts = [tuple(p) for p in xy.tolist()]
    lines, m = oracle.cat_check_edges(c, xy)
    want = edges_py(c, pts)
    assert m == len(want) and m > 4
    assert [tuple(l) for l in lines.tolist()] == want


def ccl_py(c):
    """connected_components (lib.rs:501-549) with a plain union-find; returns min-index labels and sizes"""
    h, w = c.shape
    parent = list(range(w * h))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    def union(a, b):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)

    for y in range(h):
        for x in range(1, w - 1):
            p = c[y, x]
            if p == OTHER:
                continue
            i = y * w + x
            if c[y, x - 1] == p:
                union(i, i - 1)
            if y > 0:
                if c[y - 1, x] == p:
                    union(i, i - w)
                if p == WHITE:
                    if c[y - 1, x - 1] == p:
                        union(i, i - w - 1)
                    if x < w - 1 and c[y - 1, x + 1] == p:
                        union(i, i - w + 1)
    lab = np.array([find(i) for i in range(w * h)], np.uint32).reshape(h, w)
    sizes = np.bincount(lab.reshape(-1), minlength=w * h)[lab].astype(np.uint32)
    return lab, sizes


def test_connected_components_partition(oracle):
    c = ternary_map(64, 40, 6)
    lab, sizes = oracle.cat_connected_components(c)
    wl, ws = ccl_py(c)
    assert (lab == wl).all() and (sizes == ws).all()


def test_golden_fixture_cat(oracle):
    """Regression pin: tests/golden/cat_96x72.npz (tests/golden/make_golden.py), every stage bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cat_96x72.npz"))
    color = oracle.cat_calc_otsu(g["rgb"])
    assert (color == g["otsu"]).all() and (oracle.cat_thresh(g["rgb"]) == g["thresh"]).all()
    xy, n = oracle.cat_detect_corners(color)
    assert n == len(g["corners"]) and (xy == g["corners"]).all()
    lines, m = oracle.cat_check_edges(color, xy)
    assert m == len(g["lines"]) and (lines == g["lines"]).all()
    labels, sizes = oracle.cat_connected_components(color)
    assert (labels == g["labels"]).all() and (sizes == g["sizes"]).all()


def test_external_map_entry_is_upstreams_pipeline_after_threshold(oracle):
    """orc_detect_with_map (the CAT decode oracle) fed with upstream's own threshold map is orc_detect: the entry changes where
    the ternary map comes from and nothing else."""
    from chalkydri_b200 import synth
    gray, _ = synth.render_frame(640, 480, 3, seed=7, edge_px=(50, 110))
    for f in (1.0, 2.0):
        prm = oracle.default_params(quad_decimate=f)
        ref = oracle.detect(gray, prm)
        got = oracle.detect_with_map(gray, oracle.threshold(gray, prm), prm)
        assert len(ref) >= 2 and got.tobytes() == ref.tobytes()
