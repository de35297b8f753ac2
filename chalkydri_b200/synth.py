"""Deterministic synthetic tag-field frames (SURVEY.md 8d "Synthetic inputs").

Plumbing for tests and bench.py: renders tag36h11 tags (with a one-cell white quiet zone) through a
pinhole + OpenCV-5 distortion camera onto a mid-gray background with an illumination gradient, adds
Gaussian noise (sigma 3) and a 3x3 blur (sigma 0.8).  Pure numpy, seeded, no reference code involved.
Ground truth (ids, corner pixels in the detector's corner order, tag poses) is returned beside the pixels.
"""
from __future__ import annotations

import numpy as np

from .tagfamily import TAG36H11_CODES, BIT_X, BIT_Y

TAG_SIZE_M = 0.1651  # /root/reference/crates/chalkydri_sqpnp/src/lib.rs:38

# calibration of the 1280x720 camera in /root/reference/chalkydri.ron:98 (fx,fy,cx,cy,k1,k2,p1,p2,k3)
CALIB_1280x720 = (898.994806807896, 897.9156469180645, 627.0698256482966, 357.65273282451244,
                  -0.18595770381253796, 0.4406013374445432, -0.001, -0.001, -0.3704732841830049)


def scaled_calib(width: int, height: int, distortion: bool = True):
    """The chalkydri.ron:98 model rescaled to another resolution."""
    fx, fy, cx, cy, k1, k2, p1, p2, k3 = CALIB_1280x720
    sx, sy = width / 1280.0, height / 720.0
    s = 0.5 * (sx + sy)
    if not distortion:
        k1 = k2 = p1 = p2 = k3 = 0.0
    return (fx * s, fy * s, cx * sx, cy * sy, k1, k2, p1, p2, k3)


def tag_pattern(tag_id: int, flip_bits=()) -> np.ndarray:
    """10x10 cell image (1 = white) of a tag36h11 tag including the white quiet zone.  `flip_bits`: indices (0..35, AprilTag-3
    bit order) of data bits to invert -- a tag printed / seen with that many bit errors (hamming distance len(flip_bits))."""
    g = np.ones((10, 10), np.uint8)
    g[1:9, 1:9] = 0
    code = TAG36H11_CODES[tag_id]
    for i in flip_bits:
        code ^= 1 << (35 - int(i))
    for i in range(36):
        bit = (code >> (35 - i)) & 1
        g[BIT_Y[i] + 1, BIT_X[i] + 1] = bit
    return g


def _distort(xn, yn, k):
    k1, k2, p1, p2, k3 = k
    r2 = xn * xn + yn * yn
    radial = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
    dx = 2.0 * p1 * xn * yn + p2 * (r2 + 2.0 * xn * xn)
    dy = p1 * (r2 + 2.0 * yn * yn) + 2.0 * p2 * xn * yn
    return xn * radial + dx, yn * radial + dy


def project(calib, pts_cam):
    fx, fy, cx, cy = calib[:4]
    xn = pts_cam[..., 0] / pts_cam[..., 2]
    yn = pts_cam[..., 1] / pts_cam[..., 2]
    xd, yd = _distort(xn, yn, calib[4:])
    return np.stack([fx * xd + cx, fy * yd + cy], -1)


def _undistort(xd, yd, k, iters=12):
    k1, k2, p1, p2, k3 = k
    x, y = xd.copy(), yd.copy()
    if k1 == 0 and k2 == 0 and p1 == 0 and p2 == 0 and k3 == 0:
        return x, y
    for _ in range(iters):
        r2 = x * x + y * y
        radial = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
        dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
        dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
        x = (xd - dx) / radial
        y = (yd - dy) / radial
    return x, y


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


# detector corner order (apriltag.c): tag coords (-1,1),(1,1),(1,-1),(-1,-1) with y down in the tag image
_CORNER_TAG = np.array([[-1, 1], [1, 1], [1, -1], [-1, -1]], np.float64)


def render_frame(width: int, height: int, n_tags: int, seed: int, edge_px=(40.0, 200.0), max_tilt_deg: float = 60.0,
                 calib=None, small_tags: int = 0, small_edge_px=(12.0, 24.0), noise_sigma: float = 3.0,
                 blur: bool = True, ss: int = 3, bit_errors=0):
    """Render one frame. Returns (gray u8 [H,W], truth dict).  `bit_errors`: number of inverted data bits per tag (an int, or a
    sequence cycled over the tags); the flipped positions come from their own seeded stream, so frames rendered with
    bit_errors = 0 are unchanged.  truth["hamming"] holds the number of flipped bits of every placed tag."""
    rng = np.random.default_rng(seed)
    if calib is None:
        calib = scaled_calib(width, height)
    fx, fy, cx, cy = calib[:4]
    # background: mid-gray with a linear illumination gradient of +-40
    ang = rng.uniform(0, 2 * np.pi)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    u = ((xx - width / 2) * np.cos(ang) + (yy - height / 2) * np.sin(ang)) / (0.5 * np.hypot(width, height))
    illum = (1.0 + (40.0 / 128.0) * u).astype(np.float32)
    img = 128.0 * illum

    if n_tags <= 32:
        ids = rng.permutation(np.arange(1, 33))[:n_tags]   # ids that exist in field.json
    else:
        ids = rng.permutation(587)[:n_tags]
    placed = []   # (cx, cy, radius)
    truth = {"ids": [], "corners": [], "R": [], "t": [], "edge_px": [], "hamming": []}
    half = TAG_SIZE_M / 2.0
    for ti, tag_id in enumerate(ids):
        lo, hi = (small_edge_px if ti >= n_tags - small_tags else edge_px)
        ok = False
        for _attempt in range(200):
            edge = float(np.exp(rng.uniform(np.log(lo), np.log(hi))))
            rad = edge * 0.95  # bounding circle of tag + quiet zone (10/8 * edge * sqrt2 / 2 ~ 0.88 edge)
            px = rng.uniform(rad + 4, width - rad - 4) if width > 2 * rad + 8 else None
            py = rng.uniform(rad + 4, height - rad - 4) if height > 2 * rad + 8 else None
            if px is None or py is None:
                continue
            if all((px - a) ** 2 + (py - b) ** 2 > (rad + r) ** 2 for a, b, r in placed):
                ok = True
                break
        if not ok:
            continue
        tilt = np.deg2rad(max_tilt_deg)
        R = _rot(rng.uniform(-tilt, tilt) * 0.7, rng.uniform(-tilt, tilt) * 0.7, rng.uniform(-np.pi, np.pi))
        z = 0.5 * (fx + fy) * TAG_SIZE_M / edge
        xn, yn = _undistort(np.array([(px - cx) / fx]), np.array([(py - cy) / fy]), calib[4:])
        t = np.array([xn[0] * z, yn[0] * z, z])
        # corners of the black border (tag coords +-1 -> +-half metres); tag frame: x right, y down, z into the tag
        c3 = np.concatenate([_CORNER_TAG * half, np.zeros((4, 1))], 1) @ R.T + t
        if np.any(c3[:, 2] <= 0.05):
            continue
        cpx = project(calib, c3)
        q3 = np.concatenate([_CORNER_TAG * half * 1.25, np.zeros((4, 1))], 1) @ R.T + t
        qpx = project(calib, q3)
        x0, y0 = np.floor(qpx.min(0)).astype(int) - 2
        x1, y1 = np.ceil(qpx.max(0)).astype(int) + 3
        if x0 < 0 or y0 < 0 or x1 > width or y1 > height:
            continue
        placed.append((px, py, rad))
        # inverse mapping with ss x ss supersampling
        sub = (np.arange(ss) + 0.5) / ss
        gx = (np.arange(x0, x1)[:, None] + sub[None, :]).reshape(-1)          # pixel x in [x, x+1): centre convention
        gy = (np.arange(y0, y1)[:, None] + sub[None, :]).reshape(-1)
        # pixel (ix,iy) covers [ix, ix+1) x [iy, iy+1); the detector's corner coordinates use the same convention
        X, Y = np.meshgrid(gx, gy)
        xd, yd = (X - cx) / fx, (Y - cy) / fy
        xu, yu = _undistort(xd, yd, calib[4:])
        n = R[:, 2]
        lam = (n @ t) / (n[0] * xu + n[1] * yu + n[2])
        Pc = np.stack([lam * xu - t[0], lam * yu - t[1], lam - t[2]], -1)
        Pt = Pc @ R      # R^T applied to row vectors
        cell = TAG_SIZE_M / 8.0
        ux = np.floor(Pt[..., 0] / cell + 5.0).astype(int)
        uy = np.floor(Pt[..., 1] / cell + 5.0).astype(int)
        inside = (ux >= 0) & (ux < 10) & (uy >= 0) & (uy < 10)
        nflip = int(bit_errors if np.isscalar(bit_errors) else bit_errors[ti % len(bit_errors)])
        flips = np.random.default_rng([int(seed) & 0x7fffffff, 0xB17, ti]).permutation(36)[:nflip] if nflip else ()
        pat = tag_pattern(int(tag_id), flips)
        val = np.where(inside, pat[np.clip(uy, 0, 9), np.clip(ux, 0, 9)], 0).astype(np.float32)
        hh, ww = y1 - y0, x1 - x0
        alpha = inside.astype(np.float32).reshape(hh, ss, ww, ss).mean((1, 3))
        white = val.reshape(hh, ss, ww, ss).mean((1, 3))       # fraction of the pixel that is white tag
        ill = illum[y0:y1, x0:x1]
        tagcol = (white * 215.0 + (alpha - white) * 35.0) * ill
        img[y0:y1, x0:x1] = img[y0:y1, x0:x1] * (1 - alpha) + tagcol
        truth["ids"].append(int(tag_id))
        truth["corners"].append(cpx)
        truth["R"].append(R)
        truth["t"].append(t)
        truth["edge_px"].append(edge)
        truth["hamming"].append(nflip)
    if noise_sigma > 0:
        img = img + rng.normal(0.0, noise_sigma, img.shape).astype(np.float32)
    if blur:
        k = np.exp(-np.array([-1.0, 0.0, 1.0]) ** 2 / (2 * 0.8 ** 2)).astype(np.float32)
        k /= k.sum()
        p = np.pad(img, 1, mode="edge")
        img = k[0] * p[:-2, 1:-1] + k[1] * p[1:-1, 1:-1] + k[2] * p[2:, 1:-1]
        p = np.pad(img, 1, mode="edge")
        img = k[0] * p[1:-1, :-2] + k[1] * p[1:-1, 1:-1] + k[2] * p[1:-1, 2:]
    out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    truth["ids"] = np.array(truth["ids"], np.int32)
    truth["hamming"] = np.array(truth["hamming"], np.int32)
    truth["corners"] = np.array(truth["corners"], np.float64).reshape(-1, 4, 2)
    return out, truth


def render_batch(width: int, height: int, batch: int, n_tags, seed: int, unique: int | None = None, **kw):
    """[batch, H, W] u8 frames.  `n_tags` may be an int or a (lo, hi) range drawn per frame.  With `unique` < batch
    the first `unique` frames are rendered and the rest are copies with fresh sensor noise (+-2 gray levels)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    nu = batch if unique is None else min(unique, batch)
    frames = np.empty((batch, height, width), np.uint8)
    truths = []
    for b in range(nu):
        nt = n_tags if isinstance(n_tags, int) else int(rng.integers(n_tags[0], n_tags[1] + 1))
        frames[b], t = render_frame(width, height, nt, seed * 100003 + b, **kw)
        truths.append(t)
    for b in range(nu, batch):
        src = b % nu
        noise = rng.integers(-2, 3, (height, width), dtype=np.int16)
        frames[b] = np.clip(frames[src].astype(np.int16) + noise, 0, 255).astype(np.uint8)
        truths.append(truths[src])
    return frames, truths


def gray_to_rgb(gray: np.ndarray, seed: int = 0) -> np.ndarray:
    """Packed RGB variant for the CAT API (crates/chalkydri-apriltags/src/lib.rs:265-267): channels differ by a
    small seeded tint so the fused RGB->gray conversion is exercised."""
    rng = np.random.default_rng(seed)
    tint = rng.integers(-6, 7, 3)
    rgb = np.stack([np.clip(gray.astype(np.int16) + int(t), 0, 255) for t in tint], -1).astype(np.uint8)
    return np.ascontiguousarray(rgb)
