"""Hand-built frames shared by the CPU (oracle) and GPU parity tests -- cases the random generator does not produce."""
import numpy as np

from chalkydri_b200 import synth


def nested_same_id_frame(big_id=7, small_id=7, flip_big=(), flip_small=(), big_levels=(230, 25), W=1456, H=1088, cell=100, scell=4):
    """A pixel-replicated tag of `cell` px per cell with a small copy (`scell` px per cell, quiet zone included) painted into one
    of its white data cells, away from the cell's centre so the big tag's bit still reads white.  Both decode; when the ids
    agree the two detections overlap (the small polygon lies inside the big one) and upstream's reconcile pass keeps exactly
    one of them: lower hamming first, then higher decision margin (apriltag.c, the loop after the decode workers)."""
    img = np.full((H, W), 140, np.uint8)
    g = synth.tag_pattern(big_id, flip_big)
    pat = np.kron(g, np.ones((cell, cell), np.uint8))
    n = pat.shape[0]
    x0, y0 = (W - n) // 2, (H - n) // 2
    img[y0:y0 + n, x0:x0 + n] = np.where(pat > 0, big_levels[0], big_levels[1])
    r, c = [(r, c) for r in range(2, 8) for c in range(2, 8) if g[r, c] == 1][0]
    sp = np.kron(synth.tag_pattern(small_id, flip_small), np.ones((scell, scell), np.uint8))
    m = sp.shape[0]
    yy, xx = y0 + r * cell + 6, x0 + c * cell + 6
    img[yy:yy + m, xx:xx + m] = np.where(sp > 0, 230, 25)
    return img


# (keyword arguments, which copy must survive the reconcile pass)
RECONCILE_CASES = [
    (dict(), None),                                             # both hamming 0: the higher decision margin wins (checked against the oracle)
    (dict(flip_small=(5,)), "big"),                             # small copy has one bit error: hamming decides
    (dict(flip_big=(5,)), "small"),                             # big copy has one bit error
    (dict(big_levels=(170, 110)), "small"),                     # equal hamming, low-contrast big copy: margin decides
    (dict(flip_big=(3, 17), flip_small=(8,)), "small"),         # hamming 2 against hamming 1
]


def bit_error_frame(seed=31, W=1280, H=720):
    """Four tags carrying 0, 1, 2 and 3 inverted data bits."""
    return synth.render_frame(W, H, 4, seed=seed, edge_px=(70, 150), bit_errors=(0, 1, 2, 3))


def yuv420_from_gray(grays, seed=0):
    """[B, H*3/2, W] buffers (NV12 / I420 layout: Y plane, then W*H/2 chroma bytes) whose Y planes are `grays`."""
    B, H, W = grays.shape
    rng = np.random.default_rng(seed)
    out = np.empty((B, H * 3 // 2, W), np.uint8)
    out[:, :H] = grays
    out[:, H:] = rng.integers(0, 256, (B, H // 2, W), dtype=np.uint8)
    return out
