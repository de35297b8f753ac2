"""Randomised GPU-vs-oracle parity: odd frame sizes, random shapes (rectangles, ellipses, lines, gradients, noise patches)
and pixel-replicated or warped tags.  Compares threshold map, partition, component sizes, quads (bit-level, tolerance 1e-4 px)
and detections.  usage: fuzz_parity.py [cases] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth
from chalkydri_b200.detector import DetectorBuilder
from oracle import pyoracle as po


def canon_quads(q):
    out = []
    for c in q:
        c = np.asarray(c, np.float64).reshape(4, 2)
        k = min(range(4), key=lambda i: (c[i, 0], c[i, 1]))
        out.append(tuple(np.roll(c, -k, 0).reshape(-1)))
    return sorted(out)


def random_frame(rng, W, H):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.full((H, W), float(rng.integers(60, 200)), np.float32)
    if rng.random() < 0.5:
        img += (xx / W - 0.5) * rng.uniform(-80, 80) + (yy / H - 0.5) * rng.uniform(-80, 80)
    for _ in range(int(rng.integers(2, 14))):
        kind = rng.integers(0, 5)
        v = float(rng.integers(0, 256))
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        w, h = int(rng.integers(3, max(4, W // 2))), int(rng.integers(3, max(4, H // 2)))
        if kind == 0:
            img[y0:y0 + h, x0:x0 + w] = v
        elif kind == 1:
            m = ((xx - x0) / max(w, 1)) ** 2 + ((yy - y0) / max(h, 1)) ** 2 < 0.25
            img[m] = v
        elif kind == 2:       # thin outline
            th = int(rng.integers(1, 5))
            img[y0:y0 + h, x0:x0 + w] = v
            img[y0 + th:y0 + h - th, x0 + th:x0 + w - th] = float(rng.integers(0, 256))
        elif kind == 3:       # slanted line
            a = rng.uniform(0, np.pi)
            d = np.abs((xx - x0) * np.sin(a) - (yy - y0) * np.cos(a))
            img[d < rng.uniform(0.8, 4.0)] = v
        else:                 # noise patch
            img[y0:y0 + h, x0:x0 + w] += rng.normal(0, rng.uniform(2, 30), img[y0:y0 + h, x0:x0 + w].shape)
    for _ in range(int(rng.integers(0, 4))):       # pixel-replicated tags
        cell = int(rng.integers(3, 14))
        pat = np.kron(synth.tag_pattern(int(rng.integers(0, 587))), np.ones((cell, cell), np.float32))
        n = pat.shape[0]
        if n + 2 >= min(W, H):
            continue
        x0, y0 = int(rng.integers(0, W - n)), int(rng.integers(0, H - n))
        lo, hi = sorted(rng.integers(0, 256, 2).tolist())
        if hi - lo < 40:
            lo, hi = 30, 220
        pat = np.rot90(pat, int(rng.integers(0, 4)))
        img[y0:y0 + n, x0:x0 + n] = np.where(pat > 0, hi, lo)
    if rng.random() < 0.5:
        img += rng.normal(0, rng.uniform(0.5, 4.0), img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
    bad = 0
    for c in range(cases):
        W, H = int(rng.integers(40, 900)), int(rng.integers(40, 700))
        if rng.random() < 0.3:
            W, H = (W // 16) * 16 + 16, (H // 4) * 4 + 4
        B = int(rng.integers(1, 4))
        frames = np.stack([random_frame(rng, W, H) for _ in range(B)])
        det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 256).build()
        thr = det.threshold(frames); lab, sz = det.labels(frames); q, qc, _ = det.quads(frames); out, counts = det.detect_batch(frames)
        msgs = []
        for b in range(B):
            ref, taps = po.detect(frames[b], taps=True, cap=1024)
            if not (thr[b] == taps["thresh"]).all(): msgs.append(f"frame {b}: threshold differs")
            if not (lab[b] == taps["labels"]).all(): msgs.append(f"frame {b}: partition differs")
            if not (sz[b] == taps["comp_size"]).all(): msgs.append(f"frame {b}: sizes differ")
            if qc[b] != taps["nquads"]: msgs.append(f"frame {b}: quads {qc[b]} vs {taps['nquads']}")
            else:
                gq, oq = canon_quads(q[b, :qc[b]]), canon_quads(taps["quads"]["p"])
                if gq and np.abs(np.array(gq) - np.array(oq)).max() > 1e-4: msgs.append(f"frame {b}: quad corners differ")
            g = out[b, :counts[b]]
            if g["id"].tolist() != ref["id"].tolist() or g["hamming"].tolist() != ref["hamming"].tolist():
                msgs.append(f"frame {b}: detections {g['id'].tolist()} vs {ref['id'].tolist()}")
            elif len(g) and np.abs(g["p"] - ref["p"]).max() > 1e-3:
                msgs.append(f"frame {b}: corners differ by {np.abs(g['p'] - ref['p']).max():.2e}")
        det.close()
        print(f"case {c}: {W}x{H} B={B} quads {qc.tolist()} dets {counts.tolist()} {'OK' if not msgs else 'MISMATCH ' + '; '.join(msgs)}")
        bad += bool(msgs)
    print(f"{cases - bad}/{cases} cases identical")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
