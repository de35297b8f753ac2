"""Stage-by-stage GPU vs oracle comparison with verbose diagnostics (debug aid; the assertions live in tests/)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth
from chalkydri_b200.detector import DetectorBuilder
from oracle import pyoracle as po


def canon_quads(q):
    """quads as a sorted list of corner sets, each rotated so the lexicographically smallest corner is first"""
    out = []
    for c in q:
        c = np.asarray(c, np.float64).reshape(4, 2)
        k = min(range(4), key=lambda i: (c[i, 0], c[i, 1]))
        out.append(np.roll(c, -k, 0).reshape(-1))
    return sorted(out, key=lambda v: tuple(v))


def report(W, H, ntags, seed, edge, B=2, **kw):
    frames, truths = synth.render_batch(W, H, B, ntags, seed=seed, edge_px=edge, **kw)
    det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, B, 128).build()
    print(f"== {W}x{H} tags={ntags} seed={seed} B={B}")
    thr = det.threshold(frames)
    lab, sz = det.labels(frames)
    q, qc, npts = det.quads(frames)
    out, counts = det.detect_batch(frames)
    print("timing", det.timing())
    for b in range(B):
        rd, taps = po.detect(frames[b], taps=True)
        tm = (thr[b] != taps["thresh"]).sum()
        lm = (lab[b] != taps["labels"]).sum()
        sm = (sz[b] != taps["comp_size"]).sum()
        print(f" frame {b}: thresh mismatches {tm}, label mismatches {lm}, size mismatches {sm}")
        gq = canon_quads(q[b, :qc[b]])
        oq = canon_quads(taps["quads"]["p"])
        print(f"   quads gpu {qc[b]} oracle {taps['nquads']}", end="")
        if len(gq) == len(oq) and len(gq):
            d = max(np.abs(a - c).max() for a, c in zip(gq, oq))
            print(f" max corner diff {d:.3e}")
        else:
            print(" (count differs)")
            so = set(tuple(np.round(v, 2)) for v in oq); sg = set(tuple(np.round(v, 2)) for v in gq)
            print("   only oracle:", len(so - sg), "only gpu:", len(sg - so))
        g = out[b, :counts[b]]
        print(f"   dets gpu ids {g['id'].tolist()} ham {g['hamming'].tolist()}")
        print(f"   dets orc ids {rd['id'].tolist()} ham {rd['hamming'].tolist()}")
        if len(g) == len(rd) and len(g) and (g["id"] == rd["id"]).all():
            print(f"   corner max diff {np.abs(g['p'] - rd['p']).max():.3e}  margin max rel {np.abs(g['decision_margin'] - rd['decision_margin']).max():.3e}"
                  f"  H max diff {np.abs(g['H'] - rd['H']).max():.3e}")
    det.close()


if __name__ == "__main__":
    report(1280, 720, 4, 1, (60, 150), B=2)
    report(1456, 1088, 8, 2, (40, 200), B=2)
    report(642, 486, 3, 3, (40, 100), B=1)          # decimates to 321 x 243: partial tiles, odd pitch -> generic paths
    report(4608, 2592, 40, 4, (40, 300), B=1, small_tags=10)
