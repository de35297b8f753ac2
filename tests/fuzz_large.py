"""tests/fuzz_parity.py on a few large frames (up to 4608 x 2592): exercises the big sort tiers, long prefix chains, many tiles."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fuzz_parity import random_frame, canon_quads
from chalkydri_b200.detector import DetectorBuilder
from oracle import pyoracle as po

rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
bad = 0
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for c in range(cases):
    W, H = int(rng.integers(1500, 4609)), int(rng.integers(1000, 2593))
    frame = random_frame(rng, W, H)
    det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, 1, 256).build()
    thr = det.threshold(frame[None]); lab, sz = det.labels(frame[None]); q, qc, _ = det.quads(frame[None]); out, counts = det.detect_batch(frame[None])
    ref, taps = po.detect(frame, taps=True, cap=1024, pts_cap=16_000_000)
    ok = (thr[0] == taps["thresh"]).all() and (lab[0] == taps["labels"]).all() and (sz[0] == taps["comp_size"]).all() and qc[0] == taps["nquads"]
    if ok and qc[0]:
        ok = np.abs(np.array(canon_quads(q[0, :qc[0]])) - np.array(canon_quads(taps["quads"]["p"]))).max() < 1e-4
    g = out[0, :counts[0]]
    ok = ok and g["id"].tolist() == ref["id"].tolist() and (len(g) == 0 or np.abs(g["p"] - ref["p"]).max() < 1e-3)
    cs = np.array([0])
    if "pts_cluster" in taps and taps["npoints"]:
        _, cs = np.unique(taps["pts_cluster"][:taps["npoints"]], return_counts=True)
    print(f"case {c}: {W}x{H} quads {qc[0]} dets {counts[0]} largest cluster {cs.max()} {'OK' if ok else 'MISMATCH'}")
    bad += not ok
    det.close()
print(f"{cases - bad}/{cases} large cases identical")
sys.exit(1 if bad else 0)
