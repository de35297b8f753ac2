import sys, os
import numpy as np
sys.path.insert(0, '/root/repo')
from chalkydri_b200 import synth
from chalkydri_b200.detector import DetectorBuilder
from oracle import pyoracle as po
sys.path.insert(0, '/root/repo/tools')
from gpu_parity_report import canon_quads
W,H=1280,720
frames,_=synth.render_batch(W,H,2,4,seed=1,edge_px=(60,150))
det = DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(W, H, 2, 128).build()
for rep in range(3):
    q,qc,npts=det.quads(frames)
    print(rep, qc, npts)
for b in range(2):
    rd,taps=po.detect(frames[b],taps=True, pts_cap=4000000)
    gq=canon_quads(q[b,:qc[b]]); oq=canon_quads(taps['quads']['p'])
    so={tuple(np.round(v,2)):i for i,v in enumerate(oq)}; sg={tuple(np.round(v,2)) for v in gq}
    for k in so:
        if k not in sg:
            print('missing on gpu', b, k)
            # find the oracle cluster size
            print(taps['quads'][:0].dtype)
    for k in sg:
        if k not in so: print('extra on gpu', b, k)
