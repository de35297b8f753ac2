"""Generates the committed fixtures under tests/golden/ from the oracle (run once; outputs are committed).

There is no reference binary or golden vector to import, so these fixtures pin the ORACLE
against drift; the oracle itself is pinned by ground truth + cv2 cross-checks in tests/test_oracle_*.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chalkydri_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def detector_c1():
    seed = 1
    im, _ = synth.render_frame(1280, 720, 4, seed=seed, edge_px=(60, 150))
    dets, taps = po.detect(im, taps=True)
    np.savez_compressed(os.path.join(HERE, "detector_c1.npz"), seed=seed, frame=im, ids=dets["id"], hamming=dets["hamming"],
                        corners=dets["p"], margin=dets["decision_margin"], H=dets["H"],
                        thresh_white_bits=np.packbits(taps["thresh"] == 255), thresh_black_bits=np.packbits(taps["thresh"] == 0),
                        npoints=taps["npoints"], nquads=taps["nquads"], quads=taps["quads"]["p"])


if __name__ == "__main__":
    detector_c1()
    print("golden fixtures written")
