"""Kernel time of small SQPnP calls (events around the three phases): python tools/sq_small_ab.py   (CB_SQ_SMALL=0: thread-per-system form)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200.solver import SqPnP
from tests.sqpnp_problems import make_problems
tags, bearings, n_tags, r2c, gyro, _ = make_problems(4096, 11, 0.5, 0.25)
s = SqPnP.new()
for n in (1, 8, 64, 256, 512, 1024, 4096):
    ts = []
    for _ in range(12):
        s.solve_robot_pose_batch(tags[:n], bearings[:n], n_tags[:n], r2c, gyro[:n], 600.0)
        ts.append(s.timing()["decode_ms"])
    print("CB_SQ_SMALL=%s  n=%5d  kernels min %.3f ms  median %.3f ms" % (os.environ.get("CB_SQ_SMALL", "default"), n, min(ts), float(np.median(ts))))
s.close()
