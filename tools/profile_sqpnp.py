import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200.solver import SqPnP
from tests.sqpnp_problems import make_problems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
tags, bearings, n_tags, r2c, gyro, truth = make_problems(N, 0x5EED + 5, 0.1, 0.25)
s = SqPnP.new()
for _ in range(2):
    out, ok = s.solve_robot_pose_batch(tags, bearings, n_tags, r2c, gyro, 600.0)
    print(s.timing()["decode_ms"], ok.mean())
s.close()
