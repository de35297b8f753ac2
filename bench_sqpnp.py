"""BASELINE configs[4]: batched SQPnP pose solve for 1M tag corner sets vs the CPU solver (oracle restatement)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from chalkydri_b200.solver import SqPnP
from tests.sqpnp_problems import make_problems

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tags, bearings, n_tags, r2c, gyro, truth = make_problems(N, 0x5EED + 5, 0.1, 0.25)
s = SqPnP.new()
for _ in range(2):
    out, ok = s.solve_robot_pose_batch(tags, bearings, n_tags, r2c, gyro, 600.0)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    out, ok = s.solve_robot_pose_batch(tags, bearings, n_tags, r2c, gyro, 600.0)
    wall = time.perf_counter() - t0
    t = s.timing()
    ts.append((t["decode_ms"], t["total_ms"], wall * 1e3))
k_ms, tot_ms, wall_ms = np.median(np.array(ts), 0)
res = {"workload": f"c5: {N} SQPnP problems (90% one tag, 10% two tags, 0.25 px corner noise, gyro sigma 2 deg)",
       "kernel_ms": float(k_ms), "h2d_kernel_d2h_ms": float(tot_ms), "wall_ms": float(wall_ms),
       "problems_per_s_kernel": N / (k_ms * 1e-3), "problems_per_s_e2e": N / (wall_ms * 1e-3), "ok_fraction": float(ok.mean())}
from oracle import pyoracle as po
M = min(N, 20000)
cores = os.cpu_count() or 1
t0 = time.perf_counter()
ref, rok = po.sqpnp_batch(tags[:M], bearings[:M], n_tags[:M], r2c, gyro[:M], 600.0, nthreads=1)
t1 = time.perf_counter() - t0
t0 = time.perf_counter()
po.sqpnp_batch(tags[:M], bearings[:M], n_tags[:M], r2c, gyro[:M], 600.0, nthreads=cores)
tn = time.perf_counter() - t0
m = ok[:M].astype(bool) & rok.astype(bool)
res["cpu_baseline"] = {"kind": "port", "sample": f"{M} problems", "problems_per_s_1thread": M / t1, "problems_per_s_all": M / tn, "cores": cores}
res["parity_on_sample"] = {"ok_equal": bool((ok[:M] == rok).all()), "max_pos_diff": float(np.abs(out["pos"][:M][m] - ref["pos"][m]).max()),
                           "max_rot_diff": float(np.abs(out["rot"][:M][m] - ref["rot"][m]).max())}
print(json.dumps(res))
s.close()
