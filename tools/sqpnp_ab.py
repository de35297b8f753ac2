"""A/B of the SQPnP forms: CB_SQPNP=fused (round-1 one-kernel form) against the three-phase default, bit for bit, plus timing.
usage: python tools/sqpnp_ab.py [N]   (run once per form; the second run compares with the file the first one wrote)"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200.solver import SqPnP
from tests.sqpnp_problems import make_problems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
tags, bearings, n_tags, r2c, gyro, truth = make_problems(N, 0x5EED + 5, 0.1, 0.25)
s = SqPnP.new()
ts = []
for _ in range(4):
    out, ok = s.solve_robot_pose_batch(tags, bearings, n_tags, r2c, gyro, 600.0)
    ts.append(s.timing()["decode_ms"])
form = os.environ.get("CB_SQPNP", "phases")
path = "gpurun_out/sqpnp_ab_%d.npz" % N
print(form, "kernel ms", [round(t, 2) for t in ts], "problems/s", round(N / (min(ts) * 1e-3)), "ok", float(ok.mean()))
if os.path.exists(path):
    ref = np.load(path)
    same_ok = bool((ref["ok"] == ok).all())
    m = ok.astype(bool) & ref["ok"].astype(bool)
    print("vs first run: ok equal", same_ok, "pos bit-equal", bool((out["pos"][m] == ref["pos"][m]).all()), "rot bit-equal", bool((out["rot"][m] == ref["rot"][m]).all()),
          "std bit-equal", bool((out["std_devs"][m] == ref["std_devs"][m]).all()), "max |dpos|", float(np.abs(out["pos"][m] - ref["pos"][m]).max()))
else:
    np.savez(path, ok=ok, pos=out["pos"], rot=out["rot"], std_devs=out["std_devs"])
s.close()
