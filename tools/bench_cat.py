"""The reference's only benchmark (crates/chalkydri-apriltags/bench.rs:8-26): Detector::new(703, 905, &[]) + process_frame on one
packed-RGB frame, timed per call.  GPU: cb_cat_process_frame (frame uploaded once, gray plane / colour map / corner list stay on
the device, lists come back) and cb_cat_detect_tags (CAT's threshold map, then the detector's stages); CPU: the oracle restatement
of the same three stages on one thread (the reference is single-threaded here).  The reference's test.png is not in the
repository; the frame is a synthetic 703x905 scene with four tags."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chalkydri_b200 import synth
from chalkydri_b200.cat import CatDetector

W, H = 703, 905


def run(iters=50, cpu_iters=3):
    gray, _ = synth.render_frame(W, H, 4, seed=2, edge_px=(60, 140), noise_sigma=0.0)
    hi = np.clip((gray.astype(np.float32) - 128.0) * 3.0 + 128.0, 0, 255).astype(np.uint8)       # contrast CAT's fixed 60 / 160 thresholds can use
    rgb = np.ascontiguousarray(np.repeat(hi[..., None], 3, axis=2))                             # a clean gray scene: 636 corners, 12 k lines
    d = CatDetector(W, H, ())
    for _ in range(5):
        d.process_frame(rgb, want_color=False)
    ts, dev = [], []
    for _ in range(iters):
        t0 = time.perf_counter(); d.process_frame(rgb, want_color=False); ts.append(time.perf_counter() - t0)
        dev.append(d.timing()["preprocess_ms"])
    res = {"workload": f"CAT Detector::new({W}, {H}, &[]) + process_frame on one packed-RGB frame (crates/chalkydri-apriltags/bench.rs:8-26), synthetic frame",
           "process_frame": {"p50_ms": float(np.percentile(np.array(ts) * 1e3, 50)), "kernels_ms": float(np.median(dev)),
                             "corners": int(len(d.points)), "lines": int(len(d.lines)),
                             "api": "cb_cat_process_frame: host RGB frame in, corner and line lists out"}}
    for _ in range(5):
        dets = d.detect_tags(rgb)
    tt = []
    for _ in range(iters):
        t0 = time.perf_counter(); dets = d.detect_tags(rgb); tt.append(time.perf_counter() - t0)
    res["detect_tags"] = {"p50_ms": float(np.percentile(np.array(tt) * 1e3, 50)), "tags": sorted(int(i) for i in dets["id"]),
                          "api": "cb_cat_detect_tags: CAT threshold map (thresh), then the detector's stages A3-A9, full resolution"}
    d.close()
    try:
        from oracle import pyoracle as po
        tc = []
        if cpu_iters <= 0:
            raise RuntimeError("skipped (--no-cpu-baseline)")
        for _ in range(cpu_iters):
            t0 = time.perf_counter()
            col = po.cat_calc_otsu(rgb)
            xy, n = po.cat_detect_corners(col)
            ln, m = po.cat_check_edges(col, xy[:n])
            tc.append(time.perf_counter() - t0)
        res["cpu_baseline"] = {"kind": "port", "cores": 1, "ms_per_frame": float(np.median(tc) * 1e3), "corners": int(n),
                               "lines": int(m), "sample": f"{cpu_iters} frames"}
        res["process_frame"]["same_counts_as_oracle"] = bool(n == res["process_frame"]["corners"] and m == res["process_frame"]["lines"])
    except Exception as e:                                   # noqa: BLE001
        res["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    return res


if __name__ == "__main__":
    print(json.dumps(run()))
