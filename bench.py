#!/usr/bin/env python
"""bench.py -- frames/s of the AprilTag detection hot path on N B200 GPUs (BASELINE.json metric).

A "step" is one pass of the whole detector (decimate+threshold -> union-find -> gradient clusters -> quad fit ->
refine/decode/reconcile) over one batch of synthetic frames of the configuration BASELINE.json quotes for one GPU
(configs[1]: 256 frames of 1456x1088 with 8 tags each).  Frames are independent, so N GPUs each process their own
batch (weak scaling, no collective on the data path; rank 0 gathers per-rank detection counts).

  value  : device-resident frames/s (frames already in HBM; CUDA events on the library's stream, max over ranks)
  e2e    : frames/s through the public API with HOST (pinned) frames, H2D and D2H inside the timed region; the headline uses
           the streaming form of the call (submit batch k+1, collect batch k), e2e.sync_call the blocking call
  roofline: the HBM-bound kernel the north star names (fused decimate+threshold), algorithmic bytes 0.75*W*H per frame
  cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/), timed on this box's cores

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, TAGS, BATCH = 1456, 1088, 8, 256        # BASELINE.json configs[1]
WORKLOAD = "c2: 256 x 1456x1088 gray frames, 8 tag36h11 tags each (BASELINE.json configs[1])"
# BASELINE.json's metric string, verbatim (it names the reference's own 1280x720 CPU case, configs[0]); the workload this
# bench measures is configs[1] -- the single-GPU configuration the metric is quoted on -- and is named in config.workload.
METRIC = "frames/sec at 1/2/4/8 B200 (1280\u00d7720 tag36h11); p50 per-frame latency"
try:
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "BASELINE.json")) as _f:
        METRIC = json.load(_f)["metric"]
except (OSError, KeyError, ValueError):
    pass
UNIT = "frames/s"
# dram__bytes_read.sum + dram__bytes_write.sum of threshold_f2_tma_kernel over a 128-frame launch (ncu --set full), per frame;
# the ternary map (0.25*W*H per frame) is only partly written back to DRAM inside the kernel (it stays in the 126 MB L2)
NCU_TRAFFIC_BYTES_PER_FRAME = (110.668032e6 + 20.104960e6) / 128


def make_frames(rank: int, batch: int = BATCH, unique: int = 16):
    from chalkydri_b200 import synth
    frames, truths = synth.render_batch(W, H, batch, TAGS, seed=0x5EED + 2 + 1000 * rank, unique=unique, edge_px=(40.0, 200.0))
    return frames, truths


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML (the library behind nvidia-smi; a query takes
    microseconds, so the region gets dozens of samples), falling back to the nvidia-smi command line of the profiling recipe."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.source = gpu_index, [], set(), False, None, None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # CUDA_VISIBLE_DEVICES remaps ordinals: resolve the physical device through its UUID / PCI bus id when torch knows it
        try:
            import torch
            h = nv.nvmlDeviceGetHandleByPciBusId(torch.cuda.get_device_properties(self.gpu).pci_bus_id.encode()) \
                if hasattr(torch.cuda.get_device_properties(self.gpu), "pci_bus_id") else nv.nvmlDeviceGetHandleByIndex(self.gpu)
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for n, b in bits.items():
                if r & b:
                    self.reasons.add(n)
            time.sleep(0.01)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0]))
                self.max_mhz = float(o[1])
                for n, v in zip(names, o[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


def cpu_baseline(frames: np.ndarray, seconds_target: float = 12.0):
    """oracle detector on a bounded sample of the same workload, all host threads (one frame per thread)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    po.detect_batch(frames[:1], cap=64, nthreads=1)
    t1 = time.perf_counter() - t0
    n = int(max(cores, min(len(frames), seconds_target / max(t1, 1e-3) * cores)))
    n = min(n, len(frames))
    t0 = time.perf_counter()
    _, counts = po.detect_batch(frames[:n], cap=64, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} frames of the c2 workload, one frame per thread on {cores} threads "
                      f"(single-thread first frame: {1.0 / t1:.1f} frames/s); CPU restatement of the reference path (oracle/), "
                      f"{int(counts.sum())} detections"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; the Rust reference cannot be built here) on host cores."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    sample = max(2 * cores, 32)
    frames, _ = make_frames(0, batch=sample, unique=min(sample, 16))
    for _ in range(args.warmup):
        po.detect_batch(frames[:cores], cap=64, nthreads=cores)
    t0 = time.perf_counter()
    ndet = 0
    for _ in range(args.steps):
        _, counts = po.detect_batch(frames, cap=64, nthreads=cores)
        ndet += int(counts.sum())
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "frames_per_step": sample, "note": "bounded sample of the c2 workload per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} frames per step x {args.steps} steps, one frame per thread on {cores} threads; "
                                       f"the Rust reference (un-vendored git deps, no cargo) cannot be built here, so this is oracle/"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "detections": ndet}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c1", action="store_true", help="skip the secondary 1280x720 measurement")
    ap.add_argument("--latency-iters", type=int, default=50)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG=VERSION and up) with printf when the communicator comes up; stdout must
        # carry exactly one JSON line, so fd 1 points at stderr until the first collective has run.
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from chalkydri_b200 import capi
    from chalkydri_b200.detector import DetectorBuilder, DET_DTYPE

    frames, truths = make_frames(rank)
    det = DetectorBuilder.default().add_family_bits("tag36h11", 3).device(local_rank).capacity(W, H, BATCH, 64).build()
    L = capi.lib()
    # pinned host copies (e2e arm) and a device-resident copy (value arm)
    h_frames = capi.pinned_array(frames.shape, np.uint8)
    h_frames[...] = frames
    out = capi.pinned_array((BATCH, 64), DET_DTYPE)
    counts = capi.pinned_array((BATCH,), np.int32)
    d_frames = L.cb_device_alloc(det.ctx, frames.nbytes)
    assert d_frames, "device allocation failed"
    assert L.cb_memcpy_h2d(det.ctx, d_frames, capi.ptr(h_frames), frames.nbytes) == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    for _ in range(max(args.warmup, 3)):
        det.detect_batch_device(d_frames, BATCH, H, W, out=out, counts=counts)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    stage = {k: 0.0 for k in ("threshold_ms", "ccl_ms", "cluster_ms", "quad_ms", "decode_ms", "d2h_ms", "total_ms")}
    launches = thr_launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        det.detect_batch_device(d_frames, BATCH, H, W, out=out, counts=counts)
        t = det.timing()
        for k in stage:
            stage[k] += t[k]
        launches += t["kernel_launches"]
        thr_launches += t["threshold_launches"]
    barrier()
    wall_dev = time.perf_counter() - t0
    dev_ms = stage["total_ms"]                       # CUDA events on the library's stream (kernels + list D2H)
    ndet = int(counts.sum())
    want = sum(len(t["ids"]) for t in truths)

    # ---- end-to-end arm: pinned host frames in, detection lists out ----
    for _ in range(2):
        det.detect_batch(h_frames, out=out, counts=counts)
    barrier()
    t0 = time.perf_counter()
    e2e_dev_ms = 0.0
    for _ in range(args.steps):
        det.detect_batch(h_frames, out=out, counts=counts)
        e2e_dev_ms += det.timing()["total_ms"]
    barrier()
    wall_e2e_sync = time.perf_counter() - t0

    # ---- the same arm through the streaming form of the call (cb_detect_gray_submit / _collect): batch k+1 is submitted
    #      before batch k is collected, as in a camera loop; every step's H2D and D2H are inside the timed region ----
    def stream_steps(d, hf, n):
        d.submit(hf)
        for s_ in range(n):
            if s_ + 1 < n:
                d.submit(hf)
            d.collect(out=out, counts=counts)

    stream_steps(det, h_frames, 2)
    barrier()
    t0 = time.perf_counter()
    stream_steps(det, h_frames, args.steps)
    barrier()
    wall_e2e = time.perf_counter() - t0
    ndet_stream = int(counts.sum())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- p50 single-frame latency through the public API (host frame in, list out) ----
    lat = []
    one = h_frames[:1]
    for i in range(args.latency_iters + 5):
        t0 = time.perf_counter()
        det.detect_batch(one, out=out[:1], counts=counts[:1])
        if i >= 5:
            lat.append((time.perf_counter() - t0) * 1e3)
    p50 = float(np.median(lat)) if lat else None

    # ---- the metric string names 1280x720: the same two arms on a 256-frame batch of the c1 resolution (N = 1 only) ----
    also = None
    if world == 1 and not args.no_c1:
        from chalkydri_b200 import synth
        W1, H1 = 1280, 720
        f1, _ = synth.render_batch(W1, H1, BATCH, 4, seed=0x5EED + 1, unique=8, edge_px=(60.0, 150.0))
        det1 = DetectorBuilder.default().add_family_bits("tag36h11", 3).device(local_rank).capacity(W1, H1, BATCH, 64).build()
        h1 = capi.pinned_array(f1.shape, np.uint8)
        h1[...] = f1
        d1 = L.cb_device_alloc(det1.ctx, f1.nbytes)
        assert d1 and L.cb_memcpy_h2d(det1.ctx, d1, capi.ptr(h1), f1.nbytes) == 0
        for _ in range(3):
            det1.detect_batch_device(d1, BATCH, H1, W1, out=out, counts=counts)
        ms = 0.0
        for _ in range(args.steps):
            det1.detect_batch_device(d1, BATCH, H1, W1, out=out, counts=counts)
            ms += det1.timing()["total_ms"]
        nd1 = int(counts.sum())
        for _ in range(2):
            det1.detect_batch(h1, out=out, counts=counts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            det1.detect_batch(h1, out=out, counts=counts)
        torch.cuda.synchronize()
        w1 = time.perf_counter() - t0
        stream_steps(det1, h1, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        stream_steps(det1, h1, args.steps)
        torch.cuda.synchronize()
        w1s = time.perf_counter() - t0
        also = {"workload": "256 x 1280x720 gray frames, 4 tag36h11 tags each (the resolution the metric string names)",
                "value": BATCH * args.steps / (ms / 1e3), "e2e": BATCH * args.steps / w1s, "e2e_sync_call": BATCH * args.steps / w1,
                "unit": UNIT, "detections_per_step": nd1}
        L.cb_device_free(det1.ctx, d1)
        det1.close()

    # max over ranks
    def rmax(x):
        if world == 1:
            return x
        t_ = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    def rsum(x):
        if world == 1:
            return x
        t_ = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_, op=dist.ReduceOp.SUM)
        return float(t_.item())

    dev_s = rmax(dev_ms / 1e3)
    wall_dev = rmax(wall_dev)
    wall_e2e = rmax(wall_e2e)
    wall_e2e_sync = rmax(wall_e2e_sync)
    total_det = rsum(ndet)
    frames_total = BATCH * args.steps * world
    value = frames_total / dev_s
    e2e_value = frames_total / wall_e2e

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        thr_bytes = 0.75 * W * H * BATCH                       # algorithmic bytes of one launch
        thr_ms = stage["threshold_ms"] / max(thr_launches, 1)
        achieved = thr_bytes / (thr_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": BATCH, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                       "l2": "inputs (405 MB per step) exceed the 126 MB L2", "unique_frames": 16},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
                    "d2h_bytes_per_step": int(out.nbytes + counts.nbytes), "ms_per_step_wall": wall_e2e / args.steps * 1e3,
                    "api": "cb_detect_gray_submit / cb_detect_gray_collect: pinned host frames in, detection lists out, batch k+1 "
                           "submitted before batch k is collected (two in flight); every step's copies are inside the timed region",
                    "detections_per_step": ndet_stream,
                    "sync_call": {"api": "cb_detect_gray (one blocking call per step)", "value": frames_total / wall_e2e_sync,
                                  "ms_per_step_wall": wall_e2e_sync / args.steps * 1e3,
                                  "ms_per_step_device_events": e2e_dev_ms / args.steps}},
            "also_1280x720": also,
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "roofline": {"kernel": "threshold_f2_tma_kernel (fused decimate + tile min/max + 3x3 dilate + binarise, TMA-staged)", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_FRAME * BATCH,
                         "traffic_source": "ncu --set full, profiles/r01_ncu_threshold_tma.txt: (dram read + write) / 128 frames", "ms_per_launch": thr_ms,
                         "algorithmic_bytes_per_launch": thr_bytes},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "wall_ms_per_step_device_arm": wall_dev / args.steps * 1e3,
            "p50_frame_latency_ms": p50,
            "detections_per_step": total_det, "expected_tags_per_step": want * world,
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(frames)
        print(json.dumps(line), flush=True)
    L.cb_device_free(det.ctx, d_frames)
    det.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
