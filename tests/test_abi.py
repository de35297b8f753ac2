"""CPU checks of the drop-in boundary: the shared library loads and exports every symbol include/chalkydri_b200.h declares,
record layouts match the header, and the product fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from chalkydri_b200 import capi
    return capi.lib()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "chalkydri_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(lib):
    from chalkydri_b200 import capi
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(capi.EXPORTS) == syms


def test_record_layouts():
    from chalkydri_b200 import capi
    assert capi.DET_DTYPE.itemsize == 168 and capi.DET_DTYPE.fields["H"][1] == 16 and capi.DET_DTYPE.fields["p"][1] == 104
    assert capi.ISO_DTYPE.itemsize == 56 and capi.POSE_DTYPE.itemsize == 120
    assert C.sizeof(capi.Timing) == 9 * 4 + 8


def test_vision_measurement_contract(lib):
    """The reference's one test: size_of::<VisionMeasurement>() == 64 (crates/whacknet/src/lib.rs:92-95); plus what
    AprilTags::process publishes per frame (crates/apriltags/src/lib.rs:340-376)."""
    from chalkydri_b200 import capi
    from chalkydri_b200.solver import euler_angles
    assert capi.VISION_DTYPE.itemsize == 64
    poses = np.zeros(3, capi.POSE_DTYPE)
    a = 0.7
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    poses[0]["rot"] = R.T.reshape(-1)                       # column-major
    poses[0]["pos"] = (1.5, -2.25, 0.3)
    poses[0]["std_devs"] = (0.01, 0.02, 0.03)
    poses[2] = poses[0]
    ok = np.array([1, 0, 1], np.uint8)
    counts = np.array([3, 5, 400], np.int32)
    ts = np.array([10, 20, 30], np.uint64)
    out = np.zeros(3, capi.VISION_DTYPE)
    assert lib.cb_pack_vision_measurements(capi.ptr(poses), capi.ptr(ok), capi.ptr(counts), capi.ptr(ts), 7, 3, capi.ptr(out)) == 0
    assert out["ts"].tolist() == [10, 20, 30] and out["camera_id"].tolist() == [7, 7, 7]
    assert out["tag_count"].tolist() == [3, 0, 255]                      # None -> heartbeat record; count saturates like u8::MAX
    assert out[0]["x"] == 1.5 and out[0]["y"] == -2.25 and abs(out[0]["rot"] - euler_angles(R)[2]) < 1e-15
    assert (out[0]["std_x"], out[0]["std_y"], out[0]["std_rot"]) == (0.01, 0.02, 0.03)
    assert out[1]["x"] == 0 and out[1]["rot"] == 0 and out[1]["std_x"] == 0


def test_version_and_no_cpu_fallback(lib):
    assert b"sm_100a" in lib.cb_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only hosts")
    ctx = lib.cb_create(0, 640, 480, 1, 16)
    assert not ctx
    msg = lib.cb_last_error(None).decode()
    assert "no CPU fallback" in msg or "CUDA" in msg
    from chalkydri_b200.detector import DetectorBuilder
    from chalkydri_b200.capi import ChalkydriError
    with pytest.raises(ChalkydriError):
        DetectorBuilder.default().add_family_bits("tag36h11", 3).capacity(640, 480).build()


def test_builder_argument_errors():
    from chalkydri_b200.detector import DetectorBuilder
    with pytest.raises(ValueError):
        DetectorBuilder.default().add_family_bits("tag16h5", 1)
    with pytest.raises(ValueError):
        DetectorBuilder.default().build()


def test_camera_transform_matches_oracle(lib, oracle):
    """create_solver_camera_transform is host scalar math behind the C ABI: compare with the restatement."""
    from chalkydri_b200.solver import SqPnP
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.uniform(-1, 1, 3)
        e = rng.uniform(-180, 180, 3)
        got = SqPnP.create_solver_camera_transform(*a, *e)
        ref = oracle.create_solver_camera_transform(*a, *e)
        assert np.allclose(got["t"], ref["t"], atol=1e-14) and np.allclose(got["q"], ref["q"], atol=1e-15)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chalkydri_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle/", ""), f


def test_cpp_mirror_compiles_and_fails_loudly(lib, tmp_path):
    """include/chalkydri_b200.hpp builds against the .so; without a GPU the program reports the error instead of falling back."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    so_dir = os.path.join(ROOT, "chalkydri_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "abi_smoke.cpp"),
           "-L", so_dir, "-lchalkydri_b200", "-Wl,-rpath," + so_dir, "-o", exe]
    subprocess.check_call(cmd)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "version chalkydri_b200" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_runs_on_the_gpu(lib, tmp_path):
    """The same program on a GPU box: Detector::detect, the streaming form and the AprilTags task (process / submit / collect)."""
    test_cpp_mirror_compiles_and_fails_loudly(lib, tmp_path)


def test_apriltag_detections_payload():
    """AprilTagDetections (crates/apriltags/src/lib.rs:47-141): capacity 16, strict margin filter, tuple round trip."""
    from chalkydri_b200 import capi
    from chalkydri_b200.pipeline import MAX_DETECTIONS, AprilTagDetections
    dets = np.zeros(20, capi.DET_DTYPE)
    dets["id"] = np.arange(20)
    dets["decision_margin"] = np.linspace(10, 105, 20, dtype=np.float32)
    a = AprilTagDetections.from_detections(dets)
    assert len(a) == MAX_DETECTIONS == 16
    with pytest.raises(OverflowError):
        a.push(99, np.eye(4), 1.0)
    kept = list(a.filtered_by_decision_margin(float(dets["decision_margin"][4])))
    assert [k[0] for k in kept] == list(range(5, 16))                 # strictly greater, order kept
    assert kept[0][1].dtype == np.float32 and kept[0][1].shape == (4, 4)
    b = AprilTagDetections.from_tuples(a.to_tuples())
    assert b.ids == a.ids and b.decision_margins == a.decision_margins
    assert all((p == q).all() for p, q in zip(a.poses, b.poses))
    assert list(AprilTagDetections().filtered_by_decision_margin(0.0)) == []


def test_task_publish_rules_host_logic():
    """AprilTags::process publish rules (crates/apriltags/src/lib.rs:340-376) in the Python task mirror, without a device: a solved
    frame publishes (x, y, euler yaw) with the saturating tag count and ts = now - frame time; an unsolved one publishes the
    empty heartbeat only when more than 5 ms have passed since the last heartbeat."""
    from chalkydri_b200 import capi
    from chalkydri_b200.pipeline import AprilTags, RobotPose, VisionUncertainty

    class Comm:
        def __init__(self):
            self.sent = []

        def gyro_angle(self):
            return 0.5

        def publish(self, cam_id, tag_count, ts_us, pose, unc):
            self.sent.append((cam_id, tag_count, ts_us, pose, unc))

    task = object.__new__(AprilTags)              # the publish path needs no detector
    task.comm, task.cam_id, task.last_time = Comm(), 9, None
    a = 0.4
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    poses = np.zeros(4, capi.POSE_DTYPE)
    poses["rot"][:] = R.T.reshape(-1)             # column-major like nalgebra
    poses["pos"][:] = (2.0, -1.0, 0.1)
    poses["std_devs"][:] = (0.1, 0.2, 0.3)
    ok = np.array([1, 0, 0, 1], np.uint8)
    counts = np.array([2, 0, 3, 300], np.int32)
    res = task._publish_batch(10_000_000, [9_990_000, 9_991_000, 9_992_000, 9_993_000], counts, poses, ok)
    assert [r is not None for r in res] == [True, False, False, True]
    sent = task.comm.sent
    assert [(m[0], m[1], m[2]) for m in sent] == [(9, 2, 10_000), (9, 0, 9_000), (9, 255, 7_000)]     # second unsolved frame: no heartbeat
    assert isinstance(sent[0][3], RobotPose) and abs(sent[0][3].rot - a) < 1e-12 and (sent[0][3].x, sent[0][3].y) == (2.0, -1.0)
    assert (sent[0][4].x, sent[0][4].y, sent[0][4].rot) == (0.1, 0.2, 0.3)
    assert sent[1][3] == RobotPose() and sent[1][4] == VisionUncertainty()
    # 5 ms later exactly: still quiet (the reference tests `> 5`); 6 ms later: heartbeat again
    task._publish_batch(10_005_000, [10_000_000], counts[1:2], poses[1:2], ok[1:2])
    assert len(sent) == 3
    task._publish_batch(10_006_000, [10_000_000], counts[1:2], poses[1:2], ok[1:2])
    assert len(sent) == 4 and sent[3][1] == 0
    # gyro defaults to comm.gyro_angle() for every frame, None -> NaN (no reading, lib.rs:329)
    assert task._gyro_array(None, 3).tolist() == [0.5, 0.5, 0.5]
    assert np.isnan(task._gyro_array([0.1, None], 2)[1])


def test_cluttered_frame_overflow_becomes_a_heartbeat():
    """A device-table overflow (CB_ERR_OVERFLOW) on one frame must not take the task down (upstream's detector has no fixed
    tables): `process` treats it as "nothing detected" and publishes the heartbeat; any other library error still propagates."""
    from chalkydri_b200 import capi
    from chalkydri_b200.pipeline import AprilTags

    class Comm:
        sent = []

        def gyro_angle(self):
            return 0.0

        def publish(self, *a):
            self.sent.append(a)

    class Det:
        code = capi.CB_ERR_OVERFLOW

        def detect_batch(self, frames):
            raise capi.ChalkydriError(self.code, "device table overflow")

    task = object.__new__(AprilTags)
    task.comm, task.cam_id, task.last_time, task.detector = Comm(), 3, None, Det()
    assert task.process(1_000_000, 990_000, np.zeros((8, 8), np.uint8)) is None
    assert len(task.comm.sent) == 1 and task.comm.sent[0][1] == 0 and task.overflow_batches == 1
    task.detector.code = capi.CB_ERR_CUDA
    with pytest.raises(capi.ChalkydriError):
        task.process(2_000_000, 1_990_000, np.zeros((8, 8), np.uint8))


def test_rust_extern_block_lists_the_header():
    """rust/chalkydri-b200-sys/src/ffi.rs declares exactly the symbols of include/chalkydri_b200.h (text comparison: the image has
    no rustc), and the wrapper crate uses every one of them."""
    ffi = open(os.path.join(ROOT, "rust", "chalkydri-b200-sys", "src", "ffi.rs")).read()
    block = ffi[ffi.index('unsafe extern "C" {'):]
    block = block[:block.index("\n}\n")]
    declared = re.findall(r"pub fn (cb_[a-z0-9_]+)\s*\(", block)
    assert len(declared) == len(set(declared))
    assert sorted(declared) == header_symbols()
    wrappers = open(os.path.join(ROOT, "rust", "chalkydri-b200-sys", "src", "lib.rs")).read()
    unused = [s for s in declared if not re.search(r"\b" + s + r"\b", wrappers)]
    assert unused == [], f"declared but never called from the wrapper crate: {unused}"
    # the seams the reference's task code needs (crates/apriltags/src/lib.rs:19,229,259; crates/chalkydri_sqpnp/src/lib.rs:182,194)
    for needle in ("impl FromStr for Family", "pub fn add_family_bits(mut self, family: Family, bits_corrected: usize)",
                   "impl Default for SqPnP", "impl Clone for SqPnP", "#[derive(Debug)]\npub struct SqPnP",
                   "pub fn new(width: usize, height: usize, valid_tags: &'static [usize])", "pub fn process_frame(&mut self, input: &[u8])",
                   "pub fn connected_components(&self) -> UnionFind"):
        assert needle in wrappers, needle


def test_cpp_header_covers_the_abi():
    """include/chalkydri_b200.hpp (the host mirror that IS compiled and run) calls every entry point a host needs; the ones it
    leaves to the raw header are stage taps and plumbing."""
    hpp = open(os.path.join(ROOT, "include", "chalkydri_b200.hpp")).read()
    used = {s for s in header_symbols() if re.search(r"\b" + s + r"\b", hpp)}
    for s in ("cb_detect_gray", "cb_detect_gray_submit", "cb_detect_gray_collect", "cb_detect_pose_gray", "cb_sqpnp_batch",
              "cb_cat_process_frame", "cb_cat_connected_components", "cb_pool_detect_gray", "cb_pack_vision_measurements"):
        assert s in used


def test_integration_guide_indexes_every_symbol():
    """INTEGRATION.md section 5 names every symbol the header declares, each with the reference interface it stands for."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    index = doc[doc.index("## 5. Symbol index"):]
    missing = [s for s in header_symbols() if ("`" + s + "`") not in index]
    assert missing == [], missing
    assert index.count("lib.rs:") >= 10                    # file:line citations, not prose
