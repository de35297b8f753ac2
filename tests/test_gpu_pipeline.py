"""Row B0 (AprilTags::process, crates/apriltags/src/lib.rs:293-379) through the fused device path cb_detect_pose_gray,
against the CPU restatement run stage by stage (oracle detect -> oracle un-project -> oracle solve_robot_pose)."""
import json

import numpy as np
import pytest

from chalkydri_b200 import synth

pytestmark = pytest.mark.gpu


class Comm:
    def __init__(self, gyro):
        self.gyro, self.published = gyro, []

    def gyro_angle(self):
        return self.gyro

    def publish(self, cam_id, tag_count, ts_us, pose, unc):
        self.published.append((cam_id, tag_count, ts_us, pose, unc))


def make_task(W, H, B, comm):
    from chalkydri_b200.pipeline import AprilTags
    calib = synth.scaled_calib(W, H)
    keys = ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")
    config = {"family": "tag36h11", "bits_corrected": 3, "cam_id": 7,
              "robot_to_cam": json.dumps({"x": 0.2, "y": -0.1, "z": 0.5, "roll": 0.0, "pitch": -10.0, "yaw": 15.0}),
              "calib": json.dumps({"OpenCVModel5": dict(zip(keys, calib))})}
    return AprilTags.new(config, comm, max_width=W, max_height=H, max_batch=B), np.array(calib, np.float64)


def oracle_pose(oracle, task, cam_params, frame, gyro):
    """the reference's process() body on the CPU restatement"""
    from chalkydri_b200.capi import ISO_DTYPE
    from chalkydri_b200.solver import SIGN_FLIP_CONST
    dets = oracle.detect(frame)
    world, cam = [], []
    for d in dets:
        tag = task.tags.get(int(d["id"]))
        if tag is None:
            continue
        bs = [oracle.unproject_opencv5(cam_params, float(c[0]), float(c[1])) for c in d["p"]]
        if all(b is not None for b in bs):
            world.append(tag)
            cam.append(np.array(bs))
    if gyro is None or not world:
        return len(dets), None
    res = oracle.sqpnp_solve_robot_pose(np.array(world, ISO_DTYPE), np.concatenate(cam), np.array(task.robot_to_cam, ISO_DTYPE), gyro, SIGN_FLIP_CONST)
    return len(dets), res


def test_fused_detect_pose_matches_stagewise_oracle(oracle):
    W, H, B = 1280, 720, 6
    # one tag per frame is always a consistent scene for the field layout (randomly placed tag sets are not, and mostly end in
    # solve_robot_pose -> None on both sides); frame 1 carries four tags, frame 4 none (heartbeat)
    frames, _ = synth.render_batch(W, H, B, 1, seed=21, edge_px=(90, 200))
    frames[1] = synth.render_batch(W, H, 1, 4, seed=22, edge_px=(70, 160))[0][0]
    frames[4] = 128
    comm = Comm(0.3)
    task, cam_params = make_task(W, H, B, comm)
    gyro = [0.3, 0.3, None, -1.2, 0.3, 2.0]           # frame 2: no gyro reading -> no solve
    res = task.process_batch(10_000_000, [9_990_000] * B, frames, gyro=gyro)
    out, counts, poses, ok, ntags = task.last_batch
    solved = 0
    for b in range(B):
        ndet, ref = oracle_pose(oracle, task, cam_params, frames[b], gyro[b])
        assert counts[b] == ndet
        if ref is None:
            assert res[b] is None and not ok[b]
            continue
        assert res[b] is not None and ok[b], f"frame {b}: the fused path returned None"
        rot, pos, std = ref["rot"].reshape(3, 3).T, ref["pos"], ref["std_devs"]
        got_rot = poses[b]["rot"].reshape(3, 3).T
        scale = max(1.0, float(np.abs(pos).max()))
        assert np.abs(poses[b]["pos"] - pos).max() < 1e-4 * scale            # north star: pose within 1e-4 relative
        assert np.abs(got_rot - rot).max() < 1e-4
        assert np.allclose(poses[b]["std_devs"], std, rtol=1e-4, atol=1e-9)
        solved += 1
    assert solved >= 3
    # publish contract: one message per solved frame (tag_count = detections), heartbeat for the others at most once per 5 ms
    assert sum(1 for m in comm.published if m[1] > 0) == solved
    assert all(m[0] == 7 and m[2] == 10_000 for m in comm.published)


def test_fused_path_equals_single_frame_process(oracle):
    W, H = 1280, 720
    frames, _ = synth.render_batch(W, H, 2, 4, seed=33, edge_px=(70, 160))
    c1, c2 = Comm(0.1), Comm(0.1)
    t1, _ = make_task(W, H, 2, c1)
    t2, _ = make_task(W, H, 2, c2)
    batch = t1.process_batch(5_000_000, [4_999_000, 4_999_000], frames)
    for b in range(2):
        single = t2.process(5_000_000, 4_999_000, frames[b])
        if single is None:
            assert batch[b] is None
        else:
            assert batch[b] is not None
            assert abs(single[0].x - batch[b][0].x) < 1e-9 and abs(single[0].y - batch[b][0].y) < 1e-9 and abs(single[0].rot - batch[b][0].rot) < 1e-9


def test_streaming_pose_batches_equal_the_blocking_call(oracle):
    """cb_detect_pose_gray_submit / _collect with two batches in flight (sizes 4, 2, 4): same detections, same Some/None, poses
    bit for bit those of cb_detect_pose_gray, and the same publish sequence; mixing with gray-only batches keeps the queue order."""
    from chalkydri_b200.capi import ChalkydriError, CB_ERR_STATE
    W, H = 1280, 720
    batches, gyros = [], []
    for i, n in enumerate((4, 2, 4)):
        f, _ = synth.render_batch(W, H, n, 1, seed=40 + i, edge_px=(90, 200))
        if i == 0:
            f[2] = 128                                   # heartbeat frame
        batches.append(np.ascontiguousarray(f))
        gyros.append([0.3, None, -0.7, 1.1][:n])         # a frame without a gyro reading in every batch
    c1, c2 = Comm(0.3), Comm(0.3)
    t1, _ = make_task(W, H, 4, c1)
    t2, _ = make_task(W, H, 4, c2)
    want, want_raw = [], []
    for f, g in zip(batches, gyros):
        want.append(t1.process_batch(7_000_000, [6_990_000] * len(f), f, gyro=g))
        want_raw.append([a.copy() for a in t1.last_batch])
    t2.submit_batch([6_990_000] * len(batches[0]), batches[0], gyro=gyros[0])
    got, got_raw = [], []
    for k in range(3):
        if k + 1 < 3:
            t2.submit_batch([6_990_000] * len(batches[k + 1]), batches[k + 1], gyro=gyros[k + 1])
            assert t2.detector.pending == 2
        got.append(t2.collect_batch(7_000_000))
        got_raw.append([a.copy() for a in t2.last_batch])
    solved = 0
    for w, g, wr, gr in zip(want, got, want_raw, got_raw):
        assert [r is None for r in w] == [r is None for r in g]
        assert wr[1].tolist() == gr[1].tolist() and wr[3].tolist() == gr[3].tolist() and wr[4].tolist() == gr[4].tolist()
        for b in range(len(wr[1])):
            assert wr[0][b, :wr[1][b]].tobytes() == gr[0][b, :gr[1][b]].tobytes()
            if wr[3][b]:
                assert wr[2][b].tobytes() == gr[2][b].tobytes()
                solved += 1
    assert solved >= 4
    assert [(m[0], m[1], m[2]) for m in c1.published] == [(m[0], m[1], m[2]) for m in c2.published]
    # a gray-only batch in front of a pose batch: the pose collect refuses it and leaves it queued
    t2.detector.submit(batches[1])
    t2.submit_batch([0, 0, 0, 0], batches[2], gyro=gyros[2])
    with pytest.raises(ChalkydriError) as e:
        t2.collect_batch(1)
    assert e.value.code == CB_ERR_STATE and t2.detector.pending == 2
    _, c = t2.detector.collect()
    assert c.tolist() == want_raw[1][1].tolist()
    last = t2.collect_batch(7_000_000)
    assert [r is None for r in last] == [r is None for r in want[2]] and t2.detector.pending == 0


def test_blocking_pose_call_with_a_pose_batch_in_flight_leaves_it_intact():
    """The blocking cb_detect_pose_gray shares the pose buffers with batches queued by cb_detect_pose_gray_submit: it must refuse
    (CB_ERR_STATE) BEFORE touching them, so the batch in flight still collects the poses it would have produced alone."""
    from chalkydri_b200.capi import ChalkydriError, CB_ERR_STATE
    W, H = 1280, 720
    f, _ = synth.render_batch(W, H, 4, 1, seed=40, edge_px=(90, 200))
    g = [0.3, 0.1, -0.7, 1.1]
    t1, _ = make_task(W, H, 4, Comm(0.3))
    want = t1.process_batch(7_000_000, [6_990_000] * 4, f, gyro=g)
    want_raw = [a.copy() for a in t1.last_batch]
    assert sum(r is not None for r in want) >= 2
    t2, _ = make_task(W, H, 4, Comm(0.3))
    t2.process_batch(7_000_000, [6_990_000] * 4, f, gyro=g)          # sizes the pose buffers, then the race the advisor described:
    for _ in range(3):
        t2.submit_batch([6_990_000] * 4, f, gyro=g)
        with pytest.raises(ChalkydriError) as e:
            t2.process_batch(7_000_000, [6_990_000] * 4, np.full_like(f, 128), gyro=[2.0] * 4)
        assert e.value.code == CB_ERR_STATE
        t2._device_path_ready = False
        with pytest.raises(ChalkydriError):
            t2._configure_device_path()                                 # cb_set_field / cb_set_camera are refused as well
        t2._device_path_ready = True
        got = t2.collect_batch(7_000_000)
        got_raw = t2.last_batch
        assert [r is None for r in got] == [r is None for r in want]
        assert got_raw[3].tolist() == want_raw[3].tolist()
        for b in range(4):
            if want_raw[3][b]:
                assert got_raw[2][b].tobytes() == want_raw[2][b].tobytes()


def test_graph_replay_of_the_one_frame_pose_call():
    """cb_detect_pose_gray on <= 4 frames joins its solve inside the chunk, so from the second call with a geometry on the whole detect ->
    un-project -> SQPnP sequence is one captured CUDA graph.  Six one-frame calls on one task (plain, capture, four replays; frames,
    gyro readings and a heartbeat frame changing under the graph) against a fresh task per frame (always the plain path): same
    detections, same Some/None, poses bit for bit; then a parameter change drops the graph and the next call is right again."""
    W, H = 1280, 720
    frames, _ = synth.render_batch(W, H, 6, 1, seed=41, edge_px=(90, 200))
    frames[2] = synth.render_batch(W, H, 1, 4, seed=42, edge_px=(70, 160))[0][0]
    frames[4] = 128
    gyros = [0.1, -0.4, 0.25, None, 0.0, 1.2]
    t, _ = make_task(W, H, 1, Comm(0.0))
    got = []
    for i in range(6):
        t.process_batch(1_000_000, [999_000], frames[i:i + 1], gyro=[gyros[i]])
        got.append([a.copy() for a in t.last_batch])
    for i in range(6):
        f, _ = make_task(W, H, 1, Comm(0.0))
        f.process_batch(1_000_000, [999_000], frames[i:i + 1], gyro=[gyros[i]])
        out, counts, poses, ok, ntags = f.last_batch
        assert counts.tolist() == got[i][1].tolist() and ok.tolist() == got[i][3].tolist() and ntags.tolist() == got[i][4].tolist(), i
        assert out[0, :counts[0]].tobytes() == got[i][0][0, :counts[0]].tobytes(), i
        if ok[0]:
            assert poses.tobytes() == got[i][2].tobytes(), i
    assert [int(g[3][0]) for g in got] == [1, 1, 0, 0, 0, 1]      # (frame 2: four randomly placed tags are no consistent scene -> None)
    # a solver parameter is baked into the graph: changing it drops the graph and the answer follows the new value
    from chalkydri_b200 import capi
    L = capi.lib()
    assert L.cb_sqpnp_set(t.detector.ctx, 0, 1e-8) == 0          # no Newton iteration at all
    t.process_batch(1_000_000, [999_000], frames[0:1], gyro=[gyros[0]])
    assert t.last_batch[3][0] == 0 or t.last_batch[2].tobytes() != got[0][2].tobytes()
    assert L.cb_sqpnp_set(t.detector.ctx, 15, 1e-8) == 0           # the defaults again (lib.rs:203-204)
    t.process_batch(1_000_000, [999_000], frames[0:1], gyro=[gyros[0]])
    t.process_batch(1_000_000, [999_000], frames[0:1], gyro=[gyros[0]])
    assert t.last_batch[2].tobytes() == got[0][2].tobytes()
