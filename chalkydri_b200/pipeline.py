"""Mirror of the Copper sink task `AprilTags` (/root/reference/crates/apriltags/src/lib.rs:166-379), row B0.

`AprilTags.new(config, comm)` reads the same config keys (`family`, `bits_corrected`, `cam_id`, `robot_to_cam`, `calib`;
lib.rs:227-233) and `process(now_us, frame_time_us, gray)` runs detect -> field lookup -> un-project -> one multi-tag
SQPnP -> publish, with the reference's skip rules (tags missing from field.json, lib.rs:306-308; corners that fail to
un-project, lib.rs:324-327) and the >5 ms empty heartbeat (lib.rs:365-376).  `comm` is any object with
`gyro_angle() -> float | None` and `publish(cam_id, tag_count, ts_us, pose, uncertainty)` (whacknet::Comm, whacknet/src/lib.rs:152-178).
The batched variant `process_batch` keeps detections on one device context and solves every frame's pose in one launch.
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np

from . import field
from .capi import ISO_DTYPE
from .detector import DetectorBuilder, FAMILY
from .solver import SIGN_FLIP_CONST, SqPnP, euler_angles


@dataclass
class RobotPose:             # whacknet/src/lib.rs:17-26
    x: float = 0.0
    y: float = 0.0
    rot: float = 0.0


@dataclass
class VisionUncertainty:     # whacknet/src/lib.rs (x, y, rot std-devs)
    x: float = 0.0
    y: float = 0.0
    rot: float = 0.0


MAX_DETECTIONS = 16          # crates/apriltags/src/lib.rs:42


class AprilTagDetections:
    """The task's detection payload (crates/apriltags/src/lib.rs:47-52): three parallel fixed-capacity lists (`CuArrayVec<_, 16>`)
    of tag id, tag pose (`CuPose<f32>`, a 4x4 f32 transform) and decision margin.  Serialised like the reference's serde impls
    (:69-121): a sequence of `(id, pose, decision_margin)` tuples.  Pushing beyond the capacity raises, like `ArrayVec::push`."""

    def __init__(self):
        self.ids: list[int] = []
        self.poses: list[np.ndarray] = []
        self.decision_margins: list[float] = []

    def push(self, tag_id: int, pose, decision_margin: float):
        if len(self.ids) >= MAX_DETECTIONS:
            raise OverflowError(f"AprilTagDetections holds at most {MAX_DETECTIONS} detections")
        self.ids.append(int(tag_id))
        self.poses.append(np.asarray(pose, np.float32).reshape(4, 4))
        self.decision_margins.append(float(np.float32(decision_margin)))

    @staticmethod
    def from_detections(dets, poses=None) -> "AprilTagDetections":
        """From a frame's detection records (DET_DTYPE); poses: per-detection 4x4 transforms (identity when absent)."""
        r = AprilTagDetections()
        for i, d in enumerate(dets[:MAX_DETECTIONS]):
            r.push(d["id"], np.eye(4) if poses is None else poses[i], d["decision_margin"])
        return r

    def filtered_by_decision_margin(self, threshold: float):
        """lib.rs:127-141: `(id, &pose, margin)` of the detections whose margin is strictly above the threshold, in order."""
        thr = np.float32(threshold)
        return ((i, p, m) for i, p, m in zip(self.ids, self.poses, self.decision_margins) if np.float32(m) > thr)

    def to_tuples(self):                     # impl Serialize, lib.rs:69-87
        return [(i, p.tolist(), m) for i, p, m in zip(self.ids, self.poses, self.decision_margins)]

    @staticmethod
    def from_tuples(seq) -> "AprilTagDetections":        # impl Deserialize, lib.rs:89-121
        r = AprilTagDetections()
        for i, p, m in seq:
            r.push(i, p, m)
        return r

    def __len__(self):
        return len(self.ids)


def calib_params(calib_json: str):
    m = json.loads(calib_json)["OpenCVModel5"]
    return np.array([m[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")], np.float64)


class AprilTags:
    def __init__(self, config: dict | None, comm, max_width=1600, max_height=1304, max_batch=1, device=0, field_path=None):
        self.comm = comm
        self.last_time = None
        self.tags = field.load(field_path)
        if config is not None:
            family = config.get("family", FAMILY)
            bits = int(config.get("bits_corrected", 3))
            self.cam_id = int(config["cam_id"])
            off = json.loads(config["robot_to_cam"])
            self.robot_to_cam = SqPnP.create_solver_camera_transform(off["x"], off["y"], off["z"], off["roll"], off["pitch"], off["yaw"])
            self.cam_params = calib_params(config["calib"])
            self.yaw = off["yaw"]
        else:                      # lib.rs:277-290
            family, bits, self.cam_id = FAMILY, 1, 255
            self.robot_to_cam = None
            self.cam_params = np.zeros(9)
            self.yaw = 0.0
        self.detector = (DetectorBuilder.default().add_family_bits(family, bits).device(device)
                         .capacity(max_width, max_height, max_batch, 64).build())
        self.solver = SqPnP(ctx=self.detector.ctx)

    @staticmethod
    def new(config, comm, **kw):
        return AprilTags(config, comm, **kw)

    def _correspondences(self, dets):
        world, cam = [], []
        for d in dets:
            tag = self.tags.get(int(d["id"]))
            if tag is None:
                continue
            bearings, ok = self.solver.unproject(self.cam_params, d["p"])
            if ok.all():
                world.append(tag)
                cam.append(bearings)
        return world, cam

    def _configure_device_path(self):
        """field layout + camera into the context once (cb_set_field / cb_set_camera)."""
        if getattr(self, "_device_path_ready", False):
            return
        from . import capi
        L = capi.lib()
        ids = np.array(sorted(self.tags), np.int32)
        poses = np.array([self.tags[int(i)] for i in ids], ISO_DTYPE)
        self.detector._check(L.cb_set_field(self.detector.ctx, capi.ptr(ids), capi.ptr(poses), len(ids)))
        r2c = None if self.robot_to_cam is None else np.ascontiguousarray(np.array(self.robot_to_cam, ISO_DTYPE))
        self.detector._check(L.cb_set_camera(self.detector.ctx, capi.ptr(np.ascontiguousarray(self.cam_params, np.float64)),
                                             None if r2c is None else capi.ptr(r2c)))
        self._device_path_ready = True

    def _gyro_array(self, gyro, B):
        if gyro is None:
            gy = self.comm.gyro_angle()
            gyro = [gy] * B
        return np.array([np.nan if v is None else float(v) for v in gyro], np.float64)

    def _publish_batch(self, now_us, frame_times_us, counts, poses, ok):
        """What `process` does with each frame's result (lib.rs:340-376): publish the pose, or the >5 ms heartbeat."""
        results = []
        for b in range(len(counts)):
            ts = now_us - int(frame_times_us[b])
            if ok[b]:
                rot = poses[b]["rot"].reshape(3, 3).T          # column-major like nalgebra
                pos, std = poses[b]["pos"], poses[b]["std_devs"]
                pose = RobotPose(pos[0], pos[1], euler_angles(rot)[2])
                unc = VisionUncertainty(std[0], std[1], std[2])
                self.comm.publish(self.cam_id, min(int(counts[b]), 255), ts, pose, unc)
                results.append((pose, unc))
                continue
            now_ms = now_us // 1000
            if self.last_time is None or (now_ms - self.last_time) > 5:
                self.comm.publish(self.cam_id, 0, ts, RobotPose(), VisionUncertainty())
                self.last_time = now_ms
            results.append(None)
        return results

    def process_batch(self, now_us: int, frame_times_us, grays: np.ndarray, gyro=None):
        """`process` for a batch of frames in ONE library call (cb_detect_pose_gray): detections never leave the device between
        detect, field lookup, un-projection and the per-frame SQPnP.  gyro: per-frame yaw (None / NaN = no reading, like
        comm.gyro_angle() == None); default: comm.gyro_angle() for every frame.  Publishes and returns per frame what `process`
        would: (RobotPose, VisionUncertainty) or None."""
        from . import capi
        from .capi import DET_DTYPE, POSE_DTYPE
        self._configure_device_path()
        L = capi.lib()
        grays = np.ascontiguousarray(grays)
        B, H, W = grays.shape
        gy = self._gyro_array(gyro, B)
        out = np.zeros((B, 64), DET_DTYPE); counts = np.zeros(B, np.int32)
        poses = np.zeros(B, POSE_DTYPE); ok = np.zeros(B, np.uint8); ntags = np.zeros(B, np.int32)
        try:
            self.detector._check(L.cb_detect_pose_gray(self.detector.ctx, capi.ptr(grays), W, H, W, H * W, B, capi.ptr(gy), SIGN_FLIP_CONST,
                                                       capi.ptr(out), capi.ptr(counts), capi.ptr(poses), capi.ptr(ok), capi.ptr(ntags)))
        except capi.ChalkydriError as e:
            self._overflow_to_heartbeat(e, counts, ok)
        self.last_batch = (out, counts, poses, ok, ntags)
        return self._publish_batch(now_us, frame_times_us, counts, poses, ok)

    # ---- continuous feed: submit batch k+1, then collect batch k (cb_detect_pose_gray_submit / _collect) ----
    def submit_batch(self, frame_times_us, grays: np.ndarray, gyro=None):
        """Enqueue a batch (at most two in flight; keep `grays` alive and unchanged until the matching collect_batch)."""
        from . import capi
        self._configure_device_path()
        L = capi.lib()
        if grays.dtype != np.uint8 or grays.ndim != 3 or not grays.flags.c_contiguous:
            raise ValueError("grays must be a C-contiguous [B,H,W] uint8 array")
        B, H, W = grays.shape
        gy = self._gyro_array(gyro, B)
        self.detector._check(L.cb_detect_pose_gray_submit(self.detector.ctx, capi.ptr(grays), W, H, W, H * W, B, capi.ptr(gy), SIGN_FLIP_CONST))
        if not hasattr(self, "_inflight"):
            self._inflight = []
        self._inflight.append((grays, [int(t) for t in frame_times_us]))

    def collect_batch(self, now_us: int):
        """Wait for the oldest submitted batch, publish and return per frame what `process` would."""
        from . import capi
        from .capi import DET_DTYPE, POSE_DTYPE
        L = capi.lib()
        inflight = getattr(self, "_inflight", [])
        B = inflight[0][0].shape[0] if inflight else 1                 # nothing in flight: the library reports CB_ERR_STATE
        out = np.zeros((B, 64), DET_DTYPE); counts = np.zeros(B, np.int32)
        poses = np.zeros(B, POSE_DTYPE); ok = np.zeros(B, np.uint8); ntags = np.zeros(B, np.int32)
        try:
            self.detector._check(L.cb_detect_pose_gray_collect(self.detector.ctx, capi.ptr(out), capi.ptr(counts), capi.ptr(poses), capi.ptr(ok),
                                                               capi.ptr(ntags)))
        except capi.ChalkydriError as e:
            if e.code == capi.CB_ERR_OVERFLOW and inflight:
                self._overflow_to_heartbeat(e, counts, ok)     # the batch has left the queue; its frames publish heartbeats
            else:
                if e.code != capi.CB_ERR_STATE and inflight:       # a failed batch has left the queue; a refused collect has not
                    inflight.pop(0)
                raise
        times = inflight.pop(0)[1] if inflight else []
        self.last_batch = (out, counts, poses, ok, ntags)
        return self._publish_batch(now_us, times, counts, poses, ok)

    def _overflow_to_heartbeat(self, e, counts, ok):
        """A frame so cluttered that a fixed-size device table overflowed (CB_ERR_OVERFLOW) must not take the task down -- upstream's
        detector has no such limit.  The batch is treated as "nothing detected" (heartbeats go out) and counted; every other error
        propagates."""
        from . import capi
        if e.code != capi.CB_ERR_OVERFLOW:
            raise e
        self.overflow_batches = getattr(self, "overflow_batches", 0) + 1
        counts[:] = 0
        ok[:] = 0

    def process(self, now_us: int, frame_time_us: int, gray: np.ndarray):
        from . import capi
        try:
            out, counts = self.detector.detect_batch(np.ascontiguousarray(gray)[None])
            dets = out[0, :counts[0]]
        except capi.ChalkydriError as e:
            if e.code != capi.CB_ERR_OVERFLOW:
                raise
            self.overflow_batches = getattr(self, "overflow_batches", 0) + 1
            dets = []
        if len(dets) > 0:
            world, cam = self._correspondences(dets)
            gyro = self.comm.gyro_angle()
            if gyro is not None and world:
                r2c = self.robot_to_cam if self.robot_to_cam is not None else np.array(((0, 0, 0), (1, 0, 0, 0)), ISO_DTYPE)
                res = self.solver.solve_robot_pose(np.array(world, ISO_DTYPE), np.concatenate(cam), r2c, gyro, SIGN_FLIP_CONST)
                if res is not None:
                    rot, pos, std = res
                    pose = RobotPose(pos[0], pos[1], euler_angles(rot)[2])
                    unc = VisionUncertainty(std[0], std[1], std[2])
                    self.comm.publish(self.cam_id, min(len(dets), 255), now_us - frame_time_us, pose, unc)
                    return pose, unc
        now_ms = now_us // 1000
        if self.last_time is None or (now_ms - self.last_time) > 5:
            self.comm.publish(self.cam_id, 0, now_us - frame_time_us, RobotPose(), VisionUncertainty())
            self.last_time = now_ms
        return None
