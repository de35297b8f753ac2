"""WPILib-format field layout loader (mirror of /root/reference/crates/apriltags/src/field_layout.rs:18-44).

`load()` returns {id: iso} with iso = (t[3], q[4] = w,x,y,z normalised like UnitQuaternion::from_quaternion).
The 32 tag poses of the reference's field.json are committed as a fixture under tests/golden/field.json
(data file; the reference loads it from the working directory, field_layout.rs:19).
"""
from __future__ import annotations

import json
import os

import numpy as np

ISO_DTYPE = np.dtype([("t", "<f8", (3,)), ("q", "<f8", (4,))])

DEFAULT_FIELD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "field.json")


def load(path: str | None = None) -> dict:
    with open(path or DEFAULT_FIELD) as f:
        layout = json.load(f)
    tags = {}
    for tag in layout["tags"]:
        tr = tag["pose"]["translation"]
        q = tag["pose"]["rotation"]["quaternion"]
        quat = np.array([q["W"], q["X"], q["Y"], q["Z"]], np.float64)
        quat = quat / np.sqrt((quat * quat).sum())
        iso = np.zeros((), ISO_DTYPE)
        iso["t"] = (tr["x"], tr["y"], tr["z"])
        iso["q"] = quat
        tags[int(tag["ID"])] = iso
    return tags
