"""CPU pins of the SQPnP / un-projection restatement (oracle/sqpnp_oracle.cpp; reference crates/chalkydri_sqpnp/src/lib.rs).

The reference holds no test for the solver (PARITY UNPINNED), so the restatement is pinned against
independent implementations of the same mathematics: numpy for the linear-algebra building blocks the reference takes from
nalgebra, OpenCV for the OpenCVModel5 un-projection, and the ground truth of seeded problems for the whole solve."""
import numpy as np
import pytest

from tests import sqpnp_problems as sp


def test_sym_eigen9_against_numpy(oracle):
    """nalgebra symmetric_eigen (lib.rs:398) restated as cyclic Jacobi: eigenvalues / invariant subspaces of numpy's eigh."""
    rng = np.random.default_rng(0)
    for rank in (9, 5):                         # full rank, and the rank-5 Omega of a single planar tag
        b = rng.normal(size=(9, rank))
        a = b @ b.T
        d, v = oracle.sym_eigen9(a)
        w = np.linalg.eigvalsh(a)
        assert np.abs(np.sort(d) - w).max() < 1e-10 * max(1.0, w.max())
        assert np.abs(v.T @ v - np.eye(9)).max() < 1e-12                  # orthonormal
        assert np.abs(a @ v - v * d).max() < 1e-10 * max(1.0, w.max())    # A v = v diag(d)


def test_nearest_so3_against_numpy_svd(oracle):
    """nearest_so3 (lib.rs:42-59): U Vt with the third column of U flipped when det < 0."""
    rng = np.random.default_rng(1)
    for i in range(50):
        m = rng.normal(size=(3, 3))
        if i % 2:
            m[:, 0] *= -1                      # both chiralities
        u, _, vt = np.linalg.svd(m)
        r = u @ vt
        if np.linalg.det(r) < 0:
            u[:, 2] *= -1
            r = u @ vt
        got = oracle.nearest_so3(m)
        assert abs(np.linalg.det(got) - 1) < 1e-12
        assert np.abs(got - r).max() < 1e-9


def test_omega_is_the_quadratic_form_of_the_reprojection_error(oracle):
    """build_linear_system (lib.rs:124-180): r' Omega r equals the summed squared object-space error after the optimal
    translation has been eliminated, for any 9-vector r -- checked against a direct numpy evaluation."""
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(8, 3))
    pts -= pts.mean(0)
    bear = np.c_[rng.normal(scale=0.3, size=(8, 2)), np.ones(8)]
    om, _, _ = oracle.sqpnp_omega(pts, bear)
    assert np.abs(om - om.T).max() < 1e-12 and np.linalg.eigvalsh(om).min() > -1e-9
    # direct evaluation: residual of point i is P_i (R p_i + t) with P_i = I - b b'/b'b, t solved in the least-squares sense;
    # r is R column by column (R p = col0 x + col1 y + col2 z, lib.rs:141-147)
    P = [np.eye(3) - np.outer(b, b) / (b @ b) for b in bear]
    for _ in range(5):
        R = rng.normal(size=(3, 3))
        t = -np.linalg.solve(sum(P), sum(Pi @ (R @ p) for Pi, p in zip(P, pts)))
        direct = sum(float((R @ p + t) @ Pi @ (R @ p + t)) for Pi, p in zip(P, pts))
        r = R.T.reshape(-1)
        assert abs(float(r @ om @ r) - direct) < 1e-10 * max(1.0, direct)
    # and entry by entry against Q_rr - Q_rt Q_tt^-1 Q_rt' assembled with numpy
    Qrr, Qrt, Qtt = np.zeros((9, 9)), np.zeros((9, 3)), np.zeros((3, 3))
    for Pi, p in zip(P, pts):
        M = np.kron(p.reshape(1, 3), np.eye(3))          # 3x9: R p = M r
        Qrr += M.T @ Pi @ M
        Qrt += M.T @ Pi
        Qtt += Pi
    assert np.abs(om - (Qrr - Qrt @ np.linalg.inv(Qtt) @ Qrt.T)).max() < 1e-12


def test_unproject_against_opencv(oracle):
    """OpenCVModel5::unproject (crates/apriltags/src/lib.rs:316-321) == cv2.undistortPoints on the reference's calibrations."""
    cv2 = pytest.importorskip("cv2")
    from chalkydri_b200 import synth
    for W, H in ((1456, 1088), (1280, 720)):
        p = np.array(synth.scaled_calib(W, H), np.float64)
        K = np.array([[p[0], 0, p[2]], [0, p[1], p[3]], [0, 0, 1]])
        dist = p[4:9]                                                       # k1 k2 p1 p2 k3
        rng = np.random.default_rng(3)
        px = np.c_[rng.uniform(0.1 * W, 0.9 * W, 200), rng.uniform(0.1 * H, 0.9 * H, 200)]
        crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 100, 1e-14)
        want = cv2.undistortPointsIter(px.reshape(-1, 1, 2), K, dist, None, None, crit).reshape(-1, 2)
        for (u, v), w in zip(px, want):
            b = oracle.unproject_opencv5(p, float(u), float(v))
            assert b is not None
            assert np.abs(b[:2] / b[2] - w).max() < 1e-6


@pytest.mark.parametrize("two_tag_frac", [0.0, 1.0])
def test_noise_free_problems_recover_ground_truth(oracle, two_tag_frac):
    """solve_robot_pose (lib.rs:297-377) on exact projections: the robot pose comes back (median error < 1e-9 m).  Single
    planar tags leave a few problems without a start that reaches the global minimum (SQPnP nullspace caveat)."""
    tags, bearings, n_tags, r2c, gyro, truth = sp.make_problems(300, seed=5, two_tag_frac=two_tag_frac, noise_px=0.0)
    out, ok = oracle.sqpnp_batch(tags, bearings, n_tags, r2c, gyro)
    assert ok.mean() > 0.95
    err = np.linalg.norm(out["pos"][ok > 0, :2] - truth["pos"][ok > 0, :2], axis=1)
    assert np.median(err) < 1e-9
    # the remaining ~5-8 % are the reference algorithm's own misses (six starts from an implementation-defined basis of a
    # degenerate null space, then a yaw blend with the gyro), measured: 92 % / 94.5 % within 1e-6 m
    assert (err < 1e-6).mean() > 0.9


def test_noisy_problems_and_none_cases(oracle):
    tags, bearings, n_tags, r2c, gyro, truth = sp.make_problems(300, seed=6, two_tag_frac=0.5, noise_px=0.25)
    out, ok = oracle.sqpnp_batch(tags, bearings, n_tags, r2c, gyro)
    err = np.linalg.norm(out["pos"][ok > 0, :2] - truth["pos"][ok > 0, :2], axis=1)
    assert ok.mean() > 0.9 and np.median(err) < 0.15                 # a quarter pixel of corner noise: centimetres, not metres
    assert (out["std_devs"][ok > 0] > 0).all()
    # None: no tags; every point behind the camera; the batch form agrees with the single call
    assert oracle.sqpnp_solve_robot_pose(tags[0, :0], bearings[0, :0], r2c, 0.0) is None
    behind = bearings[0, :4].copy()
    behind[:] = (0.0, 0.0, -1.0)
    assert oracle.sqpnp_solve_robot_pose(tags[0, :1], behind, r2c, 0.0) is None
    k = int(np.nonzero(ok)[0][0])
    one = oracle.sqpnp_solve_robot_pose(tags[k, :n_tags[k]], bearings[k, :4 * n_tags[k]], r2c, float(gyro[k]))
    assert one is not None and one.tobytes() == out[k].tobytes()


def load_golden_sqpnp():
    import os
    from chalkydri_b200.capi import ISO_DTYPE
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sqpnp_64.npz"))
    tags = np.zeros(g["tags_t"].shape[:2], ISO_DTYPE)
    tags["t"], tags["q"] = g["tags_t"], g["tags_q"]
    r2c = np.zeros((), ISO_DTYPE)
    r2c["t"], r2c["q"] = g["r2c_t"], g["r2c_q"]
    return g, tags, r2c


def test_golden_fixture_sqpnp(oracle):
    """Regression pin: tests/golden/sqpnp_64.npz (tests/golden/make_golden.py) -- same Some/None, same poses."""
    g, tags, r2c = load_golden_sqpnp()
    out, ok = oracle.sqpnp_batch(tags, g["bearings"], g["n_tags"], r2c, g["gyro"])
    assert ok.tolist() == g["ok"].tolist() and ok.sum() >= 60
    m = ok > 0
    assert np.abs(out["pos"][m] - g["pos"][m]).max() < 1e-9 and np.abs(out["rot"][m] - g["rot"][m]).max() < 1e-9
    assert np.allclose(out["std_devs"][m], g["std_devs"][m], rtol=1e-9)


def test_two_tag_problems_against_opencv_sqpnp(oracle):
    """Second opinion on the whole solve: cv2.solvePnP(SOLVEPNP_SQPNP) -- Terzakis & Lourakis' own
    implementation -- on exact two-tag problems.  Both must recover the same camera pose; the reference's six-start variant
    misses the global minimum on a few per cent of planar layouts (nullspace caveat), OpenCV on fewer, so the bar is
    the median and the share of problems that agree, not the maximum."""
    cv2 = pytest.importorskip("cv2")
    tags, bearings, n_tags, r2c, gyro, truth = sp.make_problems(200, seed=9, two_tag_frac=1.0, noise_px=0.0)
    out, ok = oracle.sqpnp_batch(tags, bearings, n_tags, r2c, truth["yaw"])
    t_r2c = r2c["t"]
    diff = []
    for i in range(200):
        if not ok[i] or n_tags[i] != 2:
            continue
        world = np.concatenate([sp.CORNERS @ sp.qmat(tags["q"][i, s]).T + tags["t"][i, s] for s in range(2)])   # lib.rs:379-394
        good, rvec, tvec = cv2.solvePnP(world, bearings[i, :8, :2].copy(), np.eye(3), None, flags=cv2.SOLVEPNP_SQPNP)
        assert good
        R, _ = cv2.Rodrigues(rvec)
        robot_in_world = R.T @ (t_r2c - tvec.ravel())          # (world_to_cam)^-1 * robot_to_cam, translation part (lib.rs:328-337)
        diff.append(np.linalg.norm(robot_in_world - out["pos"][i]))
    diff = np.array(diff)
    assert len(diff) > 150
    assert np.median(diff) < 1e-8
    assert (diff < 1e-5).mean() > 0.85
