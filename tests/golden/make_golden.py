"""Generates the committed fixtures under tests/golden/ from the oracle (run once; outputs are committed).

There is no reference binary or golden vector to import, so these fixtures pin the ORACLE
against drift; the oracle itself is pinned by ground truth + cv2 cross-checks in tests/test_oracle_*.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chalkydri_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def detector_c1():
    seed = 1
    im, _ = synth.render_frame(1280, 720, 4, seed=seed, edge_px=(60, 150))
    dets, taps = po.detect(im, taps=True)
    np.savez_compressed(os.path.join(HERE, "detector_c1.npz"), seed=seed, frame=im, ids=dets["id"], hamming=dets["hamming"],
                        corners=dets["p"], margin=dets["decision_margin"], H=dets["H"],
                        thresh_white_bits=np.packbits(taps["thresh"] == 255), thresh_black_bits=np.packbits(taps["thresh"] == 0),
                        npoints=taps["npoints"], nquads=taps["nquads"], quads=taps["quads"]["p"])


def detector_biterr():
    """four tags with 0 / 1 / 2 / 3 inverted data bits (tests/frames.py): the oracle's lists for bits_corrected 0..3"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import frames as fr
    im, truth = fr.bit_error_frame()
    rec = {"frame": im, "true_ids": truth["ids"], "true_hamming": truth["hamming"]}
    for bits in range(4):
        dets = po.detect(im, po.default_params(bits_corrected=bits))
        rec[f"ids_b{bits}"], rec[f"hamming_b{bits}"], rec[f"corners_b{bits}"] = dets["id"], dets["hamming"], dets["p"]
        rec[f"margin_b{bits}"] = dets["decision_margin"]
    np.savez_compressed(os.path.join(HERE, "detector_biterr.npz"), **rec)


def detector_cv2_pin():
    """INDEPENDENT pin (not produced by the oracle): OpenCV 4.13's aruco module carries its own port of the UMich AprilTag quad
    detector (CORNER_REFINE_APRILTAG: threshold -> union-find -> gradient clusters -> fit_quad).  Its corners on the c1 frame,
    un-decimated, are what quad_decimate = 1 / refine_edges = 0 of this detector must reproduce (measured: <= 0.05 px), and
    its decode with maxCorrectionBits = k is bits_corrected = k on the bit-error frame."""
    import cv2
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import frames as fr
    dic = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11)
    prm = cv2.aruco.DetectorParameters()
    prm.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_APRILTAG
    prm.aprilTagQuadDecimate = 0.0
    rec = {"cv2_version": np.array(cv2.__version__)}
    for name, seed in (("c1", 1), ("s2", 2), ("s3", 3)):
        im, _ = synth.render_frame(1280, 720, 4, seed=seed, edge_px=(60, 150))
        corners, ids, _ = cv2.aruco.ArucoDetector(dic, prm).detectMarkers(im)
        order = np.argsort(ids.ravel())
        rec[f"{name}_seed"] = seed
        rec[f"{name}_ids"] = ids.ravel()[order].astype(np.int32)
        rec[f"{name}_corners"] = np.array([corners[k].reshape(4, 2) for k in order], np.float64)      # AprilTag pixel convention
    im, _ = fr.bit_error_frame()
    for k in range(4):
        dic.maxCorrectionBits = k
        prm.errorCorrectionRate = 1.0
        _, ids, _ = cv2.aruco.ArucoDetector(dic, prm).detectMarkers(im)
        rec[f"biterr_ids_k{k}"] = np.sort(ids.ravel()).astype(np.int32) if ids is not None else np.zeros(0, np.int32)
    np.savez_compressed(os.path.join(HERE, "detector_cv2_pin.npz"), **rec)


def sqpnp_64():
    """64 seeded problems (half with two tags, 0.25 px corner noise): inputs and the oracle's Some/None + poses"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import sqpnp_problems as sp
    tags, bearings, n_tags, r2c, gyro, truth = sp.make_problems(64, seed=11, two_tag_frac=0.5, noise_px=0.25)
    out, ok = po.sqpnp_batch(tags, bearings, n_tags, r2c, gyro)
    np.savez_compressed(os.path.join(HERE, "sqpnp_64.npz"), tags_t=tags["t"], tags_q=tags["q"], bearings=bearings, n_tags=n_tags,
                        r2c_t=r2c["t"], r2c_q=r2c["q"], gyro=gyro, ok=ok, rot=out["rot"], pos=out["pos"], std_devs=out["std_devs"],
                        true_pos=truth["pos"], true_yaw=truth["yaw"])


def cat_96x72():
    """a 96x72 RGB frame through every CAT stage"""
    rng = np.random.default_rng(12)
    gray, _ = synth.render_frame(96, 72, 1, seed=12, edge_px=(40, 60))
    rgb = np.clip(gray[..., None].astype(int) + rng.integers(-6, 7, (72, 96, 3)), 0, 255).astype(np.uint8)
    color = po.cat_calc_otsu(rgb)
    xy, n = po.cat_detect_corners(color)
    lines, m = po.cat_check_edges(color, xy)
    labels, sizes = po.cat_connected_components(color)
    np.savez_compressed(os.path.join(HERE, "cat_96x72.npz"), rgb=rgb, otsu=color, thresh=po.cat_thresh(rgb), corners=xy, lines=lines,
                        labels=labels, sizes=sizes)


if __name__ == "__main__":
    detector_c1()
    detector_biterr()
    detector_cv2_pin()
    sqpnp_64()
    cat_96x72()
    print("golden fixtures written")
