"""cv::undistort (reference src/Frame.cpp:22) — CPU checks: the oracle against the golden vectors produced from cv2.undistort
(oracle/pin_undistort_against_cv2.py), and the host-side map builder of the product library against the oracle's map."""
import os

import numpy as np
import pytest

from oracle import pyoracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort_golden.npz")


def cases():
    z = np.load(GOLD)
    names = sorted({k[:-4] for k in z.files if k.endswith("_img")})
    return [(n, z[n + "_K"], z[n + "_D"], z[n + "_img"], z[n + "_out"]) for n in names]


@pytest.mark.parametrize("name,K,D,img,out", cases(), ids=[c[0] for c in cases()])
def test_oracle_matches_cv2_golden(name, K, D, img, out):
    assert np.array_equal(pyoracle.undistort(img, K, D), out)


def test_identity_distortion_is_a_copy_inside_the_frame():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    K = np.array([[300, 0, 80], [0, 300, 60], [0, 0, 1]], np.float32)
    assert np.array_equal(pyoracle.undistort(img, K, np.zeros(5, np.float32)), img)


@pytest.mark.parametrize("name,K,D,img,out", cases(), ids=[c[0] for c in cases()])
def test_product_map_builder_matches_oracle(name, K, D, img, out):
    """se2gpu_orb_set_undistort's host-side map (double arithmetic, stripes, LU inverse) == the oracle's, also for a
    benchmark-size frame; needs the built library but no GPU."""
    from se2lam_b200 import _capi
    for (w, h) in ((img.shape[1], img.shape[0]), (640, 480)):
        m1o, m2o = pyoracle.undistort_map(K, D, w, h)
        m1 = np.zeros((h, w, 2), np.int16); m2 = np.zeros((h, w), np.uint16)
        Kf = np.ascontiguousarray(K, np.float32).reshape(9); Df = np.ascontiguousarray(D, np.float32).ravel()
        rc = _capi.lib().se2gpu_orb_debug_undistort_map(Kf.ctypes.data, Df.ctypes.data, len(Df), w, h, m1.ctypes.data, m2.ctypes.data)
        assert rc == 0
        assert np.array_equal(m1, m1o) and np.array_equal(m2, m2o)
