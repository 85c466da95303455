"""Pins oracle/undistort_oracle.cpp against cv2.undistort (needs the cv2 wheel of the build container; not run on the GPU box)
and writes tests/golden/undistort_golden.npz (small frames, so the fixture stays a few hundred KB).

    python oracle/pin_undistort_against_cv2.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

CASES = {
    # name: (K, dist, w, h)
    "tum_like_5": (np.array([[520.9, 0, 325.1], [0, 521.0, 249.7], [0, 0, 1]], np.float32),
                   np.array([0.2312, -0.7849, -0.0033, -0.0001, 0.9172], np.float32), 640, 480),
    "wide_5": (np.array([[231.976, 0, 326.923], [0, 232.157, 227.838], [0, 0, 1]], np.float32),
               np.array([-0.207406, 0.032194, 0.001120, 0.000859, 0.0], np.float32), 640, 480),
    "skew_4": (np.array([[400.3, 0.7, 160.2], [0, 399.1, 119.6], [0, 0, 1]], np.float32),
               np.array([-0.3, 0.1, 0.001, -0.002], np.float32), 320, 240),
    "rational_8_odd": (np.array([[700.0, 0, 250.5], [0, 701.0, 188.5], [0, 0, 1]], np.float32),
                       np.array([0.05, 0.0, 0.0, 0.0, 0.0, 0.01, 0.0, 0.0], np.float32), 501, 377),
    "prism_12": (np.array([[300.0, 0, 80.0], [0, 300.0, 60.0], [0, 0, 1]], np.float32),
                 np.array([-0.1, 0.02, 0.001, 0.001, 0.0, 0.0, 0.0, 0.0, 0.002, -0.001, 0.001, 0.0005], np.float32), 160, 120),
    "none_5": (np.array([[500.0, 0, 80.0], [0, 500.0, 60.0], [0, 0, 1]], np.float32), np.zeros(5, np.float32), 160, 120),
}


def main():
    rng = np.random.default_rng(11)
    golden = {}
    ok = True
    for name, (K, D, w, h) in CASES.items():
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img[h // 4:h // 2, w // 4:w // 2] = 255          # a saturated patch: rounding at the clamp
        ref = cv2.undistort(img, K, D)
        mine = pyoracle.undistort(img, K, D)
        m1c, m2c = cv2.initUndistortRectifyMap(K.astype(np.float64), D.astype(np.float64), None, K.astype(np.float64), (w, h), cv2.CV_16SC2)
        nd = int((mine != ref).sum())
        print(f"{name}: {w}x{h} differing pixels {nd}")
        ok &= nd == 0
        if w * h <= 320 * 240:
            golden[name + "_K"] = K; golden[name + "_D"] = D; golden[name + "_img"] = img; golden[name + "_out"] = ref
    if not ok:
        raise SystemExit("UNDISTORT ORACLE NOT PINNED")
    out = os.path.join(ROOT, "tests", "golden", "undistort_golden.npz")
    np.savez_compressed(out, **golden)
    print("ALL PINNED; wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
