"""CPU: the ORB oracle reproduces the committed golden vectors (which were pinned against cv2 4.13 by
oracle/pin_orb_against_cv2.py in the build container) and behaves sanely on edge cases."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "orb_golden.npz"))

CASES = {
    "synth1000": lambda: synth.orb_frame(1000), "synth1001": lambda: synth.orb_frame(1001),
    "constant": lambda: synth.orb_adversarial("constant"), "noise": lambda: synth.orb_adversarial("noise"),
    "lowcontrast": lambda: synth.orb_adversarial("lowcontrast"), "gradient": lambda: synth.orb_adversarial("gradient"),
    "small_320x240": lambda: synth.orb_frame(5, 320, 240), "odd_501x377": lambda: synth.orb_frame(6, 501, 377),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    img = CASES[name]()
    if GOLD[name + "_img"].size:
        np.testing.assert_array_equal(img, GOLD[name + "_img"])   # generator drift guard
    kps, desc = pyoracle.OrbOracle().extract(img)
    assert kps.tobytes() == GOLD[name + "_kps"].tobytes()
    assert desc.tobytes() == GOLD[name + "_desc"].tobytes()


def test_constructor_tables():
    t = pyoracle.OrbOracle().tables()
    assert t["features_per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert t["features_per_level"].sum() == 1000


def test_keypoint_invariants():
    kps, desc = pyoracle.OrbOracle().extract(synth.orb_frame(1002))
    assert len(kps) == 1000 and desc.shape == (1000, 32)
    assert np.all(np.diff(kps["octave"]) >= 0)                       # levels concatenated in order
    assert np.all((kps["angle"] >= 0) & (kps["angle"] < 360))
    assert np.all(kps["class_id"] == -1)
    lvl0 = kps[kps["octave"] == 0]
    assert np.all((lvl0["x"] >= 16) & (lvl0["x"] < 640 - 16) & (lvl0["y"] >= 16) & (lvl0["y"] < 480 - 16))


def test_introselect_port_matches_std_nth_element(tmp_path):
    """se2lam_b200/csrc/introselect.h (host+device source) against this toolchain's std::nth_element."""
    exe = tmp_path / "introselect_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "native", "introselect_check.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "identical" in out.stdout
