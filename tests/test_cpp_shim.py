"""The drop-in C++ headers (include/se2lam/*.h) compile, link against libse2gpu.so and — on a GPU — produce the
oracle's results when driven exactly like the reference's call sites (tests/native/shim_demo.cpp)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from se2lam_b200 import build
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compile_demo(tmp_path):
    build.build_lib()
    exe = str(tmp_path / "shim_demo")
    libdir = os.path.dirname(build.LIB_PATH)
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "native", "shim_demo.cpp"),
           "-o", exe, "-L", libdir, "-lse2gpu", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_shim_headers_compile_and_link(tmp_path):
    compile_demo(tmp_path)


@pytest.mark.gpu
def test_shim_matches_oracle(tmp_path):
    from oracle import pyoracle
    exe = compile_demo(tmp_path)
    img = synth.orb_frame(1005)
    prob = synth.ba_config("C3")
    Rbc, tbc = synth.default_Tbc()
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("ii", 640, 480)); f.write(img.tobytes())
        f.write(struct.pack("iiiii", prob.P, prob.L, prob.E, prob.O, 10))
        for a, dt in ((prob.poses, "f8"), (prob.fixed, "u1"), (prob.points, "f8"), (prob.edge_pose, "i4"), (prob.edge_point, "i4"),
                      (prob.uv, "f8"), (prob.info, "f8"), (prob.odo_i, "i4"), (prob.odo_j, "i4"), (prob.odo_meas, "f8"), (prob.odo_info, "f8")):
            f.write(np.ascontiguousarray(a, dt).tobytes())
        f.write(np.array([prob.fx, prob.cx, prob.cy], "f8").tobytes())
        f.write(np.concatenate([Rbc.reshape(-1), tbc]).astype("f8").tobytes())
        f.write(struct.pack("d", prob.huber_delta))
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    buf = open(fout, "rb").read()
    off = 0
    (N,) = struct.unpack_from("i", buf, off); off += 4
    kps = np.frombuffer(buf, pyoracle.KP_DTYPE, N, off); off += 28 * N
    desc = np.frombuffer(buf, np.uint8, 32 * N, off).reshape(N, 32); off += 32 * N
    levels, sf, untouched = struct.unpack_from("ifi", buf, off); off += 12
    ko, do_ = pyoracle.OrbOracle().extract(img)
    assert N == len(ko) and kps.tobytes() == ko.tobytes() and desc.tobytes() == do_.tobytes()
    assert levels == 8 and abs(sf - 1.2) < 1e-6 and untouched == 3
    (done,) = struct.unpack_from("i", buf, off); off += 4
    poses = np.frombuffer(buf, "f8", 3 * prob.P, off).reshape(-1, 3); off += 24 * prob.P
    pts = np.frombuffer(buf, "f8", 3 * prob.L, off).reshape(-1, 3); off += 24 * prob.L
    (aborted,) = struct.unpack_from("i", buf, off)
    o = pyoracle.BAOracle(prob)
    n_o, _ = o.optimize(10)
    po, lo = o.get()
    assert done == n_o and aborted == 0
    np.testing.assert_allclose(poses, po, atol=1e-8)
    np.testing.assert_allclose(pts, lo, atol=1e-7)
