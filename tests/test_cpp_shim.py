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
    chi2_sum, n_edges = struct.unpack_from("di", buf, off); off += 12
    (aborted,) = struct.unpack_from("i", buf, off)
    o = pyoracle.BAOracle(prob)
    n_o, _ = o.optimize(10)
    po, lo = o.get()
    assert done == n_o and aborted == 0
    np.testing.assert_allclose(poses, po, atol=1e-8)
    np.testing.assert_allclose(pts, lo, atol=1e-7)
    # sum of the per-edge NON-robust chi2 (what removeOutlierChi2 thresholds) at the final estimate, against a numpy restatement
    assert n_edges == prob.E + prob.O
    Rcb = np.asarray(prob.Tcb[:9]).reshape(3, 3); tcb = np.asarray(prob.Tcb[9:])
    tot = 0.0
    for e in range(prob.E):
        x, y, th = po[prob.edge_pose[e]]
        c, s = np.cos(th), np.sin(th)
        lc = Rcb @ (np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]]) @ (lo[prob.edge_point[e]] - np.array([x, y, 0.0]))) + tcb
        err = prob.fx * lc[:2] / lc[2] + np.array([prob.cx, prob.cy]) - prob.uv[e]
        w = prob.info[e]
        tot += err @ np.array([[w[0], w[1]], [w[1], w[2]]]) @ err
    for k in range(prob.O):
        a, b = po[prob.odo_i[k]], po[prob.odo_j[k]]
        c, s = np.cos(a[2]), np.sin(a[2])
        d = b[:2] - a[:2]
        err = np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], b[2] - a[2]]) - prob.odo_meas[k]
        w = prob.odo_info[k]
        tot += err @ np.array([[w[0], w[1], w[2]], [w[1], w[3], w[4]], [w[2], w[4], w[5]]]) @ err
    assert abs(chi2_sum - tot) <= 1e-6 * tot


def compile_g2o_demo(tmp_path):
    build.build_lib()
    exe = str(tmp_path / "g2o_binding_demo")
    libdir = os.path.dirname(build.LIB_PATH)
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "native", "mock_g2o"),
           os.path.join(ROOT, "tests", "native", "g2o_binding_demo.cpp"), "-o", exe, "-L", libdir, "-lse2gpu", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_g2o_binding_compiles_against_the_g2o_api(tmp_path):
    compile_g2o_demo(tmp_path)


@pytest.mark.gpu
def test_g2o_binding_matches_oracle(tmp_path):
    """se2gpu::G2oGpuLevenberg (include/se2lam/g2o_gpu_levenberg.h) as g2o's OptimizationAlgorithm: the (mock) SparseOptimizer
    drives it with solve(0), solve(1), ... exactly like real g2o, one se2gpu_ba_optimize_from slice per call."""
    from oracle import pyoracle
    exe = compile_g2o_demo(tmp_path)
    prob = synth.ba_config("C3")
    img = synth.orb_frame(1005, 64, 48)
    Rbc, tbc = synth.default_Tbc()
    fin, fout = str(tmp_path / "g_in.bin"), str(tmp_path / "g_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("ii", 64, 48)); f.write(img.tobytes())
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
    done, on_gpu = struct.unpack_from("ii", buf, 0); off = 8
    poses = np.frombuffer(buf, "f8", 3 * prob.P, off).reshape(-1, 3); off += 24 * prob.P
    pts = np.frombuffer(buf, "f8", 3 * prob.L, off).reshape(-1, 3); off += 24 * prob.L
    lam, trials = struct.unpack_from("di", buf, off)
    o = pyoracle.BAOracle(prob)
    n_o, st_o = o.optimize(10)
    po, lo = o.get()
    assert on_gpu == 1 and done == n_o
    np.testing.assert_allclose(poses, po, atol=1e-8)
    np.testing.assert_allclose(pts, lo, atol=1e-7)
    assert abs(lam - st_o["lambda"][-1]) <= 1e-6 * st_o["lambda"][-1] and trials == st_o["trials"][-1]


def compile_matcher_demo(tmp_path):
    build.build_lib()
    exe = str(tmp_path / "matcher_demo")
    libdir = os.path.dirname(build.LIB_PATH)
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "native", "stub"),
           os.path.join(ROOT, "tests", "native", "matcher_demo.cpp"), "-o", exe, "-L", libdir, "-lse2gpu", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_matcher_shim_compiles_and_links(tmp_path):
    compile_matcher_demo(tmp_path)


@pytest.mark.gpu
def test_matcher_shim_matches_oracle(tmp_path):
    """include/se2lam/ORBmatcher.h driven like Track.cpp:129-132, LocalMapper.cpp:117-118 and GlobalMapper.cpp:274-276."""
    from oracle import pyoracle
    from tests.matcher_cases import GRID, make_bow_case, make_frame_pair, make_projection_case
    exe = compile_matcher_demo(tmp_path)
    f1, f2, prev = make_frame_pair(seed=4)
    a = make_projection_case(seed=7)["args"]
    uvc = a["mp_uv"]      # KeyFrame::inImgBound is part of the flattened mp_valid (ORBmatcher.cpp:398-399): apply it to the oracle's input
    a["mp_valid"] = (a["mp_valid"].astype(bool) & (uvc[:, 0] >= 0) & (uvc[:, 0] <= 640) & (uvc[:, 1] >= 0) & (uvc[:, 1] <= 480)).astype(np.uint8)
    k1, k2 = make_bow_case(seed=3)
    fin, fout = str(tmp_path / "m_in.bin"), str(tmp_path / "m_out.bin")

    def put_frame(f, kp, desc):
        f.write(struct.pack("i", len(kp))); f.write(np.ascontiguousarray(kp).tobytes()); f.write(np.ascontiguousarray(desc, np.uint8).tobytes())
    with open(fin, "wb") as f:
        put_frame(f, f1["kp"], f1["desc"]); put_frame(f, f2["kp"], f2["desc"])
        put_frame(f, a["kfkp"], a["kfdesc"])
        f.write(np.ascontiguousarray(a["kf_observed"], np.uint8).tobytes())
        M = len(a["mp_valid"])
        f.write(struct.pack("i", M)); f.write(np.ascontiguousarray(a["mp_valid"], np.uint8).tobytes())
        f.write(np.ascontiguousarray(a["mp_uv"], np.float32).tobytes()); f.write(np.ascontiguousarray(a["mp_octave"], np.int32).tobytes())
        f.write(np.ascontiguousarray(a["mp_desc"], np.uint8).tobytes())
        for k in (k1, k2):
            kp = np.zeros(len(k["desc"]), pyoracle.KP_DTYPE); kp["angle"] = k["angle"]
            put_frame(f, kp, k["desc"])
            f.write(np.ascontiguousarray(k["has_mp"], np.uint8).tobytes())
            f.write(struct.pack("i", len(k["node"]))); f.write(k["node"].astype(np.int32).tobytes()); f.write(k["ptr"].astype(np.int32).tobytes())
            f.write(k["feat"].astype(np.int32).tobytes())
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    buf = open(fout, "rb").read()
    off = 0
    nm, n = struct.unpack_from("ii", buf, off); off += 8
    m12 = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    pv = np.frombuffer(buf, np.float32, 2 * n, off).reshape(n, 2); off += 8 * n
    prev0 = np.stack([f1["kp"]["x"], f1["kp"]["y"]], axis=1).astype(np.float32)
    n_o, m_o, prev_o = pyoracle.match_by_window(f1["kp"], f1["desc"], f2["kp"], f2["desc"], prev0, GRID, 20, 1, 0, 8, 0.9)
    assert nm == n_o and n == len(m_o)
    np.testing.assert_array_equal(m12, m_o); np.testing.assert_array_equal(pv, prev_o)
    nm, n = struct.unpack_from("ii", buf, off); off += 8
    mp = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    a2 = dict(a); a2["nnratio"] = 0.6
    n_o, m_o = pyoracle.match_by_projection(**a2)
    assert nm == n_o
    np.testing.assert_array_equal(mp, m_o)
    nm, sz = struct.unpack_from("ii", buf, off); off += 8
    pairs = np.frombuffer(buf, np.int32, 2 * sz, off).reshape(sz, 2); off += 8 * sz
    (dd,) = struct.unpack_from("i", buf, off)
    n_o, m_o = pyoracle.search_by_bow(k1, k2, True, 0.6, True)
    assert nm == n_o == sz
    ref = np.flatnonzero(m_o >= 0)
    np.testing.assert_array_equal(pairs[:, 0], ref); np.testing.assert_array_equal(pairs[:, 1], m_o[ref])
    assert dd == int(np.unpackbits(k1["desc"][0] ^ k2["desc"][0]).sum())
