"""Self-consistency pins of the BA oracle (no golden data exists for g2o: SURVEY.md section 8c)."""
import numpy as np
import pytest

from oracle import ba_numpy, pyoracle
from tools import synth


@pytest.fixture(scope="module")
def small():
    return synth.ba_window(n_kf=6, n_lm=60, seed=3)


def test_analytic_jacobians_match_central_differences(small):
    o = pyoracle.BAOracle(small)
    for e in range(0, small.E, 7):
        err, Ji, Jj = o.edge_xyz(e)
        i, j = small.edge_pose[e], small.edge_point[e]
        np.testing.assert_allclose(err, ba_numpy.xyz_error(small, small.poses[i], small.points[j], e), rtol=0, atol=1e-9)
        Jn_i = ba_numpy._numjac(lambda x: ba_numpy.xyz_error(small, x, small.points[j], e), small.poses[i].copy())
        Jn_j = ba_numpy._numjac(lambda x: ba_numpy.xyz_error(small, small.poses[i], x, e), small.points[j].copy())
        np.testing.assert_allclose(Ji, Jn_i, rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(Jj, Jn_j, rtol=1e-6, atol=1e-5)
    for k in range(small.O):
        err, Ji, Jj = o.edge_odo(k)
        i, j = small.odo_i[k], small.odo_j[k]
        np.testing.assert_allclose(err, ba_numpy.odo_error(small, small.poses[i], small.poses[j], k), atol=1e-12)
        Jn_i = ba_numpy._numjac(lambda x: ba_numpy.odo_error(small, x, small.poses[j], k), small.poses[i].copy())
        Jn_j = ba_numpy._numjac(lambda x: ba_numpy.odo_error(small, small.poses[i], x, k), small.poses[j].copy())
        np.testing.assert_allclose(Ji, Jn_i, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(Jj, Jn_j, rtol=1e-6, atol=1e-7)


def test_chi2_matches_numpy(small):
    o = pyoracle.BAOracle(small)
    assert o.chi2() == pytest.approx(ba_numpy.total_chi2(small, small.poses, small.points), rel=1e-12)


def test_schur_solve_equals_dense_full_system_solve(small):
    o = pyoracle.BAOracle(small)
    lin = o.linearize()
    H, b, hidx, lidx, nf, nl = ba_numpy.build_full_system(small, small.poses, small.points)
    assert nf == o.nf
    np.testing.assert_allclose(lin["Hpp"], H[:3 * nf, :3 * nf], rtol=1e-5, atol=1e-5 * np.abs(H).max())
    lam = 1e-5 * np.abs(np.diag(H)).max()
    dx = np.linalg.solve(H + lam * np.eye(len(b)), b)
    ss = o.schur_solve(lam)
    assert ss["ok"] == 1
    scale = np.abs(dx).max()
    np.testing.assert_allclose(ss["dx_p"], dx[:3 * nf], rtol=0, atol=2e-5 * scale)
    for j in range(small.L):
        if lidx[j] >= 0:
            np.testing.assert_allclose(ss["dx_l"][j], dx[3 * nf + 3 * lidx[j]:3 * nf + 3 * lidx[j] + 3], rtol=0, atol=2e-5 * scale)
    # Schur with analytic Jacobians vs its own dense assembly: exact consistency to 1e-10
    n = 3 * nf
    Hfull = np.zeros((n + 3 * small.L, n + 3 * small.L)); bfull = np.concatenate([lin["bp"], lin["bl"].reshape(-1)])
    Hfull[:n, :n] = lin["Hpp"]
    free = -np.ones(small.P, int); free[np.flatnonzero(small.fixed == 0)] = np.arange(nf)
    for j in range(small.L):
        Hfull[n + 3 * j:n + 3 * j + 3, n + 3 * j:n + 3 * j + 3] = lin["Hll"][j]
    for e in range(small.E):
        a = free[small.edge_pose[e]]
        if a >= 0:
            j = small.edge_point[e]
            Hfull[3 * a:3 * a + 3, n + 3 * j:n + 3 * j + 3] += lin["Hpl"][e]
            Hfull[n + 3 * j:n + 3 * j + 3, 3 * a:3 * a + 3] += lin["Hpl"][e].T
    act = np.ones(len(bfull), bool)
    for j in range(small.L):
        if lidx[j] < 0:
            act[n + 3 * j:n + 3 * j + 3] = False
    dx2 = np.linalg.solve((Hfull + lam * np.eye(len(bfull)))[np.ix_(act, act)], bfull[act])
    got = np.concatenate([ss["dx_p"], ss["dx_l"].reshape(-1)])[act]
    np.testing.assert_allclose(got, dx2, rtol=0, atol=1e-10 * max(1.0, np.abs(dx2).max()))


def test_lm_trajectory_matches_numpy_restatement(small):
    o = pyoracle.BAOracle(small)
    n, st, tp, tl = o.optimize(5, trace=True)
    poses, points, stats = ba_numpy.lm_optimize(small, 5)
    assert n == len(stats)
    for k in range(n):
        assert st["trials"][k] == stats[k]["trials"]
        assert st["chi2_after"][k] == pytest.approx(stats[k]["chi2_after"], rel=1e-6)
        assert st["lambda"][k] == pytest.approx(stats[k]["lam"], rel=1e-4)
    np.testing.assert_allclose(tp[-1], poses, atol=1e-6)
    np.testing.assert_allclose(tl[-1], points, atol=1e-5)


def test_noise_free_converges_to_ground_truth_and_gauge_is_fixed():
    prob = synth.ba_window(n_kf=8, n_lm=120, seed=5, noise=False)
    rng = np.random.default_rng(0)
    prob.points = prob.points + rng.normal(0, 0.02, prob.points.shape)
    prob.poses[1:] += rng.normal(0, [0.01, 0.01, 0.003], (prob.P - 1, 3))
    o = pyoracle.BAOracle(prob)
    chi0 = o.chi2()
    n, st = o.optimize(15)
    poses, pts = o.get()
    assert st["chi2_after"][-1] < 1e-6 * chi0
    np.testing.assert_array_equal(poses[0], prob.poses[0])          # fixed pose never moves
    np.testing.assert_allclose(poses, prob.gt_poses.astype(np.float32).astype(np.float64), atol=2e-4)


@pytest.mark.parametrize("cfg", ["C1", "C3"])
def test_configs_reduce_cost(cfg):
    prob = synth.ba_config(cfg)
    o = pyoracle.BAOracle(prob)
    n, st = o.optimize(10)
    assert n >= 1
    assert st["chi2_after"][-1] < st["chi2_before"][0]
    assert np.all(st["chi2_after"] <= st["chi2_before"] + 1e-9)
