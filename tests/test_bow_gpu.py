"""GPU parity of the N1 / N4 rows against the CPU oracle, through the C ABI: DBoW2 transform,
median-Hamming main descriptor, per-edge information of the local-graph loader, float write-back."""
import numpy as np
import pytest

from oracle import pyoracle
from se2lam_b200 import _capi
from se2lam_b200.ba import LocalBA
from se2lam_b200.bow import Vocabulary, median_descriptor
from tests.bow_cases import make_features, make_observation_lists, make_voc
from tests.test_bow_oracle import loader_inputs
from tools import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,levels,ragged,levelsup,n", [(10, 4, False, 2, 1000), (10, 6, False, 4, 2000), (7, 5, True, 3, 700), (40, 2, False, 1, 300)])
def test_voc_transform(k, levels, ragged, levelsup, n):
    """k = 10, L = 6, levelsup = 4 is the ORBvoc shape KeyFrame::ComputeBoW assumes (KeyFrame.cpp:249-251): 1.1 M nodes."""
    voc = make_voc(k, levels, seed=levels, ragged=ragged)
    feats = make_features(voc, n, seed=k)
    w_o, wt_o, nd_o = pyoracle.voc_transform(voc, feats, levelsup)
    v = Vocabulary(voc["desc"], voc["child_ptr"], voc["children"], voc["word_id"], voc["weight"], voc["levels"])
    w_g, wt_g, nd_g = v.transform_features(feats, levelsup)
    np.testing.assert_array_equal(w_g, w_o); np.testing.assert_array_equal(wt_g, wt_o); np.testing.assert_array_equal(nd_g, nd_o)
    bow, fv = v.transform(feats, levelsup)
    assert abs(sum(bow.values()) - 1.0) < 1e-12 and sum(len(x) for x in fv.values()) == int((wt_o > 0).sum())
    assert list(bow) == sorted(bow) and all(nd_o[i] == node for node, idx in fv.items() for i in idx)


def test_median_descriptor():
    desc, ptr = make_observation_lists(M=500, seed=4, max_obs=60)
    i_o, m_o = pyoracle.median_descriptor(desc, ptr)
    i_g, m_g = median_descriptor(desc, ptr)
    np.testing.assert_array_equal(i_g, i_o); np.testing.assert_array_equal(m_g, m_o)
    big, bptr = make_observation_lists(M=3, seed=5, max_obs=300)      # > 48 KB of shared memory per map point
    i_o, m_o = pyoracle.median_descriptor(big, bptr)
    i_g, m_g = median_descriptor(big, bptr)
    np.testing.assert_array_equal(i_g, i_o); np.testing.assert_array_equal(m_g, m_o)


@pytest.mark.parametrize("cfg", ["C3", "C4"])
def test_edge_information_and_float_writeback(cfg):
    """N1: Omega per edge on the device from the loader's float data (Map.cpp:1024-1049), fed straight into
    se2gpu_ba_set_problem; after optimize() the estimates come back narrowed like Map::optimizeLocalGraph (Map.cpp:768-779)."""
    import copy
    import ctypes as C
    prob = synth.ba_config(cfg)
    li = loader_inputs(prob)
    ref = pyoracle.edge_information(**li)
    got = np.zeros((prob.E, 3))
    c = np.ascontiguousarray
    args = [c(li["view_mp"], np.float32), c(li["edge_pose"], np.int32), c(li["edge_point"], np.int32), c(li["octave"], np.int32),
            c(li["kf_Rcw"], np.float32), c(li["kf_twb_xy"], np.float32), c(li["mp_pos"], np.float32), c(li["level_sigma2"], np.float32)]
    _capi.check(_capi.lib().se2gpu_ba_build_information(prob.P, prob.L, prob.E, *[_capi.ptr(a) for a in args], 8, float(li["fx"]), 1e6, 1.0,
                                                        _capi.ptr(got), 0), "se2gpu_ba_build_information")
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-18)
    q = copy.copy(prob); q.info = got
    g = LocalBA.from_problem(q)
    g.optimize(5)
    qo = copy.copy(prob); qo.info = ref
    o = pyoracle.BAOracle(qo); o.optimize(5)
    pf = np.zeros((prob.P, 3), np.float32); lf = np.zeros((prob.L, 3), np.float32)
    _capi.check(_capi.lib().se2gpu_ba_get_f32(g.h, _capi.ptr(pf), _capi.ptr(lf)), "se2gpu_ba_get_f32")
    p64, l64 = g.get()
    np.testing.assert_array_equal(pf[:, :2], p64[:, :2].astype(np.float32)); np.testing.assert_array_equal(lf, l64.astype(np.float32))
    po, lo = o.writeback_f32()
    assert np.abs(pf - po).max() <= 2e-6 and np.abs(lf - lo).max() <= 2e-5      # float ulps of the 1e-8-close double estimates
    th = p64[:, 2].astype(np.float32).astype(np.float64)
    assert np.all(pf[:, 2] >= -np.pi - 1e-6) and np.all(np.abs(np.sin(pf[:, 2]) - np.sin(th)) < 1e-6)
