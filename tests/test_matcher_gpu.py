"""GPU parity of the matchers against the CPU oracle (bit-exact integer/index work), through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle
from tools import synth
from se2lam_b200.matcher import FrameView, ORBmatcher
from tests.matcher_cases import GRID, make_bow_case, make_frame_pair, make_projection_case

pytestmark = pytest.mark.gpu


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5000, 32), dtype=np.uint8); b = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    got = ORBmatcher.DescriptorDistance(a, b)
    ref = np.unpackbits(a ^ b, axis=1).sum(axis=1)
    np.testing.assert_array_equal(got, ref)
    assert ORBmatcher.DescriptorDistance(a[0], b[0]) == pyoracle.descriptor_distance(a[0], b[0])


@pytest.mark.parametrize("seed,ratio,win", [(1, 0.9, 20), (4, 0.9, 20), (5, 0.6, 35), (6, 0.95, 8)])
def test_match_by_window(seed, ratio, win):
    f1, f2, prev = make_frame_pair(seed=seed)
    n_o, m_o, prev_o = pyoracle.match_by_window(f1["kp"], f1["desc"], f2["kp"], f2["desc"], prev, GRID, win, 1, 0, 8, ratio)
    prev_g = prev.copy()
    n_g, m_g = ORBmatcher(ratio).MatchByWindow(FrameView(f1["kp"], f1["desc"]), FrameView(f2["kp"], f2["desc"]), prev_g, win)
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)
    np.testing.assert_array_equal(prev_g, prev_o)


def test_match_by_window_on_real_extractions():
    """Frames from the extractor itself (shifted copy of the same texture), as Track::mTrack feeds the matcher."""
    from se2lam_b200.orb import ORBextractor
    img1 = synth.orb_frame(1000)
    img2 = np.roll(img1, (3, 5), axis=(0, 1))
    e = ORBextractor(1000, 1.2, 8)
    k1, d1 = e(img1); k2, d2 = e(img2)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32).copy()
    n_o, m_o, prev_o = pyoracle.match_by_window(k1, d1, k2, d2, prev, GRID, 20, 1, 0, 8, 0.9)
    prev_g = prev.copy()
    n_g, m_g = ORBmatcher(0.9).MatchByWindow(FrameView(k1, d1), FrameView(k2, d2), prev_g, 20)
    assert n_o > 300
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)
    np.testing.assert_array_equal(prev_g, prev_o)


@pytest.mark.parametrize("seed", [2, 7])
def test_match_by_projection(seed):
    a = make_projection_case(seed=seed)["args"]
    n_o, m_o = pyoracle.match_by_projection(**a)
    n_g, m_g = ORBmatcher(a["nnratio"]).MatchByProjection(FrameView(a["kfkp"], a["kfdesc"]), a["kf_observed"], a["mp_valid"], a["mp_uv"],
                                                         a["mp_octave"], a["mp_desc"], a["win_size"], a["level_offset"])
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)


@pytest.mark.parametrize("seed,mp_only,ori", [(3, True, True), (8, False, True), (9, True, False)])
def test_search_by_bow(seed, mp_only, ori):
    k1, k2 = make_bow_case(seed=seed)
    n_o, m_o = pyoracle.search_by_bow(k1, k2, mp_only, 0.6, ori)
    n_g, m_g = ORBmatcher(0.6, ori).SearchByBoW(k1, k2, mp_only)
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)


def test_empty_inputs():
    f1, f2, prev = make_frame_pair(seed=1, n=200)
    n, m = ORBmatcher(0.9).MatchByWindow(FrameView(f1["kp"], f1["desc"]), FrameView(f2["kp"][:0], f2["desc"][:0]), prev.copy(), 20)
    assert n == 0 and np.all(m == -1)
