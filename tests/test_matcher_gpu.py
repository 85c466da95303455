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


def test_owned_context_and_rounds():
    """An ORBmatcher that owns its context (se2gpu_matcher_create) gives the same results; the speculative resolve needs a
    handful of rounds, not one per query, and does not take the sequential fallback on ordinary frames."""
    f1, f2, prev = make_frame_pair(seed=4)
    n_o, m_o, prev_o = pyoracle.match_by_window(f1["kp"], f1["desc"], f2["kp"], f2["desc"], prev, GRID, 20, 1, 0, 8, 0.9)
    mt = ORBmatcher(0.9, max_queries=1024, max_db=1024)
    for _ in range(3):       # the context is reused call after call
        prev_g = prev.copy()
        n_g, m_g = mt.MatchByWindow(FrameView(f1["kp"], f1["desc"]), FrameView(f2["kp"], f2["desc"]), prev_g, 20)
        assert n_g == n_o
        np.testing.assert_array_equal(m_g, m_o)
        np.testing.assert_array_equal(prev_g, prev_o)
    rounds, fallback = mt.last_rounds()
    assert not fallback and 1 <= rounds <= 32, (rounds, fallback)
    a = make_projection_case(seed=7)["args"]
    n_o, m_o = pyoracle.match_by_projection(**a)
    n_g, m_g = mt.MatchByProjection(FrameView(a["kfkp"], a["kfdesc"]), a["kf_observed"], a["mp_valid"], a["mp_uv"], a["mp_octave"],
                                    a["mp_desc"], a["win_size"], a["level_offset"])
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)
    k1, k2 = make_bow_case(seed=8)
    n_o, m_o = pyoracle.search_by_bow(k1, k2, False, 0.6, True)
    n_g, m_g = ORBmatcher(0.6, True, max_queries=1024, max_db=1024).SearchByBoW(k1, k2, False)
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)
    with pytest.raises(Exception):      # capacity is checked, not silently exceeded
        big1, big2, bprev = make_frame_pair(seed=1, n=1500)
        mt.MatchByWindow(FrameView(big1["kp"], big1["desc"]), FrameView(big2["kp"], big2["desc"]), bprev, 20)


def test_steal_chain_takes_the_exact_fallback():
    """40 queries that each beat the previous claim on ONE database keypoint (a chain of steals, :331-334): more simultaneous
    claims than the shared-memory claim table holds, so the exact sequential kernel must take over - same result."""
    rng = np.random.default_rng(5)
    n = 200
    f1, f2, prev = make_frame_pair(seed=9, n=n)
    kp1, d1, kp2, d2 = f1["kp"].copy(), f1["desc"].copy(), f2["kp"].copy(), f2["desc"].copy()
    kp2["x"][0], kp2["y"][0], kp2["octave"][0] = 300.0, 200.0, 0
    for q in range(40):
        kp1["x"][q], kp1["y"][q], kp1["octave"][q] = 300.0 + 0.1 * q, 200.0, 0
        kp1["angle"][q] = np.float32((float(kp2["angle"][0]) - 7.0) % 360.0)      # same rotation bin as the true pairs
        d1[q] = d2[0]
        for b in rng.choice(256, 40 - q, replace=False):     # distance to keypoint 0 of frame 2 decreases with q
            d1[q, b // 8] ^= np.uint8(1 << (b % 8))
    prev = np.stack([kp1["x"], kp1["y"]], axis=1).astype(np.float32).copy()
    n_o, m_o, prev_o = pyoracle.match_by_window(kp1, d1, kp2, d2, prev, GRID, 20, 1, 0, 8, 0.9)
    mt = ORBmatcher(0.9, max_queries=256, max_db=256)
    prev_g = prev.copy()
    n_g, m_g = mt.MatchByWindow(FrameView(kp1, d1), FrameView(kp2, d2), prev_g, 20)
    assert (m_o[:40] == 0).sum() == 1            # only the last claimant keeps keypoint 0
    assert n_g == n_o
    np.testing.assert_array_equal(m_g, m_o)
    np.testing.assert_array_equal(prev_g, prev_o)
    assert mt.last_rounds()[1], "expected the claim table to overflow on this input"


def test_device_resident_extract_then_match():
    """N2: extractor -> matcher without leaving HBM (reference Track.cpp:129-132): the matcher consumes the extractor's device
    keypoint / descriptor buffers and its device-side keypoint counts; only the final matches come back."""
    import torch
    from se2lam_b200.orb import ORBextractor
    from se2lam_b200._capi import KP_DTYPE
    dev = torch.device("cuda", 0)
    img1 = synth.orb_frame(1003)
    img2 = np.roll(img1, (-4, 6), axis=(0, 1))
    nf = 1000
    e = ORBextractor(nf, 1.2, 8, max_batch=2)
    d_img = torch.from_numpy(np.stack([img1, img2])).to(dev)
    d_kps = torch.zeros(2 * nf * 28, dtype=torch.uint8, device=dev)
    d_desc = torch.zeros(2 * nf * 32, dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(2, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    e.extract_device(d_img, 2, 480, 640, d_kps, d_desc, d_counts, stream=s)
    mt = ORBmatcher(0.9, max_queries=nf, max_db=nf)
    d_prev = torch.zeros(2 * nf, dtype=torch.float32, device=dev)
    d_m = torch.zeros(nf, dtype=torch.int32, device=dev)
    d_nm = torch.zeros(1, dtype=torch.int32, device=dev)
    kp1_ptr, kp2_ptr = d_kps.data_ptr(), d_kps.data_ptr() + nf * 28
    de1_ptr, de2_ptr = d_desc.data_ptr(), d_desc.data_ptr() + nf * 32
    ORBmatcher.KeypointsToPointsDevice(kp1_ptr, nf, d_prev, d_n=d_counts.data_ptr(), stream=s)
    mt.MatchByWindowDevice(kp1_ptr, de1_ptr, nf, kp2_ptr, de2_ptr, nf, d_prev, FrameView(None, None).grid(), 20, d_m, d_nm,
                           d_n1=d_counts.data_ptr(), d_n2=d_counts.data_ptr() + 4, stream=s)
    torch.cuda.synchronize()
    c = d_counts.cpu().numpy()
    kps = d_kps.cpu().numpy().view(KP_DTYPE).reshape(2, nf)
    desc = d_desc.cpu().numpy().reshape(2, nf, 32)
    k1, k2, de1, de2 = kps[0, :c[0]], kps[1, :c[1]], desc[0, :c[0]], desc[1, :c[1]]
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32).copy()
    n_o, m_o, prev_o = pyoracle.match_by_window(k1, de1, k2, de2, prev, GRID, 20, 1, 0, 8, 0.9)
    assert n_o > 300
    assert int(d_nm.item()) == n_o
    np.testing.assert_array_equal(d_m.cpu().numpy()[:c[0]], m_o)
    np.testing.assert_array_equal(d_prev.cpu().numpy().reshape(-1, 2)[:c[0]], prev_o)
