"""CPU: the bag-of-words oracle and the Omega / write-back oracle against independent numpy restatements."""
import numpy as np
import pytest

from oracle import pyoracle
from tests.bow_cases import make_features, make_observation_lists, make_voc
from tools import synth


def _ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


@pytest.mark.parametrize("k,levels,ragged,levelsup", [(10, 3, False, 1), (10, 3, False, 4), (6, 4, True, 2), (3, 5, True, 4)])
def test_voc_transform_against_bruteforce(k, levels, ragged, levelsup):
    """TemplatedVocabulary::transform (TemplatedVocabulary.h:1220-1262) restated with explicit python loops."""
    voc = make_voc(k, levels, seed=k + levels, ragged=ragged)
    feats = make_features(voc, 120, seed=levels)
    word, weight, node = pyoracle.voc_transform(voc, feats, levelsup)
    cp, ch = voc["child_ptr"], voc["children"]
    for f in range(len(feats)):
        cur, lvl = 0, 0
        nid_level = voc["levels"] - levelsup
        nid = 0 if nid_level <= 0 else -1
        while cp[cur + 1] > cp[cur]:
            lvl += 1
            kids = ch[cp[cur]:cp[cur + 1]]
            dist = [_ham(feats[f], voc["desc"][c]) for c in kids]
            cur = int(kids[int(np.argmin(dist))])          # argmin returns the FIRST minimum == the reference's strict <
            if lvl == nid_level:
                nid = cur
        assert word[f] == voc["word_id"][cur] and weight[f] == voc["weight"][cur] and node[f] == nid


def test_median_descriptor_against_bruteforce():
    """MapPoint::updateMainKFandDescriptor (MapPoint.cpp:245-267) with numpy sort."""
    desc, ptr = make_observation_lists(M=120, seed=3, max_obs=25)
    idx, med = pyoracle.median_descriptor(desc, ptr)
    for m in range(len(ptr) - 1):
        d = desc[ptr[m]:ptr[m + 1]]
        N = len(d)
        if N == 0:
            assert idx[m] == -1
            continue
        D = np.array([[0 if i == j else _ham(d[i], d[j]) for j in range(N)] for i in range(N)])
        meds = np.sort(D, axis=1)[:, int(0.5 * (N - 1))]
        assert idx[m] == int(np.argmin(meds)) and med[m] == int(meds.min())


def loader_inputs(prob, seed=0):
    """What Map::loadLocalGraph reads per edge (Map.cpp:1004-1037) for a synthetic window: float Tcw rotation and Twb per
    keyframe, float landmark positions, float camera-frame measurements mViewMPs and the keypoint octave."""
    rng = np.random.default_rng(seed)
    Rcb = np.asarray(prob.Tcb[:9]).reshape(3, 3); tcb = np.asarray(prob.Tcb[9:])
    P = prob.P
    Rcw = np.zeros((P, 3, 3), np.float32); twb = np.zeros((P, 2), np.float32)
    for i in range(P):
        x, y, th = prob.poses[i]
        c, s = np.cos(th), np.sin(th)
        Rcw[i] = (Rcb @ np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])).astype(np.float32)
        twb[i] = (x, y)
    mp = prob.points.astype(np.float32)
    view = np.zeros((prob.E, 3), np.float32)
    for e in range(prob.E):
        p = prob.edge_pose[e]
        d = mp[prob.edge_point[e]].astype(np.float64) - np.array([twb[p, 0], twb[p, 1], 0.0])
        view[e] = (Rcw[p].astype(np.float64) @ d + tcb).astype(np.float32)
    octave = np.minimum(rng.geometric(1 - 1 / 1.44, prob.E) - 1, 7).astype(np.int32)
    sigma2 = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32) ** 2
    return dict(view_mp=view, edge_pose=prob.edge_pose, edge_point=prob.edge_point, octave=octave, kf_Rcw=Rcw.reshape(P, 9), kf_twb_xy=twb,
                mp_pos=mp, level_sigma2=sigma2.astype(np.float32), fx=np.float32(prob.fx))


def test_edge_information_against_numpy():
    """Omega of Map.cpp:1024-1049: the C++ oracle against the numpy restatement of tools/synth.py (np.linalg.inv)."""
    prob = synth.ba_config("C3")
    li = loader_inputs(prob)
    info = pyoracle.edge_information(**li)
    Rcb = np.asarray(prob.Tcb[:9]).reshape(3, 3); tcb = np.asarray(prob.Tcb[9:])
    for e in range(0, prob.E, 37):
        p, j = prob.edge_pose[e], prob.edge_point[e]
        Rcw = li["kf_Rcw"][p].reshape(3, 3).astype(np.float64)
        lc = li["view_mp"][e].astype(np.float64)
        fx = float(li["fx"])
        zi = 1.0 / lc[2]
        Jpi = np.array([[fx * zi, 0, -fx * lc[0] * zi * zi], [0, fx * zi, -fx * lc[1] * zi * zi]])
        M = Jpi @ Rcw
        d = li["mp_pos"][j].astype(np.float64) - np.array([li["kf_twb_xy"][p, 0], li["kf_twb_xy"][p, 1], 0.0])
        S = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        Jr = (M @ S)[:, :2]; Jz = -M[:, 2:3]
        Sig = float(np.float32(1.0 / 1e6)) * Jr @ Jr.T + float(np.float32(1.0)) * Jz @ Jz.T + np.eye(2) * float(li["level_sigma2"][li["octave"][e]])
        Om = np.linalg.inv(Sig)
        np.testing.assert_allclose(info[e], [Om[0, 0], 0.5 * (Om[0, 1] + Om[1, 0]), Om[1, 1]], rtol=1e-11, atol=1e-15)
