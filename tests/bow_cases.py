"""Synthetic vocabulary trees / observation lists for the bag-of-words tests (the real ORBvoc.bin is not available offline)."""
import numpy as np


def make_voc(k=10, levels=4, seed=0, ragged=False):
    """k-ary tree with `levels` levels below the root, nodes in breadth-first order like DBoW2's m_nodes (root = 0).
    Children are noisy copies of their parent (as k-means centres would be: about 32 of the 256 bits flipped), leaves carry
    word ids in creation order and idf-like weights (a few are 0 = stopped words). ragged=True prunes random subtrees and
    varies the branching factor, so some leaves are shallow. Vectorised: the ORBvoc shape (k = 10, L = 6) has 1.1 M nodes."""
    rng = np.random.default_rng(seed)
    desc = [rng.integers(0, 256, (1, 32), dtype=np.uint8)]
    nkids = []                       # per node, in node order
    depth = [np.zeros(1, np.int32)]
    frontier = desc[0]
    for lvl in range(1, levels + 1):
        npar = len(frontier)
        if ragged:
            cnt = rng.integers(2, k + 1, npar)
            if lvl > 1:
                cnt[rng.random(npar) < 0.15] = 0                  # those parents stay (shallow) leaves
        else:
            cnt = np.full(npar, k)
        nkids.append(cnt)
        par = np.repeat(np.arange(npar), cnt)
        n = len(par)
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        frontier = frontier[par] ^ flips
        desc.append(frontier); depth.append(np.full(n, lvl, np.int32))
    nkids.append(np.zeros(len(frontier), np.int64))
    desc = np.concatenate(desc); depth = np.concatenate(depth)
    cnt_all = np.concatenate(nkids)
    n = len(desc)
    child_ptr = np.zeros(n + 1, np.int32); child_ptr[1:] = np.cumsum(cnt_all)
    children = np.arange(1, n, dtype=np.int32)                    # breadth-first: the children of node i are consecutive
    leaf = cnt_all == 0
    word_id = -np.ones(n, np.int32); word_id[leaf] = np.arange(int(leaf.sum()))
    weight = np.zeros(n)
    w = rng.uniform(0.5, 9.0, int(leaf.sum())); w[rng.random(len(w)) < 0.03] = 0.0
    weight[leaf] = w
    return dict(desc=desc, child_ptr=child_ptr, children=children, word_id=word_id, weight=weight, levels=levels, depth=depth)


def make_features(voc, n, seed=1):
    """descriptors near random leaves (plus some pure noise) so that descents are non-trivial and ties occur"""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(voc["word_id"] >= 0)
    f = voc["desc"][rng.choice(leaves, n)].copy()
    for i in range(n):
        for b in rng.integers(0, 256, int(rng.integers(0, 60))):
            f[i, b // 8] ^= np.uint8(1 << (b % 8))
    f[: n // 10] = rng.integers(0, 256, (n // 10, 32), dtype=np.uint8)
    return f


def make_observation_lists(M=300, seed=2, max_obs=40):
    rng = np.random.default_rng(seed)
    ptr = [0]; rows = []
    for m in range(M):
        N = int(rng.integers(0, max_obs + 1)) if m % 17 else int(rng.integers(1, 4))
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        for _ in range(N):
            d = base.copy()
            for b in rng.integers(0, 256, int(rng.integers(0, 40))):
                d[b // 8] ^= np.uint8(1 << (b % 8))
            rows.append(d)
        ptr.append(len(rows))
    return np.stack(rows), np.asarray(ptr, np.int32)
