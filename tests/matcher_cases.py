"""Shared synthetic cases for the matcher tests (CPU oracle tests and GPU parity tests)."""
import numpy as np

from oracle.pyoracle import KP_DTYPE

f32 = np.float32
GRID = (f32(0.0), f32(0.0), f32(f32(64) / f32(640.0)), f32(f32(48) / f32(480.0)))


def random_keypoints(rng, n):
    kp = np.zeros(n, KP_DTYPE)
    octave = np.minimum(rng.geometric(0.35, n) - 1, 7)
    scale = (f32(1.2) ** octave).astype(f32)
    kp["x"] = (rng.integers(16, 500, n).astype(f32) * scale).clip(0, 639.5)
    kp["y"] = (rng.integers(16, 380, n).astype(f32) * scale).clip(0, 479.5)
    kp["octave"] = octave
    kp["angle"] = rng.uniform(0, 360, n).astype(f32)
    kp["size"] = 31 * scale
    kp["response"] = rng.integers(20, 120, n)
    kp["class_id"] = -1
    return kp


def make_frame_pair(seed=1, n=900, shift=(6.0, -4.0), drot=7.0, flip=12):
    """frame2 = frame1 moved by `shift` px with a few descriptor bits flipped, shuffled, plus clutter."""
    rng = np.random.default_rng(seed)
    kp1 = random_keypoints(rng, n)
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    kp2 = kp1[perm].copy()
    kp2["x"] = (kp2["x"] + f32(shift[0]) + rng.normal(0, 1.5, n).astype(f32)).astype(f32)
    kp2["y"] = (kp2["y"] + f32(shift[1]) + rng.normal(0, 1.5, n).astype(f32)).astype(f32)
    kp2["angle"] = np.mod(kp2["angle"] + f32(drot) + rng.normal(0, 2, n).astype(f32), 360).astype(f32)
    d2 = d1[perm].copy()
    for i in range(n):
        bits = rng.integers(0, 256, rng.integers(0, 2 * flip))
        for b in bits:
            d2[i, b // 8] ^= np.uint8(1 << (b % 8))
    # duplicates / near-duplicates so that the "already matched better" and steal-back paths fire
    dup = rng.choice(n, 120, replace=False)
    kp2[dup[:60]] = kp2[dup[60:]]
    d2[dup[:60]] = d2[dup[60:]] ^ rng.integers(0, 2, (60, 32), dtype=np.uint8)
    prev = np.stack([kp1["x"], kp1["y"]], axis=1).astype(f32).copy()
    return dict(kp=kp1, desc=d1), dict(kp=kp2, desc=d2), prev


def brute_candidates(kp2, x, y, r, min_level, max_level, minX=0.0, minY=0.0):
    """GetFeaturesInArea by explicit grid construction (Frame.cpp:64-77, 222-286) in numpy float32."""
    invW, invH = GRID[2], GRID[3]
    posx = np.floor(f32((kp2["x"] - f32(minX)) * invW) + f32(0.5)).astype(int)
    posy = np.floor(f32((kp2["y"] - f32(minY)) * invH) + f32(0.5)).astype(int)
    x, y, r = f32(x), f32(y), f32(r)
    x0 = max(0, int(np.floor(f32(f32(f32(x - f32(minX)) - r) * invW)))); x1 = min(63, int(np.ceil(f32(f32(f32(x - f32(minX)) + r) * invW))))
    y0 = max(0, int(np.floor(f32(f32(f32(y - f32(minY)) - r) * invH)))); y1 = min(47, int(np.ceil(f32(f32(f32(y - f32(minY)) + r) * invH))))
    if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
        return []
    out = []
    for ix in range(x0, x1 + 1):
        for iy in range(y0, y1 + 1):
            for i in np.flatnonzero((posx == ix) & (posy == iy)):
                if not (min_level == -1 and max_level == -1):
                    if kp2["octave"][i] < min_level or kp2["octave"][i] > max_level:
                        continue
                if abs(f32(kp2["x"][i] - x)) > r or abs(f32(kp2["y"][i] - y)) > r:
                    continue
                out.append(int(i))
    return out


def make_projection_case(seed=2, n=900, nmp=700):
    rng = np.random.default_rng(seed)
    kp = random_keypoints(rng, n)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    src = rng.integers(0, n, nmp)
    mp_uv = np.stack([kp["x"][src] + rng.normal(0, 4, nmp), kp["y"][src] + rng.normal(0, 4, nmp)], axis=1).astype(f32)
    mp_octave = np.clip(kp["octave"][src] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    mp_desc = desc[src].copy()
    flips = rng.integers(0, 256, (nmp, 20))
    for i in range(nmp):
        for b in flips[i, :rng.integers(0, 20)]:
            mp_desc[i, b // 8] ^= np.uint8(1 << (b % 8))
    mp_valid = (rng.random(nmp) > 0.1).astype(np.uint8)
    kf_observed = (rng.random(n) < 0.15).astype(np.uint8)
    return dict(args=dict(kfkp=kp, kfdesc=desc, kf_observed=kf_observed, mp_valid=mp_valid, mp_uv=mp_uv, mp_octave=mp_octave,
                          mp_desc=mp_desc, grid=GRID, win_size=15, level_offset=2, nnratio=0.6))


def make_bow_case(seed=3, n=800, nnodes=120):
    rng = np.random.default_rng(seed)

    def kf(desc, angle, node_of):
        nodes = np.unique(node_of)
        ptr = [0]; feat = []
        for nd in nodes:
            idx = np.flatnonzero(node_of == nd)
            feat.extend(idx.tolist()); ptr.append(len(feat))
        return dict(angle=angle.astype(f32), desc=desc, has_mp=(rng.random(len(desc)) < 0.8).astype(np.uint8), node=nodes.astype(np.int32),
                    ptr=np.asarray(ptr, np.int32), feat=np.asarray(feat, np.int32))
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    a1 = rng.uniform(0, 360, n)
    node1 = rng.integers(0, nnodes, n) * 3
    perm = rng.permutation(n)
    d2 = d1[perm].copy()
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, 24)):
            d2[i, b // 8] ^= np.uint8(1 << (b % 8))
    a2 = np.mod(a1[perm] + 11 + rng.normal(0, 3, n), 360)
    node2 = node1[perm].copy()
    moved = rng.random(n) < 0.1
    node2[moved] = rng.integers(0, nnodes + 20, moved.sum()) * 3 + 1     # nodes that exist in only one KF
    return kf(d1, a1, node1), kf(d2, a2, node2)
