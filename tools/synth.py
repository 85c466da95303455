"""Deterministic synthetic inputs for the two hot paths - TEST / BENCH INFRASTRUCTURE.

Used by tests/, bench.py, __graft_entry__.smoke() and the golden-vector generators; not part of the product package.
Pure numpy; no GPU, no oracle.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


# --------------------------------------------------------------------------------------------
# ORB frames
# --------------------------------------------------------------------------------------------
def orb_frame(seed: int, w: int = 640, h: int = 480) -> np.ndarray:
    """Blocky-texture frame: 8x nearest-upsampled random 80x60 base + U(-8,8) noise (section 8(d))."""
    rng = np.random.default_rng(seed)
    bw, bh = (w + 7) // 8, (h + 7) // 8
    base = rng.integers(0, 256, (bh, bw), dtype=np.uint8)
    img = np.repeat(np.repeat(base, 8, axis=0), 8, axis=1)[:h, :w].astype(np.int16)
    img = img + rng.integers(-8, 9, (h, w), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def orb_batch(n: int, first_seed: int = 1000, w: int = 640, h: int = 480) -> np.ndarray:
    return np.stack([orb_frame(first_seed + i, w, h) for i in range(n)])


def orb_adversarial(kind: str, w: int = 640, h: int = 480, seed: int = 7) -> np.ndarray:
    """Parity-only frames: 'constant' (0 kps), 'noise' (quota saturation, massive ties),
    'lowcontrast' (threshold-7 fallback everywhere), 'gradient' (no corners but non-constant)."""
    rng = np.random.default_rng(seed)
    if kind == "constant":
        return np.full((h, w), 117, np.uint8)
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "lowcontrast":
        base = rng.integers(0, 2, ((h + 7) // 8, (w + 7) // 8), dtype=np.uint8)
        img = np.repeat(np.repeat(base, 8, axis=0), 8, axis=1)[:h, :w].astype(np.int16) * 12 + 100
        img += rng.integers(-1, 2, (h, w), dtype=np.int16)
        return np.clip(img, 0, 255).astype(np.uint8)
    if kind == "gradient":
        x = np.arange(w, dtype=np.int32)[None, :] * 255 // max(w - 1, 1)
        return np.broadcast_to(x, (h, w)).astype(np.uint8).copy()
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------
# Local-BA windows
# --------------------------------------------------------------------------------------------
from se2lam_b200.problem import BAProblem  # noqa: E402  (the product package owns the argument bundle)


def default_Tbc():
    """Camera looking along body +x, 0.3 m up: cam z->body x, cam x->body -y, cam y->body -z."""
    Rbc = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    tbc = np.array([0.1, 0.0, 0.3])
    return Rbc, tbc


def _rotz(th):
    c, s = math.cos(th), math.sin(th)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def edge_information(pose, lw, Rcb, tcb, fx, sigma2, xrot_info=1e6, z_info=1.0):
    """2x2 information of one EdgeSE2XYZ exactly as Map.cpp:1024-1049 builds it (float inputs
    widened to double there; here everything is double, evaluated at the *initial* estimate)."""
    x, y, th = pose
    Rbw = _rotz(-th)
    Rcw = Rcb @ Rbw
    p = np.array([x, y, 0.0])
    lc = Rcw @ (lw - p) + tcb
    zc_inv = 1.0 / lc[2]
    zc_inv2 = zc_inv * zc_inv
    J_pi = np.array([[fx * zc_inv, 0.0, -fx * lc[0] * zc_inv2], [0.0, fx * zc_inv, -fx * lc[1] * zc_inv2]])
    J_pi_Rcw = J_pi @ Rcw
    d = lw - p
    skew = np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
    J_rotxy = (J_pi_Rcw @ skew)[:, 0:2]
    J_z = -J_pi_Rcw[:, 2:3]
    # Map.cpp:1040-1041 stores the variances in float
    s_rot = float(np.float32(1.0 / xrot_info))
    s_z = float(np.float32(1.0 / z_info))
    Sigma_all = s_rot * J_rotxy @ J_rotxy.T + s_z * J_z @ J_z.T + np.eye(2) * sigma2
    return np.linalg.inv(Sigma_all)


def ba_window(n_kf: int, n_lm: int, seed: int = 42, obs_per_lm: int = 6, layout: str = "circle",
              outlier_frac: float = 0.05, noise: bool = True) -> BAProblem:
    """Synthetic local-BA window of SURVEY.md section 8(d).

    circle layout: poses on a 5 m circle, heading tangent, 0.25 m spacing; 'zigzag' layout: long
    zig-zag sweep for the large scale test.  Each landmark is placed inside the frustum of a
    'home' keyframe and observed by up to obs_per_lm consecutive keyframes that actually see it.
    """
    rng = np.random.default_rng(seed)
    fx, cx, cy, W, H = 520.0, 320.0, 240.0, 640, 480
    Rbc, tbc = default_Tbc()
    Rcb = Rbc.T
    tcb = -Rcb @ tbc

    gt = np.zeros((n_kf, 3))
    if layout == "circle":
        # 5 m circle, 0.25 m spacing, heading tangent; headings centred on 0 so that they stay well
        # inside (-pi, pi) (PreEdgeSE2 does not normalise its angle error, EdgeSE2XYZ.h:80)
        R = 5.0
        dth = 0.25 / R
        for i in range(n_kf):
            hd = (i - n_kf / 2) * dth
            a = hd - math.pi / 2
            gt[i] = (R * math.cos(a), R * math.sin(a), hd)
    elif layout == "zigzag":
        # large-scale test: 0.25 m steps, heading +-0.7 rad alternating every 240 KFs, drifting along +x
        x = y = 0.0
        for i in range(n_kf):
            hd = 0.7 if (i // 240) % 2 == 0 else -0.7
            gt[i] = (x, y, hd)
            x += 0.25 * math.cos(hd); y += 0.25 * math.sin(hd)
    else:
        raise ValueError(layout)

    def project(pose, lw):
        Rcw = Rcb @ _rotz(-pose[2])
        lc = Rcw @ (lw - np.array([pose[0], pose[1], 0.0])) + tcb
        if lc[2] <= 0.5:
            return None, lc
        u = fx * lc[0] / lc[2] + cx
        v = fx * lc[1] / lc[2] + cy
        return (u, v), lc

    # initial (drifted) poses: truth (+) accumulated odometry noise
    init = gt.copy()
    if noise:
        drift = np.cumsum(rng.normal(0.0, [0.02, 0.02, 0.01], (n_kf, 3)) * 0.3, axis=0)
        drift[0] = 0.0
        init = gt + drift

    lw_gt = np.zeros((n_lm, 3))
    depths = np.zeros(n_lm)
    e_pose, e_pt, e_uv, e_oct = [], [], [], []
    half = obs_per_lm // 2
    for j in range(n_lm):
        home = int(rng.integers(0, n_kf))
        depth = rng.uniform(2.0, 10.0)
        u = rng.uniform(40, W - 40)
        v = rng.uniform(40, H - 40)
        lc = np.array([(u - cx) / fx * depth, (v - cy) / fx * depth, depth])
        Rcw = Rcb @ _rotz(-gt[home, 2])
        lw = Rcw.T @ (lc - tcb) + np.array([gt[home, 0], gt[home, 1], 0.0])
        lw_gt[j] = lw
        depths[j] = depth
        octave = int(min(7, rng.geometric(1 - 1 / 1.44) - 1))
        sigma = 1.2 ** octave
        start = max(0, min(home - half, n_kf - obs_per_lm))
        for k in range(start, min(start + obs_per_lm, n_kf)):
            uvk, lc = project(gt[k], lw)
            if uvk is None or not (0 <= uvk[0] < W and 0 <= uvk[1] < H):
                continue
            m = np.array(uvk)
            if noise:
                m = m + rng.normal(0.0, sigma, 2)
                if rng.random() < outlier_frac:
                    m = m + rng.uniform(10, 40, 2) * rng.choice([-1.0, 1.0], 2)
            e_pose.append(k); e_pt.append(j); e_uv.append(m); e_oct.append(octave)

    pts_init = lw_gt.copy()
    if noise:
        pts_init += rng.normal(0.0, 1.0, (n_lm, 3)) * (0.03 * depths)[:, None]

    # the reference stores poses / points / uv as float32 and widens them (converter.cpp)
    init = init.astype(np.float32).astype(np.float64)
    pts_init = pts_init.astype(np.float32).astype(np.float64)
    uv = np.asarray(e_uv, np.float64).reshape(-1, 2).astype(np.float32).astype(np.float64)
    e_pose = np.asarray(e_pose, np.int32)
    e_pt = np.asarray(e_pt, np.int32)

    info = np.zeros((len(e_pose), 3))
    for e in range(len(e_pose)):
        sigma2 = float(np.float32(np.float32(1.2) ** e_oct[e]) ** 2)
        Om = edge_information(init[e_pose[e]], pts_init[e_pt[e]], Rcb, tcb, fx, sigma2)
        info[e] = (Om[0, 0], 0.5 * (Om[0, 1] + Om[1, 0]), Om[1, 1])

    # PreEdgeSE2 between consecutive KFs: noisy relative SE(2), cov = diag(0.01^2,0.01^2,0.005^2)
    oi, oj, om, oinf = [], [], [], []
    for i in range(n_kf - 1):
        ci, si = math.cos(gt[i, 2]), math.sin(gt[i, 2])
        d = gt[i + 1, :2] - gt[i, :2]
        rel = np.array([ci * d[0] + si * d[1], -si * d[0] + ci * d[1], gt[i + 1, 2] - gt[i, 2]])
        if noise:
            rel = rel + rng.normal(0.0, [0.01, 0.01, 0.005])
        cov = np.diag([0.01 ** 2, 0.01 ** 2, 0.005 ** 2])
        Om = np.linalg.inv(cov)
        oi.append(i); oj.append(i + 1); om.append(rel)
        oinf.append([Om[0, 0], Om[0, 1], Om[0, 2], Om[1, 1], Om[1, 2], Om[2, 2]])

    fixed = np.zeros(n_kf, np.uint8)
    fixed[0] = 1  # mirrors Map.cpp:927 (min-id KF fixed when there are no reference KFs)
    return BAProblem(
        poses=init, fixed=fixed, points=pts_init, edge_pose=e_pose, edge_point=e_pt, uv=uv, info=info,
        odo_i=np.asarray(oi, np.int32), odo_j=np.asarray(oj, np.int32),
        odo_meas=np.asarray(om, np.float64).reshape(-1, 3), odo_info=np.asarray(oinf, np.float64).reshape(-1, 6),
        fx=fx, cx=cx, cy=cy, Tcb=np.concatenate([Rcb.reshape(-1), tcb]), huber_delta=math.sqrt(5.991),
        gt_poses=gt, gt_points=lw_gt)


BA_CONFIGS = {
    "C1": dict(n_kf=2, n_lm=200, obs_per_lm=2),
    "C3": dict(n_kf=20, n_lm=2000),
    "C4": dict(n_kf=50, n_lm=5000),
    "C5": dict(n_kf=2000, n_lm=50000, layout="zigzag"),
}


def ba_config(name: str, seed: int = 42) -> BAProblem:
    return ba_window(seed=seed, **BA_CONFIGS[name])
