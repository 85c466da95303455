"""Multi-GPU decomposition of the local BA: landmark j belongs to rank j % world.

`shard_problem` is the host-side statement of the rule `se2gpu_ba_set_problem` applies internally when
`se2gpu_ba_set_shard(rank, world)` was called: a rank keeps ALL vertices but only the EdgeSE2XYZ edges of its own
landmarks; PreEdgeSE2 odometry edges (and the lambda damping of the pose block) live on rank 0. Summing the ranks'
reduced systems [S | b] reproduces the single-GPU system, which is what the per-trial all-reduce relies on.
"""
from __future__ import annotations

import copy

import numpy as np


def landmark_owner(j, world: int):
    return j % world


def shard_problem(prob, rank: int, world: int):
    if world == 1:
        return prob
    keep = landmark_owner(prob.edge_point, world) == rank
    q = copy.copy(prob)
    q.edge_pose, q.edge_point, q.uv, q.info = prob.edge_pose[keep], prob.edge_point[keep], prob.uv[keep], prob.info[keep]
    if rank != 0:
        q.odo_i, q.odo_j = prob.odo_i[:0], prob.odo_j[:0]
        q.odo_meas, q.odo_info = prob.odo_meas[:0], prob.odo_info[:0]
    return q


def reduced_system(lin: dict, prob, lam: float, damp_poses: bool):
    """[S | b_s] from a linearisation dict (Hpp, bp, Hll, bl, Hpl as the oracle / se2gpu_ba_debug_system return them)."""
    free = -np.ones(prob.P, int)
    free[np.flatnonzero(prob.fixed == 0)] = np.arange(int((prob.fixed == 0).sum()))
    n = lin["Hpp"].shape[0]
    S = lin["Hpp"].copy()
    b = lin["bp"].copy()
    if damp_poses:
        S[np.arange(n), np.arange(n)] += lam
    active = np.zeros(prob.L, bool)
    active[prob.edge_point] = True
    Dinv = np.zeros((prob.L, 3, 3))
    for j in np.flatnonzero(active):
        Dinv[j] = np.linalg.inv(lin["Hll"][j] + lam * np.eye(3))
    by_lm = {}
    for e in range(prob.E):
        by_lm.setdefault(int(prob.edge_point[e]), []).append(e)
    for j, edges in by_lm.items():
        db = Dinv[j] @ lin["bl"][j]
        for e1 in edges:
            a = free[prob.edge_pose[e1]]
            if a < 0:
                continue
            BD = lin["Hpl"][e1] @ Dinv[j]
            b[3 * a:3 * a + 3] -= lin["Hpl"][e1] @ db
            for e2 in edges:
                c = free[prob.edge_pose[e2]]
                if c < 0:
                    continue
                S[3 * a:3 * a + 3, 3 * c:3 * c + 3] -= BD @ lin["Hpl"][e2].T
    return S, b
