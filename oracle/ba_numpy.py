"""Independent numpy restatement of the SE(2)-XYZ local BA (TEST INFRASTRUCTURE ONLY).

Deliberately written differently from oracle/ba_oracle.cpp so the two can pin each other:
* edge errors follow the literal transform chain of EdgeSE2XYZ::computeError
  (reference src/EdgeSE2XYZ.cpp:61-72: Tbw = SE2ToSE3(v1.inverse()); Tcw = Tcb*Tbw; cam_map(Tcw.map(lw)) - uv),
* Jacobians are *numeric* (central differences through the additive oplus of VertexSE2 /
  VertexSBAPointXYZ), and
* the damped normal equations are solved as ONE dense system (no Schur complement).
The LM control flow restates g2o's OptimizationAlgorithmLevenberg::solve [upstream, not in repo].
"""
from __future__ import annotations

import math

import numpy as np


def _se2_inverse(x, y, th):
    c, s = math.cos(th), math.sin(th)
    # R(-th), t' = -R(-th) t
    return (-(c * x + s * y), -(-s * x + c * y), -th)


def _se2_to_se3(x, y, th):
    c, s = math.cos(th), math.sin(th)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return R, np.array([x, y, 0.0])


def xyz_error(prob, pose, lw, e):
    Rcb = prob.Tcb[:9].reshape(3, 3)
    tcb = prob.Tcb[9:]
    Rbw, tbw = _se2_to_se3(*_se2_inverse(*pose))
    Rcw = Rcb @ Rbw
    tcw = Rcb @ tbw + tcb
    lc = Rcw @ lw + tcw
    return np.array([lc[0] / lc[2] * prob.fx + prob.cx, lc[1] / lc[2] * prob.fx + prob.cy]) - prob.uv[e]


def odo_error(prob, pi, pj, o):
    c, s = math.cos(pi[2]), math.sin(pi[2])
    Ri = np.array([[c, -s], [s, c]])
    e = np.zeros(3)
    e[:2] = Ri.T @ (pj[:2] - pi[:2]) - prob.odo_meas[o, :2]
    e[2] = pj[2] - pi[2] - prob.odo_meas[o, 2]
    return e


def _info2(w):
    return np.array([[w[0], w[1]], [w[1], w[2]]])


def _info3(w):
    return np.array([[w[0], w[1], w[2]], [w[1], w[3], w[4]], [w[2], w[4], w[5]]])


def huber(chi2, delta):
    d2 = delta * delta
    if chi2 <= d2:
        return chi2, 1.0
    s = math.sqrt(chi2)
    return 2 * s * delta - d2, delta / s


def total_chi2(prob, poses, points):
    chi = 0.0
    for o in range(prob.O):
        e = odo_error(prob, poses[prob.odo_i[o]], poses[prob.odo_j[o]], o)
        chi += e @ _info3(prob.odo_info[o]) @ e
    for k in range(prob.E):
        e = xyz_error(prob, poses[prob.edge_pose[k]], points[prob.edge_point[k]], k)
        chi += huber(e @ _info2(prob.info[k]) @ e, prob.huber_delta)[0]
    return chi


def _numjac(f, x, eps=1e-6):
    f0 = f(x)
    J = np.zeros((len(f0), len(x)))
    for i in range(len(x)):
        xp = x.copy(); xm = x.copy()
        xp[i] += eps; xm[i] -= eps
        J[:, i] = (f(xp) - f(xm)) / (2 * eps)
    return J


def build_full_system(prob, poses, points):
    """Dense H, b over [free poses ; active landmarks] with numeric Jacobians."""
    hidx = -np.ones(prob.P, int)
    nf = 0
    for i in range(prob.P):
        if not prob.fixed[i]:
            hidx[i] = nf; nf += 1
    active = np.zeros(prob.L, bool)
    active[prob.edge_point] = True
    lidx = -np.ones(prob.L, int)
    nl = 0
    for j in range(prob.L):
        if active[j]:
            lidx[j] = nl; nl += 1
    n = 3 * nf + 3 * nl
    H = np.zeros((n, n)); b = np.zeros(n)

    def add(idx_a, Ja, idx_b, Jb, W, r):
        for (ia, JA) in ((idx_a, Ja), (idx_b, Jb)):
            if ia < 0:
                continue
            b[ia:ia + 3] += JA.T @ r
            for (ib, JB) in ((idx_a, Ja), (idx_b, Jb)):
                if ib < 0:
                    continue
                H[ia:ia + 3, ib:ib + 3] += JA.T @ W @ JB

    for o in range(prob.O):
        i, j = prob.odo_i[o], prob.odo_j[o]
        W = _info3(prob.odo_info[o])
        e = odo_error(prob, poses[i], poses[j], o)
        Ji = _numjac(lambda x: odo_error(prob, x, poses[j], o), poses[i].copy())
        Jj = _numjac(lambda x: odo_error(prob, poses[i], x, o), poses[j].copy())
        add(3 * hidx[i] if hidx[i] >= 0 else -1, Ji, 3 * hidx[j] if hidx[j] >= 0 else -1, Jj, W, -W @ e)
    for k in range(prob.E):
        i, j = prob.edge_pose[k], prob.edge_point[k]
        W0 = _info2(prob.info[k])
        e = xyz_error(prob, poses[i], points[j], k)
        _, rho1 = huber(e @ W0 @ e, prob.huber_delta)
        W = rho1 * W0
        Ji = _numjac(lambda x: xyz_error(prob, x, points[j], k), poses[i].copy())
        Jj = _numjac(lambda x: xyz_error(prob, poses[i], x, k), points[j].copy())
        add(3 * hidx[i] if hidx[i] >= 0 else -1, Ji, 3 * nf + 3 * lidx[j], Jj, W, -rho1 * (W0 @ e))
    return H, b, hidx, lidx, nf, nl


def lm_optimize(prob, iters, poses=None, points=None):
    """g2o-style LM on the dense full system. Returns (poses, points, stats list)."""
    poses = prob.poses.copy() if poses is None else poses.copy()
    points = prob.points.copy() if points is None else points.copy()
    lam, ni = 0.0, 2.0
    stats = []
    for it in range(iters):
        cur = total_chi2(prob, poses, points)
        H, b, hidx, lidx, nf, nl = build_full_system(prob, poses, points)
        if it == 0:
            lam = 1e-5 * np.abs(np.diag(H)).max(); ni = 2.0
        rho = 0.0
        q = 0
        chi_before = cur
        while True:
            bak = (poses.copy(), points.copy())
            try:
                Lc = np.linalg.cholesky(H + lam * np.eye(len(b)))
                dx = np.linalg.solve(Lc.T, np.linalg.solve(Lc, b))
                ok2 = True
            except np.linalg.LinAlgError:
                ok2 = False
                dx = np.zeros_like(b)
            if ok2:
                for i in range(prob.P):
                    if hidx[i] >= 0:
                        d = dx[3 * hidx[i]:3 * hidx[i] + 3]
                        poses[i, :2] += d[:2]
                        th = poses[i, 2] + d[2]
                        poses[i, 2] = (th + math.pi) % (2 * math.pi) - math.pi if not (-math.pi <= th < math.pi) else th
                for j in range(prob.L):
                    if lidx[j] >= 0:
                        points[j] += dx[3 * nf + 3 * lidx[j]:3 * nf + 3 * lidx[j] + 3]
            tmp = total_chi2(prob, poses, points) if ok2 else np.finfo(float).max
            scale = (dx @ (lam * dx + b) if ok2 else 0.0) + 1e-3
            rho = (cur - tmp) / scale
            if rho > 0 and np.isfinite(tmp):
                alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha); ni = 2.0; cur = tmp
                accepted = True
            else:
                lam *= ni; ni *= 2; poses, points = bak
                accepted = False
            q += 1
            if not (rho < 0 and q < 10):
                break
        stats.append(dict(chi2_before=chi_before, chi2_after=cur, lam=lam, rho=rho, trials=q, accepted=accepted))
        if q == 10 or rho == 0:
            break
    return poses, points, stats
