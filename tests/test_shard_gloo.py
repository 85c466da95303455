"""N>1 host logic on CPU: world_size-2 gloo processes each linearise their landmark shard (with the oracle as the
compute stand-in), all-reduce [S | b | chi2] and must reproduce the unsharded reduced system."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tools import shard, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    prob = synth.ba_window(n_kf=8, n_lm=150, seed=9)
    lam = 0.37
    mine = shard.shard_problem(prob, rank, world)
    # every landmark's edges live on exactly one rank
    assert set(np.unique(mine.edge_point) % world) <= {rank}
    o = pyoracle.BAOracle(mine)
    lin = o.linearize()
    S, b = shard.reduced_system(lin, mine, lam, damp_poses=(rank == 0))
    buf = torch.from_numpy(np.concatenate([S.reshape(-1), b, [lin["chi2"]]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)          # the one collective of a lambda-trial
    if rank == 0:
        full = pyoracle.BAOracle(prob)
        lin_f = full.linearize()
        ss = full.schur_solve(lam)
        n = S.shape[0]
        got = buf.numpy()
        S_sum, b_sum, chi = got[:n * n].reshape(n, n), got[n * n:n * n + n], got[-1]
        tril = np.tril(np.ones((n, n), bool))
        out["S"] = float(np.abs(S_sum[tril] - ss["S"][tril]).max() / np.abs(ss["S"]).max())
        out["b"] = float(np.abs(b_sum - ss["bs"]).max() / np.abs(ss["bs"]).max())
        out["chi"] = float(abs(chi - lin_f["chi2"]) / lin_f["chi2"])
        out["edges"] = int(mine.E)
    dist.destroy_process_group()


def test_landmark_sharding_sums_to_the_full_reduced_system():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out["S"] < 1e-12 and out["b"] < 1e-12 and out["chi"] < 1e-12
    assert 0 < out["edges"]


def test_shard_problem_partitions_edges():
    prob = synth.ba_config("C3")
    parts = [shard.shard_problem(prob, r, 4) for r in range(4)]
    assert sum(p.E for p in parts) == prob.E
    assert sum(p.O for p in parts) == prob.O and parts[0].O == prob.O
    for r, p in enumerate(parts):
        assert np.all(p.edge_point % 4 == r)
        assert p.P == prob.P and p.L == prob.L
