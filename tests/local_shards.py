"""Drives the REAL sharded BA path (se2gpu_ba_set_shard + the all-reduce callback) with several handles on ONE CUDA device.

Every "rank" is a se2gpu_ba context living in its own host thread with its own stream; the all-reduce callback is a
rendezvous between the threads that sums (or maxes) the ranks' device buffers in rank order - the same kernels, the same
host loop and the same collective sequence as a multi-GPU run, without needing a second GPU. Test infrastructure only.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from se2lam_b200.ba import LocalBA


def _as_tensor(ptr, count, dev):
    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(a, device=dev)


def run_local_shards(prob, world, iters, device=0, stop_flags=None, setup=None, mode=0):
    """Returns per-rank (n, stats, trace_poses, trace_points, poses, points). `stop_flags[r]` (optional) is the abort word
    rank r polls; `setup(bas)` (optional) runs after all contexts exist (e.g. to attach a fused exchange)."""
    dev = torch.device("cuda", device)
    barrier = threading.Barrier(world)
    slots = [None] * world
    result = [None] * world
    errors = []
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]

    def make_cb(rank):
        def allreduce(ptr, count, op, stream):
            torch.cuda.synchronize(dev)                 # everything this rank enqueued so far is complete
            slots[rank] = _as_tensor(ptr, count, dev)
            barrier.wait()
            parts = list(slots)
            acc = parts[0].clone()
            for t in parts[1:]:                         # rank order: every rank computes the identical sum
                acc = acc + t if op == 0 else torch.maximum(acc, t)
            torch.cuda.synchronize(dev)
            barrier.wait()                              # nobody overwrites its buffer before everyone has read it
            slots[rank].copy_(acc)
            torch.cuda.synchronize(dev)
        return allreduce

    bas = [None] * world

    def worker(rank):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(streams[rank]):
                ba = LocalBA.from_problem(prob, device=device, rank=rank, world=world, allreduce=make_cb(rank),
                                          stream=streams[rank].cuda_stream)
                ba.set_mode(mode)
                bas[rank] = ba
                barrier.wait()
                if setup is not None and rank == 0:
                    setup(bas)
                barrier.wait()
                n, st, tp, tl = ba.optimize(iters, trace=True, stop_flag=None if stop_flags is None else stop_flags[rank])
                p, l = ba.get()
                result[rank] = (n, st, tp, tl, p, l)
        except Exception as e:      # noqa: BLE001
            errors.append((rank, e))
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    if errors:
        raise errors[0][1]
    return result


def merge_landmarks(prob, results, world):
    """Landmark j is owned (and only updated) by rank j % world."""
    pts = np.array(results[0][5])
    for r in range(1, world):
        own = np.arange(prob.L) % world == r
        pts[own] = results[r][5][own]
    return pts
