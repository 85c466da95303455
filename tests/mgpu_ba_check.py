"""Launched by tests/test_multigpu.py under torch.distributed.run (one rank per GPU): landmark-sharded BA with an
NCCL all-reduce of the reduced system must follow the single-device oracle trajectory."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from tools import synth  # noqa: E402
from se2lam_b200.ba import LocalBA  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cache = {}

    def allreduce(ptr, count, op, strm):
        if (ptr, count) not in cache:
            class A:
                pass
            a = A()
            a.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
            cache[(ptr, count)] = torch.as_tensor(a, device=dev)
        dist.all_reduce(cache[(ptr, count)], op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)

    def all_gather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    ok = True
    for cfg, fused in (("C3", False), ("C4", False), ("C3", True), ("C4", True)):
        prob = synth.ba_config(cfg)
        ba = LocalBA.from_problem(prob, device=local, rank=rank, world=world, allreduce=allreduce,
                                  stream=torch.cuda.current_stream().cuda_stream)
        if fused:   # reduced system summed inside the solve kernel over NVLink peer mappings instead of an NCCL all-reduce
            ba.enable_peer_exchange(all_gather_bytes)
        import time
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        n, st, tp, tl = ba.optimize(10, trace=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        # every rank holds all poses and its own landmarks; gather the landmark estimates
        _, pts = ba.get()
        t = torch.from_numpy(pts).to(dev)
        own = torch.from_numpy((np.arange(prob.L) % world == rank)).to(dev)
        t = torch.where(own[:, None], t, torch.zeros_like(t))
        # landmarks without edges are owned by nobody's kernels but keep their loaded value on every rank
        dist.all_reduce(t)
        if rank == 0:
            o = pyoracle.BAOracle(prob)
            n_o, st_o, tp_o, tl_o = o.optimize(10, trace=True)
            po, lo = o.get()
            active = np.zeros(prob.L, bool); active[prob.edge_point] = True
            good = (n == n_o and np.array_equal(st["trials"], st_o["trials"]) and np.allclose(st["lambda"], st_o["lambda"], rtol=1e-6)
                    and np.abs(tp[-1] - po).max() < 1e-8 and np.abs(t.cpu().numpy()[active] - lo[active]).max() < 1e-7)
            print(f"{cfg} ({'fused peer exchange' if fused else 'NCCL all-reduce'}, {dt * 1e3:.2f} ms): world={world} iters {n}/{n_o} pose err {np.abs(tp[-1] - po).max():.2e} "
                  f"landmark err {np.abs(t.cpu().numpy()[active] - lo[active]).max():.2e} -> {'OK' if good else 'MISMATCH'}")
            ok &= bool(good)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
