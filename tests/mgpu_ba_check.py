"""Launched by tests/test_multigpu.py under torch.distributed.run (one rank per process): landmark-sharded BA must follow
the single-device oracle trajectory, both through the all-reduce callback (multi-launch path) and through the sharded
persistent kernel that exchanges over CUDA-IPC peer mappings.

With one GPU per rank the process group is NCCL. On a box with fewer GPUs than ranks every rank uses GPU 0: the process group
is gloo (the callback sums through host memory), the IPC mappings are those of another process on the SAME device, and each
rank's cooperative grid is limited to a share of the SMs (SE2GPU_BA_PK_GRID) - the same code path as the multi-GPU run."""
import os
import sys
import time

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
    shared_gpu = torch.cuda.device_count() < world
    if shared_gpu:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if shared_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    cache = {}

    def allreduce(ptr, count, op, strm):
        if (ptr, count) not in cache:
            class A:
                pass
            a = A()
            a.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
            cache[(ptr, count)] = torch.as_tensor(a, device=dev)
        t = cache[(ptr, count)]
        rop = dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX
        if shared_gpu:                      # gloo: through host memory
            torch.cuda.synchronize()
            h = t.cpu()
            dist.all_reduce(h, op=rop)
            t.copy_(h)
            torch.cuda.synchronize()
        else:
            dist.all_reduce(t, op=rop)

    def all_gather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    ok = True
    for cfg, fused in (("C3", False), ("C4", False), ("C3", True), ("C4", True)):
        prob = synth.ba_config(cfg)
        ba = LocalBA.from_problem(prob, device=local, rank=rank, world=world, allreduce=allreduce,
                                  stream=torch.cuda.current_stream().cuda_stream)
        if fused:   # sharded persistent kernel: [S | b] and the scalars exchanged over peer mappings inside the kernel
            ba.enable_peer_exchange(all_gather_bytes)
            ba.set_mode(2)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        n, st, tp, tl = ba.optimize(10, trace=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        # every rank holds all poses and its own landmarks; gather the landmark estimates
        _, pts = ba.get()
        t = torch.from_numpy(pts)
        own = torch.from_numpy((np.arange(prob.L) % world == rank))
        t = torch.where(own[:, None], t, torch.zeros_like(t))
        if not shared_gpu:
            t = t.to(dev)
        dist.all_reduce(t)
        if rank == 0:
            o = pyoracle.BAOracle(prob)
            n_o, st_o, tp_o, tl_o = o.optimize(10, trace=True)
            po, lo = o.get()
            active = np.zeros(prob.L, bool); active[prob.edge_point] = True
            good = (n == n_o and np.array_equal(st["trials"], st_o["trials"]) and np.allclose(st["lambda"], st_o["lambda"], rtol=1e-6)
                    and np.abs(tp[-1] - po).max() < 1e-8 and np.abs(t.cpu().numpy()[active] - lo[active]).max() < 1e-7)
            print(f"{cfg} ({'persistent kernel, peer exchange' if fused else 'all-reduce callback'}, {dt * 1e3:.2f} ms, "
                  f"{'ranks share GPU 0 (gloo)' if shared_gpu else 'one GPU per rank (NCCL)'}): world={world} iters {n}/{n_o} pose err {np.abs(tp[-1] - po).max():.2e} "
                  f"landmark err {np.abs(t.cpu().numpy()[active] - lo[active]).max():.2e} -> {'OK' if good else 'MISMATCH'}")
            ok &= bool(good)
        del ba
    flag = torch.tensor([1 if ok else 0])
    if not shared_gpu:
        flag = flag.to(dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
