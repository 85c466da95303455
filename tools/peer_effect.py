"""Does enabling peer access (opening another rank's CUDA-IPC mappings) slow down the single-GPU persistent BA kernel?
torchrun --nproc-per-node 2 tools/peer_effect.py   (diagnostic for the sharded-kernel phase times)"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth
from se2lam_b200.ba import LocalBA

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
prob = synth.ba_config("C4")


def run(tag):
    ba = LocalBA.from_problem(prob, device=local)
    for _ in range(3):
        ba.reset(); ba.optimize(10)
    ba.profile(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        ba.reset(); ba.optimize(10)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    pr = ba.profile_read()
    if rank == 0:
        print(tag, f"{dt * 1e3:.3f} ms/optimize", {k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items() if v[1]}, flush=True)
    del ba


run("before peer access:")
dist.barrier()
# a sharded handle pair only to open the IPC mappings (enables peer access between the two devices)
dummy = LocalBA.from_problem(synth.ba_config("C1"), device=local, rank=rank, world=world, allreduce=lambda *a: None)
def gather(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out
dummy.enable_peer_exchange(gather)
dist.barrier()
run("after peer access: ")
dist.barrier()
dist.destroy_process_group()
