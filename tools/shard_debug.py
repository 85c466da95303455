"""Diagnostic: per-phase busy cycles of the sharded persistent kernel (two contexts on ONE GPU, 70 CTAs each) next to a single
context limited to 70 CTAs. SE2GPU_BA_DEBUG=1 python tools/shard_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SE2GPU_BA_PK_GRID"] = "70"
os.environ["SE2GPU_BA_DEBUG"] = "1"
from tools import synth
from se2lam_b200.ba import LocalBA
from tests.local_shards import run_local_shards
prob = synth.ba_config("C4")
print("=== single context, 70 CTAs", flush=True)
ba = LocalBA.from_problem(prob)
ba.optimize(10); ba.reset(); ba.optimize(10)
del ba
print("=== two sharded contexts, 70 CTAs each", flush=True)
sys.stderr.write("=== sharded\n"); sys.stderr.flush()
run_local_shards(prob, 2, 10, setup=LocalBA.attach_local, mode=2)
