"""Streaming host path (se2gpu_orb_submit / _wait, two 64-frame batches in flight, page-locked buffers) for several chunk counts
(SE2GPU_ORB_SUBMIT_CHUNKS, read once per process -> one child per setting). Prints ms per batch and a checksum of the results of
every setting (they must agree)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import hashlib, json, os, sys, time
sys.path.insert(0, %(root)r)
import numpy as np, torch
from tools import synth
from se2lam_b200 import _capi
from se2lam_b200.orb import ORBextractor
B, NF, W, H, NROT = 64, 1000, 640, 480, 8
lib = _capi.lib()
base = synth.orb_batch(B, first_seed=1000)
hosts = [torch.from_numpy(np.ascontiguousarray(np.roll(base, r * 7, axis=2) if r else base)).pin_memory() for r in range(NROT)]
ext = ORBextractor(NF, 1.2, 8, fastTh=20, max_width=W, max_height=H, max_batch=B, device=0)
def pinned_out():
    return (torch.empty(B * NF * 28, dtype=torch.uint8).pin_memory().numpy(), torch.empty(B * NF * 32, dtype=torch.uint8).pin_memory().numpy(),
            torch.zeros(B, dtype=torch.int32).pin_memory().numpy())
outs = [pinned_out(), pinned_out()]
def submit(k):
    hb = hosts[k %% NROT].numpy(); o = outs[k & 1]
    _capi.check(lib.se2gpu_orb_submit(ext.h, hb.ctypes.data, B, W, H, W, W * H, o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data), "submit")
def wait(k):
    _capi.check(lib.se2gpu_orb_wait(ext.h), "wait")
sha = hashlib.sha256()
for k in range(NROT):            # warm-up + checksum, one batch at a time
    submit(k); wait(k)
    o = outs[k & 1]; c = o[2]
    sha.update(c.tobytes())
    kk = o[0].reshape(B, NF, 28); dd = o[1].reshape(B, NF, 32)
    for i in range(B):
        sha.update(kk[i, :c[i]].tobytes()); sha.update(dd[i, :c[i]].tobytes())
for k in range(2): submit(k)
for k in range(2): wait(k)
torch.cuda.synchronize()
R = 40
t0 = time.perf_counter()
for k in range(R):
    submit(k)
    if k >= 1: wait(k - 1)
wait(R - 1)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / R
print(json.dumps({"submit_chunks": os.environ.get("SE2GPU_ORB_SUBMIT_CHUNKS"), "ms_per_batch": round(ms, 4), "mkps": round(B * NF / ms / 1e3, 2), "sha256": sha.hexdigest()[:16]}))
'''
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "orb_e2e_submit.jsonl")
with open(out, "w") as f:
    for c in ("4", "1", "2", "3", "8"):
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=dict(os.environ, SE2GPU_ORB_SUBMIT_CHUNKS=c), capture_output=True, text=True, timeout=150)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"submit_chunks": c, "error": r.stderr[-400:]})
        print(line, flush=True); f.write(line + "\n"); f.flush()
