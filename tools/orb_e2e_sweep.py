"""Time the host-buffer ORB entry point (se2gpu_orb_extract) for several pipeline chunk counts (SE2GPU_ORB_CHUNKS).
Each setting runs in a child process because the library reads the variable once."""
import os, subprocess, sys, json

CHILD = r'''
import os, time, json, numpy as np, torch
from se2lam_b200 import _capi
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth
from se2lam_b200.orb import ORBextractor
n, NF, W, H = 64, 1000, 640, 480
ex = ORBextractor(NF, 1.2, 8, max_batch=n)
lib = _capi.lib()
hosts = [torch.from_numpy(synth.orb_batch(n, first_seed=100 + 64 * k)).pin_memory() for k in range(4)]
pinned = os.environ.get("PINNED_OUT", "1") == "1"
mk = (lambda nb: torch.empty(nb, dtype=torch.uint8).pin_memory().numpy()) if pinned else (lambda nb: np.zeros(nb, np.uint8))
kps, desc, counts = mk(n * NF * 28), mk(n * NF * 32), np.zeros(n, np.int32)
def step(k):
    hb = hosts[k % 4].numpy()
    _capi.check(lib.se2gpu_orb_extract(ex.h, hb.ctypes.data, n, W, H, W, W * H, kps.ctypes.data, desc.ctypes.data, counts.ctypes.data), "x")
for k in range(4): step(k)
torch.cuda.synchronize()
t0 = time.perf_counter(); R = 30
for k in range(R): step(k)
dt = (time.perf_counter() - t0) / R
print(json.dumps({"chunks": os.environ.get("SE2GPU_ORB_CHUNKS"), "pinned_out": pinned, "ms": round(dt * 1e3, 4), "kps": int(counts.sum())}))
'''

for lanes, c, f in (("4", "4", "50"), ("4", "3", "50"), ("4", "3", "100"), ("4", "4", "100"), ("4", "5", "50"), ("4", "5", "100"), ("4", "6", "50"), ("4", "6", "100"),
                    ("3", "3", "50"), ("3", "3", "100"), ("2", "2", "50"), ("4", "4", "35"), ("4", "5", "35")):
    env = dict(os.environ, SE2GPU_ORB_CHUNKS=c, SE2GPU_ORB_FIRST=f, SE2GPU_ORB_LANES=lanes, PINNED_OUT="1")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("lanes", lanes, "first", f, r.stdout.strip() or r.stderr[-400:], flush=True)
