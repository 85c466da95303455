"""Scale test (BASELINE config 5: 2000 KF / 50k landmarks / ~300k edges): GPU LM vs the CPU oracle, timing both."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle
from tools import synth
from se2lam_b200.ba import LocalBA

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
t0 = time.time(); prob = synth.ba_config("C5"); print("synth", round(time.time() - t0, 2), "s", "P", len(prob.poses), "L", len(prob.points), "E", len(prob.edge_pose), flush=True)
t0 = time.time(); g = LocalBA.from_problem(prob); print("set_problem", round(time.time() - t0, 3), "s", flush=True)
t0 = time.time(); n_g, st_g, tp_g, tl_g = g.optimize(iters, trace=True); t_g = time.time() - t0
print("gpu", n_g, "iterations", round(t_g, 4), "s", "trials", st_g["trials"][:n_g], "chi2", st_g["chi2_after"][:n_g], flush=True)
g.reset(); t0 = time.time(); n2 = g.optimize(iters)[0]; print("gpu (warm)", n2, round(time.time() - t0, 4), "s", flush=True)
t0 = time.time(); o = pyoracle.BAOracle(prob); n_o, st_o, tp_o, tl_o = o.optimize(iters, trace=True); t_o = time.time() - t0
print("oracle", n_o, "iterations", round(t_o, 3), "s", "chi2", st_o["chi2_after"][:n_o], flush=True)
assert n_g == n_o
print("trials equal", np.array_equal(st_g["trials"][:n_o], st_o["trials"][:n_o]), "lambda rel", np.abs(st_g["lambda"][:n_o] / st_o["lambda"][:n_o] - 1).max())
prev_p, prev_l = prob.poses, prob.points
for k in range(n_o):
    dp_o, dp_g = tp_o[k] - prev_p, tp_g[k] - prev_p
    dl_o, dl_g = tl_o[k] - prev_l, tl_g[k] - prev_l
    print(k, "pose step rel", np.abs(dp_g - dp_o).max() / max(np.abs(dp_o).max(), 1e-12), "lm step rel", np.abs(dl_g - dl_o).max() / max(np.abs(dl_o).max(), 1e-12))
    prev_p, prev_l = tp_o[k], tl_o[k]
g.reset(); g.profile(True); g.optimize(iters); prof = g.profile_read(); g.profile(False)
print("per-kernel ms per launch:", {k: (round(v[0] / max(v[1], 1), 4), v[1]) for k, v in prof.items() if v[1]})
