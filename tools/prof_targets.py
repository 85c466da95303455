"""Small, fixed workloads for `ncu -k regex:<kernel>` captures (profiles/): python tools/prof_targets.py {ba|c5|match|orb}."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools import synth  # noqa: E402

what = sys.argv[1]
if what == "ba":
    from se2lam_b200.ba import LocalBA
    ba = LocalBA.from_problem(synth.ba_config("C4"))
    for _ in range(3):
        ba.reset(); ba.optimize(10)
elif what == "c5":
    from se2lam_b200.ba import LocalBA
    prob = synth.ba_window(n_kf=int(sys.argv[2]) if len(sys.argv) > 2 else 2000, n_lm=int(sys.argv[3]) if len(sys.argv) > 3 else 50000, seed=42, layout="zigzag")
    ba = LocalBA.from_problem(prob)
    for _ in range(2):
        ba.reset(); ba.optimize(3)
elif what == "match":
    from se2lam_b200.matcher import FrameView, ORBmatcher
    from se2lam_b200.orb import ORBextractor
    e = ORBextractor(1000, 1.2, 8)
    img = synth.orb_frame(2000)
    k1, d1 = e(img); k2, d2 = e(np.roll(img, (3, -5), axis=(0, 1)))
    mt = ORBmatcher(0.9, max_queries=1000, max_db=1000)
    for _ in range(4):
        prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
        mt.MatchByWindow(FrameView(k1, d1), FrameView(k2, d2), prev, 20)
elif what == "orb":
    from se2lam_b200.orb import ORBextractor
    dev = torch.device("cuda", 0)
    B = 64
    e = ORBextractor(1000, 1.2, 8, max_batch=B)
    imgs = torch.from_numpy(synth.orb_batch(B)).to(dev)
    kps = torch.empty(B * 1000 * 28, dtype=torch.uint8, device=dev); desc = torch.empty(B * 1000 * 32, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    for _ in range(3):
        e.extract_device(imgs, B, 480, 640, kps, desc, cnt)
    torch.cuda.synchronize()
torch.cuda.synchronize()
print("done", what)
