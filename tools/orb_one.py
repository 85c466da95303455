"""One 640x480 frame through the extractor (whatever kernel variants the SE2GPU_ORB_* variables select) against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from se2lam_b200.orb import ORBextractor
from se2lam_b200 import _capi
img = synth.orb_frame(1000)
ext = ORBextractor(1000, 1.2, 8, fastTh=20, max_width=640, max_height=480, max_batch=1, device=0)
try:
    kps, desc = ext(img)
except Exception as e:
    print("extract failed:", e, "| last error:", _capi.last_error()); sys.exit(1)
print("extracted", len(kps))
if "--oracle" in sys.argv:
    from oracle import pyoracle
    ko, do_ = pyoracle.OrbOracle(1000, 1.2, 8, 20).extract(img)
    print("bit-exact vs oracle:", kps.tobytes() == ko.tobytes() and desc.tobytes() == do_.tobytes(), len(ko))
