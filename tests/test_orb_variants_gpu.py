"""Every selectable ORB kernel variant against the CPU oracle, one child process per setting because the
library reads the switches once per process: the round-1 kernels, the 4-pixel TMA kernel, the defaults without programmatic dependent
launch, without the window-form resize, without batched orientation loads. tests/test_orb_gpu.py covers the defaults in depth; this keeps
the alternatives (and the fallback the library takes when the driver has no tensor-map encoder) bit-exact too."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [   # combinations that tools/orb_variants.py has run on a B200 (profiles/r02c / r02e / r02f _orb_variants.jsonl)
    {"SE2GPU_ORB_FAST_TMA": "0", "SE2GPU_ORB_ORIENT_BATCH": "0", "SE2GPU_ORB_RESIZE_W": "0", "SE2GPU_ORB_PDL": "0", "SE2GPU_ORB_BLUR_B_AFTER_FAST": "0"},
    {"SE2GPU_ORB_FAST_TMA": "1", "SE2GPU_ORB_ORIENT_BATCH": "0", "SE2GPU_ORB_RESIZE_W": "0", "SE2GPU_ORB_PDL": "0", "SE2GPU_ORB_BLUR_B_AFTER_FAST": "0"},
    {"SE2GPU_ORB_PDL": "0"},
    {"SE2GPU_ORB_PDL": "1", "SE2GPU_ORB_BLUR_B_AFTER_FAST": "0"},
    {"SE2GPU_ORB_RESIZE_W": "0", "SE2GPU_ORB_PDL": "0"},
    {"SE2GPU_ORB_ORIENT_BATCH": "0"},
]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[11:]}={v}" for k, v in e.items()))
def test_variant_is_bit_exact(env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "orb_one.py"), "--oracle"], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-exact vs oracle: True" in r.stdout, r.stdout + r.stderr
