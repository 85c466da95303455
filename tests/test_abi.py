"""CPU-side checks of the drop-in boundary: the library loads and exports every declared symbol."""
import os
import re

import pytest

from se2lam_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return _capi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "se2gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(se2gpu_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert declared_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"libse2gpu.so does not export {name}"


def test_no_device_is_reported_not_emulated(lib):
    """Without a GPU the product path must fail loudly (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.se2gpu_device_count() == 0
    h = lib.se2gpu_ba_create(4, 4, 4, 4, 0)
    assert not h
    assert "CUDA" in _capi.last_error() or "device" in _capi.last_error()
    h = lib.se2gpu_orb_create(1000, 1.2, 8, 20, 640, 480, 1, 0)
    assert not h


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "se2lam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
