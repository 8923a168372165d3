"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/playaid_b200.h
declares. No compute calls: this runs without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from playaid_core_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from playaid_core_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "playaid_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", hdr)))
    assert declared and set(declared) == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_strings(lib):
    assert lib.pa_abi_version() == 2
    assert lib.pa_status_string(0) == b"ok" and b"workspace" in lib.pa_status_string(-5)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must raise, not compute."""
    import torch

    from playaid_core_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PlayaidLibraryError):
        _lib.Context.get(0)
    from playaid_core_b200.fighter import YoloCrop
    import numpy as np

    with pytest.raises(Exception):
        YoloCrop(0.5, 0.5, 0.1, 0.1).square_crop(np.zeros((64, 64, 3), np.uint8), 128)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "playaid_core_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "from workloads" not in src, f
