"""CPU-side checks of the C-ABI library: it loads, and exports every symbol include/llmrec_b200.h declares."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, "include", "llmrec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llmrec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from llmrec_b200 import _native
    from llmrec_b200.build import build
    build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/llmrec_b200.h but not exported"
    assert set(names) == set(_native.SIGNATURES), set(names) ^ set(_native.SIGNATURES)
    lib.llmrec_abi_version.restype = ctypes.c_int
    assert lib.llmrec_abi_version() == 2


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from llmrec_b200 import _native
    lib = _native.lib()
    assert lib.llmrec_device_ok() == 0
    rc = lib.llmrec_fill_f32(None, 16, 0.0, None)
    assert rc != 0 and b"no sm_100" in lib.llmrec_last_error()
    from llmrec_b200.main import Trainer
    with pytest.raises(RuntimeError):
        Trainer()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "llmrec_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "/root/reference" not in txt, f
