"""Import the UNMODIFIED reference from /root/reference on CPU.  Container-only helper.

Used by tests/golden/make_golden.py (golden-vector generation) and
tests/test_oracle_vs_reference.py (skipped when /root/reference is absent, e.g. on the GPU
box).  Nothing is copied: the reference modules are imported in place under three shims
(SURVEY.md "Facts established by probing"):

  1. torch.Tensor.cuda / nn.Module.cuda -> identity      (no GPU in the build container)
  2. np.asfarray re-added                                 (removed in NumPy 2; metrics.py:50,75)
  3. sys.argv set before import                           (4 modules call parse_args() at import)

The dataset directory must be named netflix_valid_item / preprocessed_raw_MovieLens
(main.py:69-72) and --debug avoids the ./logs/ requirement (utility/logging.py:12-14).
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main.py"))


def import_reference(argv):
    """Returns the reference's `main` module (Trainer, data_generator, test_torch, ...)."""
    import numpy as np
    import torch

    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    sys.argv = ["main.py"] + list(argv)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("main", "Models", "utility", "utility.parser", "utility.batch_test", "utility.load_data",
                 "utility.metrics", "utility.logging", "utility.norm"):
        sys.modules.pop(name, None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)
    import main as ref_main  # noqa: E402  (the reference's main.py)
    return ref_main
