"""API-level compatibility with the reference (container only for the live comparisons; skipped on the GPU box)."""
import os
import sys

import numpy as np
import pytest

from oracle.ref_shim import REFERENCE_ROOT, reference_available


def test_flags_match_reference_parser():
    """Every flag of utility/parser.py exists here with the same default and type (drop-in CLI)."""
    if not reference_available():
        pytest.skip("reference tree not present")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_parser", os.path.join(REFERENCE_ROOT, "utility", "parser.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["x"]
    try:
        spec.loader.exec_module(mod)
        ref = vars(mod.parse_args())
    finally:
        sys.argv = argv
    from llmrec_b200.utility.parser import parse_args
    mine = vars(parse_args([]))
    assert len(ref) == 42
    for k, v in ref.items():
        assert k in mine, k
        assert mine[k] == v and type(mine[k]) is type(v), (k, mine[k], v)
    # same parsing behaviour on a typical command line, including the type=bool quirk of --mask
    sys.argv = ["x", "--dataset", "netflix", "--embed_size", "128", "--weight_size", "[128,128]", "--lr", "0.01", "--mask", "False"]
    try:
        r = vars(mod.parse_args())
    finally:
        sys.argv = argv
    m = vars(parse_args(["--dataset", "netflix", "--embed_size", "128", "--weight_size", "[128,128]", "--lr", "0.01", "--mask", "False"]))
    for k, v in r.items():
        assert m[k] == v, k
    assert m["mask"] is True          # any non-empty string is True upstream (parser.py:39)


def test_data_matches_reference_loader(tiny_root):
    if not reference_available():
        pytest.skip("reference tree not present")
    from llmrec_b200.utility.load_data import Data
    ddir = os.path.join(tiny_root, "netflix_valid_item")
    d = Data(ddir, 128)
    import json
    tr = json.load(open(os.path.join(ddir, "train.json")))
    assert d.n_users == max(int(k) for k, v in tr.items() if v) + 1            # load_data.py:35,55
    assert d.n_items == np.load(os.path.join(ddir, "text_feat.npy")).shape[0]   # load_data.py:57-58
    assert d.n_train == sum(len(v) for v in tr.values())
    assert d.exist_users == [int(k) for k, v in tr.items() if v]
    R = d.R
    assert R.shape == (d.n_users, d.n_items) and R.nnz == d.n_train and R[0, d.train_items[0][0]] == 1.0
    rp, col = d.csr("train")
    assert col[rp[5]:rp[6]].tolist() == d.train_items[5]                       # JSON order kept for the sampler
    rps, cols = d.csr("train", sorted_rows=True)
    assert cols[rps[5]:rps[6]].tolist() == sorted(d.train_items[5])


def test_dataset_alias_resolution(tmp_path):
    from llmrec_b200.utility.parser import resolve_dataset_dir
    root = str(tmp_path) + "/"
    os.makedirs(root + "netflix_valid_item")
    assert resolve_dataset_dir(root, "netflix").endswith("netflix_valid_item")
    assert resolve_dataset_dir(root, "netflix_valid_item").endswith("netflix_valid_item")
    os.makedirs(root + "netflix")
    assert resolve_dataset_dir(root, "netflix").endswith("/netflix")


def test_out_of_scope_flags_fail_loudly():
    import torch
    from llmrec_b200.Models import MM_Model
    from llmrec_b200.runtime import set_args
    from llmrec_b200.utility.parser import parse_args
    set_args(parse_args(["--mask_rate", "0.1"]))
    m = MM_Model(4, 5, 64, [64], [0.1], np.zeros((5, 8), np.float32), np.zeros((5, 8), np.float32), np.zeros((4, 8), np.float32),
                 {"title": np.zeros((5, 8), np.float32)})
    with pytest.raises(NotImplementedError):
        m.forward(torch.zeros(1), torch.zeros(1))
    set_args(parse_args([]))
