"""Binary CSR interaction store: same Data as the json walk, loaded without parsing."""
import json
import os
import random
import shutil
import time

import numpy as np
import pytest

from llmrec_b200.synth import make_dataset
from llmrec_b200.utility import csr_store
from llmrec_b200.utility.load_data import Data


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("csr"))
    p = make_dataset(root, n_users=257, n_items=190, n_inter=1900, dims=(8, 8, 8), seed=3)
    # shuffle the json key order and leave one user without train items: both must survive the conversion
    for name in ("train", "test", "val"):
        d = json.load(open(os.path.join(p, name + ".json")))
        keys = list(d)
        random.Random(7).shuffle(keys)
        d = {k: d[k] for k in keys}
        if name == "train":
            d["5"] = []
        json.dump(d, open(os.path.join(p, name + ".json"), "w"))
    q = p + "_csr"
    shutil.copytree(p, q)
    csr_store.convert_json(q)
    return Data(p, 64), Data(q, 64)


def test_attributes_equal(pair):
    a, b = pair
    for k in ("n_users", "n_items", "n_train", "n_test", "exist_users"):
        assert getattr(a, k) == getattr(b, k), k
    assert 5 not in b.train_items and 5 not in b.exist_users
    with pytest.raises(KeyError):
        b.train_items[5]
    for name in ("train_items", "test_set", "val_set"):
        x, y = getattr(a, name), getattr(b, name)
        assert list(x.keys()) == list(y.keys()), name           # json key order kept
        assert dict(x) == dict(y.items()), name
        assert len(x) == len(y)
    assert b.train_items.get(10 ** 9, []) == []


def test_csr_and_R_equal(pair):
    a, b = pair
    for w in ("train", "test", "val"):
        for s in (False, True):
            for x, y in zip(a.csr(w, s), b.csr(w, s)):
                assert x.dtype == y.dtype and np.array_equal(x, y)
    assert (a.R != b.R).nnz == 0


@pytest.mark.parametrize("mode", ["python", "native"])
def test_same_batches(pair, mode):
    a, b = pair
    a._sampler = b._sampler = mode
    random.seed(11); np.random.seed(11)
    x = [a.sample() for _ in range(4)]
    random.seed(11); np.random.seed(11)
    y = [b.sample() for _ in range(4)]
    assert x == y


def test_store_rejects_bad_input(tmp_path):
    rp = np.array([0, 2, 1], dtype=np.int64)
    with pytest.raises(ValueError):
        csr_store.write(str(tmp_path), {"train": (rp, np.zeros(1, np.int32))}, 2, 4)
    with pytest.raises(ValueError):
        csr_store.write(str(tmp_path), {"train": (np.array([0, 1]), np.zeros(1, np.int32))}, 2, 4)
    with pytest.raises(ValueError):
        csr_store.rows_from_dict({"9": [1]}, 4)
    csr_store.write(str(tmp_path), {"train": (np.array([0, 1, 1]), np.array([3], np.int32))}, 2, 4)
    meta = json.load(open(os.path.join(str(tmp_path), csr_store.META)))
    meta["format"] = "other"
    json.dump(meta, open(os.path.join(str(tmp_path), csr_store.META), "w"))
    with pytest.raises(ValueError):
        csr_store.read(str(tmp_path))


def test_large_store_loads_without_parsing(tmp_path):
    """2 M users x 20 M interactions: the json walk takes minutes; the store must open in well under a few seconds."""
    nu, ni, per = 2_000_000, 100_000, 10
    rng = np.random.default_rng(0)
    rowptr = np.arange(nu + 1, dtype=np.int64) * per
    col = rng.integers(0, ni, size=nu * per, dtype=np.int32)
    csr_store.write(str(tmp_path), {"train": (rowptr, col)}, nu, ni, validate=False)
    t = time.time()
    d = Data(str(tmp_path), 1024, sampler="native")
    dt = time.time() - t
    assert d.n_users == nu and d.n_items == ni and d.n_train == nu * per and len(d.test_set) == 0
    random.seed(0); np.random.seed(0)
    users, pos, neg = d.sample()
    assert len(users) == 1024 and all(p in d.train_items[u] for u, p in zip(users, pos))
    assert all(n not in d.train_items[u] for u, n in zip(users, neg))
    assert dt < 15.0, dt          # ~0.5 s here; the json walk of the same data takes minutes


def test_store_shards_tile_the_graph(tmp_path):
    """dist.store_shard: the rank slices of a store are disjoint, cover every (deduplicated) edge and use local rows."""
    import torch
    from llmrec_b200.dist import shard_bounds, store_shard
    rng = np.random.default_rng(5)
    nu, ni = 103, 41
    deg = rng.integers(0, 9, size=nu)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, ni, size=int(rowptr[-1]), dtype=np.int32)              # duplicates inside rows on purpose
    csr_store.write(str(tmp_path), {"train": (rowptr, col)}, nu, ni)
    want = np.unique(np.repeat(np.arange(nu), deg) * ni + col)
    for world in (1, 2, 3, 8):
        got = []
        b = shard_bounds(nu, world)
        for r in range(world):
            u, it, lo, hi, n_users, n_items = store_shard(str(tmp_path), r, world, "cpu")
            assert (lo, hi, n_users, n_items) == (b[r], b[r + 1], nu, ni)
            assert u.numel() == 0 or (0 <= int(u.min()) and int(u.max()) < hi - lo)
            got.append(((u + lo) * ni + it).numpy())
        got = np.concatenate(got)
        assert got.shape == want.shape and np.array_equal(np.sort(got), want)
