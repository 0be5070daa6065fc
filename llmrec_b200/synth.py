"""Synthetic datasets in the reference's on-disk format.

No real dataset is available offline, so every run (tests, bench, oracle) uses data
written by this module.  File formats follow what the reference reads:

* ``train.json / val.json / test.json``  ``{"uid": [item, ...]}``   (utility/load_data.py:15-28)
* ``train_mat``                          pickled scipy sparse [nu x ni]   (main.py:59)
* ``image_feat.npy``, ``text_feat.npy``  [ni x 512], [ni x 768]           (main.py:54-55)
* ``augmented_user_init_embedding``      pickle, indexable[nu] -> vec     (main.py:61-65)
* ``augmented_atttribute_embedding_dict``pickle {key: indexable[ni]->vec} (main.py:69-79)
* ``augmented_sample_dict``              pickle {uid: {0: pos, 1: neg}}   (main.py:216-220)

Directory names must be ``netflix_valid_item`` / ``preprocessed_raw_MovieLens`` for the
unmodified reference (main.py:69-72); ours also accepts ``netflix`` / ``movielens``.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import scipy.sparse as sp

NETFLIX_KEYS = ("year", "title", "director", "country", "language")
MOVIELENS_KEYS = ("title", "genre", "director", "country", "language")

DATASET_DIR = {"netflix": "netflix_valid_item", "movielens": "preprocessed_raw_MovieLens"}

SHAPES = {
    # name: (n_users, n_items, interactions) -- image/datasets.png of the reference
    "netflix": (13187, 17366, 68933),
    "movielens": (12495, 10322, 57960),
}


def attribute_keys(dataset: str):
    d = dataset.lower()
    if "movielens" in d:
        return MOVIELENS_KEYS
    return NETFLIX_KEYS


def _sample_interactions(rng, nu, ni, n_inter, zipf_a=0.8):
    """>=3 distinct items per user; item popularity ~ 1/(rank+8)^a."""
    extra = max(n_inter - 3 * nu, 0)
    deg = 3 + rng.multinomial(extra, np.full(nu, 1.0 / nu))
    deg = np.minimum(deg, ni)
    w = 1.0 / np.power(np.arange(ni, dtype=np.float64) + 8.0, zipf_a)
    perm = rng.permutation(ni)
    p = np.empty(ni)
    p[perm] = w / w.sum()
    cdf = np.cumsum(p)
    rows = []
    for u in range(nu):
        need = int(deg[u])
        got = np.unique(np.searchsorted(cdf, rng.random(need * 2 + 8)).clip(0, ni - 1))
        while got.size < need:
            more = np.searchsorted(cdf, rng.random(need * 4 + 8)).clip(0, ni - 1)
            got = np.unique(np.concatenate([got, more]))
        rows.append(rng.permutation(got)[:need])
    return rows


def make_dataset(root, dataset="netflix", n_users=None, n_items=None, n_inter=None,
                 dims=(512, 768, 1536), seed=0, feat_dtype=np.float32):
    """Write one dataset under ``root/<reference dir name>/``; returns that path."""
    base = SHAPES.get(dataset, SHAPES["netflix"])
    nu = int(n_users or base[0])
    ni = int(n_items or base[1])
    ne = int(n_inter or base[2])
    d_img, d_txt, d_llm = dims
    rng = np.random.default_rng(seed)
    dname = DATASET_DIR.get(dataset, dataset)
    path = os.path.join(root, dname)
    os.makedirs(path, exist_ok=True)

    rows = _sample_interactions(rng, nu, ni, ne)
    train, val, test = {}, {}, {}
    r_idx, c_idx = [], []
    for u, items in enumerate(rows):
        items = [int(x) for x in items]
        test[str(u)] = items[:1]
        val[str(u)] = items[1:2]
        tr = items[2:]
        train[str(u)] = tr
        r_idx.extend([u] * len(tr))
        c_idx.extend(tr)
    for name, obj in (("train", train), ("val", val), ("test", test)):
        with open(os.path.join(path, name + ".json"), "w") as f:
            json.dump(obj, f)
    mat = sp.csr_matrix((np.ones(len(r_idx), dtype=np.float32), (r_idx, c_idx)), shape=(nu, ni))
    with open(os.path.join(path, "train_mat"), "wb") as f:
        pickle.dump(mat, f)

    np.save(os.path.join(path, "image_feat.npy"), rng.standard_normal((ni, d_img)).astype(feat_dtype))
    np.save(os.path.join(path, "text_feat.npy"), rng.standard_normal((ni, d_txt)).astype(feat_dtype))
    with open(os.path.join(path, "augmented_user_init_embedding"), "wb") as f:
        pickle.dump(rng.standard_normal((nu, d_llm)).astype(feat_dtype), f)
    att = {k: rng.standard_normal((ni, d_llm)).astype(feat_dtype) for k in attribute_keys(dname)}
    with open(os.path.join(path, "augmented_atttribute_embedding_dict"), "wb") as f:
        pickle.dump(att, f)
    aug = {u: {0: int(a), 1: int(b)} for u, (a, b) in enumerate(rng.integers(0, ni, size=(nu, 2)))}
    with open(os.path.join(path, "augmented_sample_dict"), "wb") as f:
        pickle.dump(aug, f)
    return path
