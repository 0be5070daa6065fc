"""Binary CSR form of the interaction files  --  SURVEY.md 8f-2 (scalable loader).

The reference reads `train.json / test.json / val.json` (`{"uid": [item, ...]}`, utility/load_data.py:15-28) and walks
them one interaction at a time in Python (load_data.py:29-52, 63-92): minutes at 200 M edges.  This module stores the
same three relations as memory-mappable arrays next to (or instead of) the json files:

    interactions.csr.json                     {"format": "llmrec-csr-v1", "n_users": .., "n_items": .., "splits": [...]}
    <split>.rowptr.npy   int64 [n_users + 1]  row u = col[rowptr[u] : rowptr[u+1]], items in the json list's order
    <split>.col.npy      int32 [nnz]
    <split>.order.npy    int32 [rows]         (optional) non-empty users in the json file's key order; absent = ascending

`Data` (utility/load_data.py here) picks this form up automatically when `interactions.csr.json` exists, and exposes the
same attributes: `train_items / test_set / val_set` become read-only mappings over the arrays (`CsrRows`), so
`train_items[u]` is still a list of ints and `list(test_set.keys())` still enumerates the users with a non-empty row in
file order.  Loading is O(1) Python work + page faults; nothing is parsed.

    python -m llmrec_b200.utility.csr_store <dataset dir>      # convert the json files in place (keeps them)
"""
from __future__ import annotations

import json
import os
from collections.abc import Mapping

import numpy as np

FORMAT = "llmrec-csr-v1"
META = "interactions.csr.json"
SPLITS = ("train", "test", "val")


def present(path: str) -> bool:
    return os.path.exists(os.path.join(path, META))


class CsrRows(Mapping):
    """Read-only `{uid: [items]}` view over CSR arrays; only users with a non-empty row are keys (the reference
    drops empty lists while loading, load_data.py:30-31,39-40,47-48)."""

    def __init__(self, rowptr, col, order=None):
        self.rowptr, self.col = rowptr, col
        self._order = order
        self._n = None

    def order(self) -> np.ndarray:
        if self._order is None:
            self._order = np.flatnonzero(np.diff(self.rowptr) > 0).astype(np.int32)
        return self._order

    def __getitem__(self, u):
        u = int(u)
        if not 0 <= u < self.rowptr.shape[0] - 1:
            raise KeyError(u)
        lo, hi = int(self.rowptr[u]), int(self.rowptr[u + 1])
        if lo == hi:
            raise KeyError(u)
        return self.col[lo:hi].tolist()

    def __contains__(self, u):
        try:
            u = int(u)
        except (TypeError, ValueError):
            return False
        return 0 <= u < self.rowptr.shape[0] - 1 and self.rowptr[u + 1] > self.rowptr[u]

    def __iter__(self):
        return iter(self.order().tolist())

    def __len__(self):
        if self._n is None:
            self._n = int(self.order().shape[0])
        return self._n


def _check_split(name, rowptr, col, n_users, n_items):
    if rowptr.ndim != 1 or rowptr.shape[0] != n_users + 1:
        raise ValueError("%s.rowptr: expected %d entries, found %s" % (name, n_users + 1, rowptr.shape))
    if int(rowptr[0]) != 0 or int(rowptr[-1]) != col.shape[0]:
        raise ValueError("%s: rowptr[0]=%d, rowptr[-1]=%d, nnz=%d" % (name, rowptr[0], rowptr[-1], col.shape[0]))
    if col.shape[0] >= 2 ** 31:
        raise ValueError("%s: %d interactions do not fit int32 offsets" % (name, col.shape[0]))


def write(path, splits, n_users, n_items, orders=None, validate=True):
    """splits: {name: (rowptr[n_users+1], col[nnz])}; orders: {name: user order} (optional)."""
    os.makedirs(path, exist_ok=True)
    names = []
    for name, (rowptr, col) in splits.items():
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        _check_split(name, rowptr, col, n_users, n_items)
        if validate and col.shape[0]:
            if np.any(np.diff(rowptr) < 0):
                raise ValueError("%s: rowptr is not non-decreasing" % name)
            if int(col.min()) < 0:
                raise ValueError("%s: negative item id" % name)
        np.save(os.path.join(path, name + ".rowptr.npy"), rowptr)
        np.save(os.path.join(path, name + ".col.npy"), col)
        o = (orders or {}).get(name)
        opath = os.path.join(path, name + ".order.npy")
        if o is not None:
            np.save(opath, np.ascontiguousarray(o, dtype=np.int32))
        elif os.path.exists(opath):
            os.remove(opath)
        names.append(name)
    with open(os.path.join(path, META), "w") as f:
        json.dump({"format": FORMAT, "n_users": int(n_users), "n_items": int(n_items), "splits": names}, f)


def read(path, mmap=True):
    """-> (meta, {name: CsrRows}); arrays stay memory-mapped (read-only) unless mmap=False."""
    with open(os.path.join(path, META)) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT:
        raise ValueError("%s: unknown format %r" % (os.path.join(path, META), meta.get("format")))
    mode = "r" if mmap else None
    out = {}
    for name in meta["splits"]:
        rowptr = np.load(os.path.join(path, name + ".rowptr.npy"), mmap_mode=mode)
        col = np.load(os.path.join(path, name + ".col.npy"), mmap_mode=mode)
        _check_split(name, rowptr, col, meta["n_users"], meta["n_items"])
        opath = os.path.join(path, name + ".order.npy")
        order = np.load(opath, mmap_mode=mode) if os.path.exists(opath) else None
        out[name] = CsrRows(rowptr, col, order)
    return meta, out


def rows_from_dict(d, n_users):
    """{"uid" | uid: [items]} -> (rowptr int64, col int32, order int32); empty lists are dropped like the reference does."""
    keys = np.fromiter((int(k) for k, v in d.items() if len(v)), dtype=np.int64)
    if keys.size and int(keys.max()) >= n_users:
        raise ValueError("user id %d outside [0, %d)" % (int(keys.max()), n_users))
    counts = np.zeros(n_users, dtype=np.int64)
    lens = np.fromiter((len(v) for v in d.values() if len(v)), dtype=np.int64, count=keys.size)
    if np.unique(keys).size != keys.size:
        raise ValueError("duplicate user keys")
    counts[keys] = lens
    rowptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    col = np.empty(int(rowptr[-1]), dtype=np.int32)
    for k, v in d.items():
        if len(v):
            u = int(k)
            col[rowptr[u]:rowptr[u + 1]] = v
    return rowptr, col, keys.astype(np.int32)


def convert_json(path, n_items=None):
    """Write the CSR form of `path`'s train/test/val json files next to them.  n_users follows the reference
    (max TRAIN uid + 1, load_data.py:35,55); n_items = rows of text_feat.npy when present (load_data.py:57-58)."""
    loaded = {}
    for name in SPLITS:
        p = os.path.join(path, name + ".json")
        if os.path.exists(p):
            with open(p) as f:
                loaded[name] = json.load(f)
    if "train" not in loaded:
        raise FileNotFoundError(os.path.join(path, "train.json"))
    n_users = max(int(k) for k, v in loaded["train"].items() if len(v)) + 1
    rows = n_users          # a test/val user >= n_users has no embedding row upstream either; rows_from_dict rejects it
    if n_items is None:
        tf = os.path.join(path, "text_feat.npy")
        if os.path.exists(tf):
            n_items = int(np.load(tf, mmap_mode="r").shape[0])
        else:
            n_items = max(max(v) for d in loaded.values() for v in d.values() if len(v)) + 1
    splits, orders = {}, {}
    for name, d in loaded.items():
        rowptr, col, order = rows_from_dict(d, rows)
        splits[name], orders[name] = (rowptr, col), order
    write(path, splits, rows, n_items, orders)
    return {"format": FORMAT, "n_users": rows, "n_items": n_items, "splits": list(splits)}

if __name__ == "__main__":
    import sys
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    print(convert_json(sys.argv[1]))
