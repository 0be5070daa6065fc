"""Interaction data + BPR triplet sampler with the reference's interface (utility/load_data.py:10-202).

`Data(path, batch_size)` exposes n_users / n_items / n_train / n_test / exist_users / train_items /
test_set / val_set / R and `sample()`.  Differences from the reference, all behaviour-preserving:
  * `R` is assembled vectorised as CSR and only converted to dok on first access (the reference fills
    a dok_matrix one interaction at a time in Python at import, load_data.py:63-74; LLMRec never
    reads it);
  * CSR copies of train_items / test_set / val_set are kept for the device-side masking and hit
    lookup kernels;
  * `sample()` consumes the GLOBAL `random` and `np.random` streams in exactly the reference's order
    (load_data.py:157-195), so seeded runs draw identical batches.  `sampler="native"` runs the same
    algorithm in C (llmrec_b200/csrc/host_sampler.c): MT19937 + numpy's legacy masked-rejection
    bounded integers, state handed over with np.random.get_state()/set_state() -- bit-identical.
  * when the directory holds the binary CSR form of the three json files (utility/csr_store.py)
    it is memory-mapped instead of parsed: same attributes, `train_items / test_set / val_set` are read-only
    mappings over the arrays, no per-interaction Python work (the json walk is minutes at 200 M edges).
"""
import json
import os
import random as rd

import numpy as np
import scipy.sparse as sp

from . import csr_store


def _csr_from_dict(d, n_rows):
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    for u, items in d.items():
        rowptr[u + 1] = len(items)
    np.cumsum(rowptr, out=rowptr)
    col = np.empty(int(rowptr[-1]), dtype=np.int32)
    for u, items in d.items():
        col[rowptr[u]:rowptr[u + 1]] = items
    return rowptr.astype(np.int32), col


class Data(object):
    def __init__(self, path, batch_size, sampler="python"):
        self.path, self.batch_size = path, batch_size
        self.n_users = self.n_items = self.n_train = self.n_test = 0
        self.neg_pools = {}
        self.exist_users = []
        self._R = None
        self._csr = {}
        self._sampler = sampler
        self._native = None
        if csr_store.present(path):
            self._init_from_csr(path)
            return

        def load(name):
            with open(os.path.join(path, name + ".json")) as f:
                return json.load(f)

        train, test, val = load("train"), load("test"), load("val")
        self.train_items, self.test_set, self.val_set = {}, {}, {}
        max_item = 0
        for key, items in train.items():                       # load_data.py:29-36,66-74
            if len(items) == 0:
                continue
            uid = int(key)
            self.exist_users.append(uid)
            self.train_items[uid] = items
            max_item = max(max_item, max(items))
            self.n_users = max(self.n_users, uid)
            self.n_train += len(items)
        for key, items in test.items():                        # :38-44,76-83
            if len(items) == 0:
                continue
            max_item = max(max_item, max(items))
            self.n_test += len(items)
            self.test_set[int(key)] = items
        for key, items in val.items():                         # :46-52,85-92 (n_val is never defined upstream)
            if len(items) == 0:
                continue
            max_item = max(max_item, max(items))
            self.val_set[int(key)] = items
        self.n_users += 1
        text = np.load(os.path.join(path, "text_feat.npy"), mmap_mode="r")
        self.n_items = int(text.shape[0])                      # :57-58 overrides the json maximum
        self.print_statistics()

    def _init_from_csr(self, path):
        meta, rows = csr_store.read(path)
        if "train" not in rows:
            raise ValueError("%s: the CSR store has no train split" % path)
        empty = csr_store.CsrRows(np.zeros(meta["n_users"] + 1, dtype=np.int64), np.zeros(0, dtype=np.int32))
        self.train_items = rows["train"]
        self.test_set, self.val_set = rows.get("test", empty), rows.get("val", empty)
        self.exist_users = self.train_items.order().tolist()
        self.n_users = int(meta["n_users"])                    # = max train uid + 1 (csr_store.convert_json)
        self.n_train = int(self.train_items.col.shape[0])
        self.n_test = int(self.test_set.col.shape[0])
        text = os.path.join(path, "text_feat.npy")
        self.n_items = int(np.load(text, mmap_mode="r").shape[0]) if os.path.exists(text) else int(meta["n_items"])
        for which, r in (("train", self.train_items), ("test", self.test_set), ("val", self.val_set)):
            self._csr[which] = (np.asarray(r.rowptr, dtype=np.int32), np.asarray(r.col, dtype=np.int32))
        self.print_statistics()

    # -- matrices ---------------------------------------------------------------------------------
    def csr(self, which="train", sorted_rows=False):
        """(rowptr int32[n_users+1], col int32[nnz]) of train_items / test_set / val_set.  Rows keep the JSON order
        (the sampler indexes into them); sorted_rows=True returns a copy with every row ascending (device masks)."""
        if which not in self._csr:
            src = {"train": self.train_items, "test": self.test_set, "val": self.val_set}[which]
            self._csr[which] = _csr_from_dict(src, self.n_users)
        if not sorted_rows:
            return self._csr[which]
        key = which + ":sorted"
        if key not in self._csr:
            rp, col = self._csr[which]
            rows = np.repeat(np.arange(self.n_users), np.diff(rp))
            order = np.lexsort((col, rows))
            self._csr[key] = (rp, np.ascontiguousarray(col[order]))
        return self._csr[key]

    @property
    def R(self):
        if self._R is None:
            rowptr, col = self.csr("train")
            m = sp.csr_matrix((np.ones(col.shape[0], dtype=np.float32), col, rowptr), shape=(self.n_users, self.n_items))
            self._R = m.todok()
        return self._R

    # -- sampler ------------------------------------------------------------------------------------
    def sample(self):
        if self.batch_size <= self.n_users:
            users = rd.sample(self.exist_users, self.batch_size)
        else:
            users = [rd.choice(self.exist_users) for _ in range(self.batch_size)]
        if self._sampler == "native":
            return self._sample_items_native(users)
        pos_items, neg_items = [], []
        randint = np.random.randint
        for u in users:
            mine = self.train_items[u]
            pos_items.append(mine[randint(low=0, high=len(mine), size=1)[0]])      # 1 positive (:167-178)
            while True:                                                            # 1 rejection-sampled negative (:180-187)
                neg = randint(low=0, high=self.n_items, size=1)[0]
                if neg not in mine:
                    neg_items.append(neg)
                    break
        return users, pos_items, neg_items

    def _sample_items_native(self, users):
        from ..host_native import sample_items
        rowptr, col = self.csr("train")
        pos, neg = sample_items(np.asarray(users, dtype=np.int32), rowptr, col, self.n_items)
        return users, pos.tolist(), neg.tolist()

    def print_statistics(self):
        print("n_users=%d, n_items=%d" % (self.n_users, self.n_items))
        print("n_interactions=%d" % (self.n_train + self.n_test))
        print("n_train=%d, n_test=%d, sparsity=%.5f" % (self.n_train, self.n_test,
                                                        (self.n_train + self.n_test) / (self.n_users * self.n_items)))
