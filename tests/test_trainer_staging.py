"""Host-side batch staging of Trainer (sample -> pinned ring -> device views) without a GPU: the CUDA pieces (pinned
allocation, events, the step itself) are replaced by CPU stand-ins; the logic under test is which batch lands where."""
import os
import pickle
import random
from types import SimpleNamespace

import numpy as np
import torch

from llmrec_b200 import main as M
from llmrec_b200.host_native import BatchSampler
from llmrec_b200.utility.load_data import Data


class _Slot:
    syncs = 0

    def __init__(self, cap):
        self.host = torch.zeros((4, cap), dtype=torch.int32)
        self.np = self.host.numpy()
        self.event = SimpleNamespace(synchronize=self._sync, record=lambda: None)

    def _sync(self):
        _Slot.syncs += 1


class _Hot:
    """Stand-in for engine.HotPath's static index buffer (what Trainer stages batches into)."""

    def __init__(self, batch):
        self.batch, self.buf = batch, None

    def batch_capacity(self):
        return (self.batch + int(self.batch * 0.1) + 7) // 8 * 8

    def index_buffer(self, need):
        need = max(int(need), self.batch_capacity())
        if self.buf is None or self.buf.shape[1] < need:
            self.buf = torch.zeros((4, need), dtype=torch.int32)
        return self.buf

    def meta_row(self, B):
        return int(B), int((1 - 0.71) * B)


def _trainer(tiny_root, sampler, monkeypatch, batch=128):
    ddir = os.path.join(tiny_root, "netflix_valid_item")
    gen = Data(ddir, batch, sampler=sampler)
    tr = M.Trainer.__new__(M.Trainer)
    tr.args = SimpleNamespace(aug_sample_rate=0.1)
    tr.data_generator, tr.device, tr.batch_size = gen, torch.device("cpu"), batch
    tr.n_users, tr.n_items = gen.n_users, gen.n_items
    tr.augmented_sample_dict = pickle.load(open(os.path.join(ddir, "augmented_sample_dict"), "rb"))
    tr._slots, tr._slot_i, tr._idx_dev, tr._batch_sampler = [], 0, None, None
    tr.hot = _Hot(batch)
    if sampler == "native":
        rp, col = gen.csr("train")
        tr._batch_sampler = BatchSampler(gen.exist_users, rp, col, gen.n_items, batch,
                                         *BatchSampler.aug_tables(tr.augmented_sample_dict, gen.n_users), aug_limit=tr.n_items)
        tr._batch_np = np.empty((3, 2 * batch + 8), dtype=np.int32)
    monkeypatch.setattr(M, "_StagingSlot", _Slot)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    return tr


def test_staged_batches_equal_sampled_batches(tiny_root, golden, monkeypatch):
    py = _trainer(tiny_root, "python", monkeypatch)
    nat = _trainer(tiny_root, "native", monkeypatch)
    random.seed(2022); np.random.seed(2022)
    want = [py.sample_batch() for _ in range(7)]
    for b in range(3):
        np.testing.assert_array_equal(np.asarray(want[b]), golden[f"sampler/{b}"])
    random.seed(2022); np.random.seed(2022)
    got_lists = [nat.sample_batch() for _ in range(7)]
    assert got_lists == want and all(isinstance(x, list) for x in got_lists[0])
    random.seed(2022); np.random.seed(2022)
    for w in want:                                                   # fused path: more batches than ring slots
        u, p, n = nat.stage_batch()
        assert (u.tolist(), p.tolist(), n.tolist()) == tuple(w)
        assert nat.new_batch_size == len(w[0]) - 128
    assert len(nat._slots) == 4 and _Slot.syncs >= 7
    random.seed(2022); np.random.seed(2022)
    for w in want[:2]:                                               # python-sampler trainer goes through lists
        u, p, n = py.stage_batch()
        assert (u.tolist(), p.tolist(), n.tolist()) == tuple(w)


def test_upload_grows_the_ring(tiny_root, monkeypatch):
    tr = _trainer(tiny_root, "native", monkeypatch)
    u, p, n = tr.upload_batch([1, 2, 3], [4, 5, 6], [7, 8, 9])
    assert (u.tolist(), p.tolist(), n.tolist()) == ([1, 2, 3], [4, 5, 6], [7, 8, 9])
    cap0 = tr._slots[0].host.shape[1]
    big = list(range(cap0 + 5))
    u, p, n = tr.upload_batch(big, big, big)
    assert tr._slots[0].host.shape[1] >= cap0 + 5 and u.tolist() == big and tr._idx_dev.shape[1] >= cap0 + 5
