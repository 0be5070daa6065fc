"""Device-side batch sampler (`--device_sampler 1`): Data.sample() (utility/load_data.py:157-195) and the
augmented-edge step (main.py:216-224) as one kernel that fills the engine's static index buffer, so a replayed training step needs no
host sampler and no H2D copy.  Same distributions as the reference, NOT its RNG streams -- the parity default stays the host replay
(host_native.BatchSampler)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import ops


class DeviceSampler:
    def __init__(self, exist_users, train_rowptr, train_col_sorted, n_items, batch_size, aug_pos, aug_neg, aug_limit, aug_rate, device, seed=0):
        dev = torch.device(device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
        self.exist, self.rowptr, self.col = t(exist_users), t(train_rowptr), t(train_col_sorted)
        self.n_items, self.batch = int(n_items), int(batch_size)
        self.n_aug = int(self.batch * aug_rate) if aug_pos is not None else 0          # int(len(users) * rate), main.py:218
        self.aug_pos = t(aug_pos) if aug_pos is not None else None
        self.aug_neg = t(aug_neg) if aug_neg is not None else None
        self.aug_limit = int(aug_limit)
        self.state = torch.tensor([int(seed), 0], dtype=torch.int64, device=dev)       # {seed, step}; the kernel advances step
        self.keys = torch.zeros(max(int(self.exist.numel()), self.batch), dtype=torch.int32, device=dev)

    def fill(self, index_buffer, meta_table):
        """index_buffer: the engine's [4 x cap] int32 buffer; meta_table: [cap + 1, 2] int32 {B', n_keep}.  One launch on the current stream."""
        cap = int(index_buffer.shape[1])
        N.check(N.lib().llmrec_device_sample_batch(
            C.c_void_p(self.exist.data_ptr()), self.exist.numel(), self.batch, C.c_void_p(self.rowptr.data_ptr()), C.c_void_p(self.col.data_ptr()),
            self.n_items, self.n_aug, C.c_void_p(self.aug_pos.data_ptr()) if self.n_aug else None, C.c_void_p(self.aug_neg.data_ptr()) if self.n_aug else None,
            self.aug_pos.numel() if self.n_aug else 0, self.aug_limit, C.c_void_p(meta_table.data_ptr()), cap, C.c_void_p(self.state.data_ptr()),
            C.c_void_p(index_buffer.data_ptr()), C.c_void_p(self.keys.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "device_sample_batch")
        ops._count()
