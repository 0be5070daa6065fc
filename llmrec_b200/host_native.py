"""Host-side native helpers (no GPU involved): the bit-identical C sampler of csrc/host_sampler.cu."""
import numpy as np

from . import _native as N


def _np_global_mt():
    """Address of numpy's GLOBAL legacy generator state, `struct { uint32_t key[624]; int pos; }` (numpy/random/src/mt19937/
    mt19937.h), through the bit generator's ctypes interface: the C samplers advance it in place, no get_state/set_state
    round trip (2 x 40 us).  None when the global generator is not an MT19937 exposing that interface."""
    try:
        bg = np.random.mtrand._rand._bit_generator
        if type(bg).__name__ != "MT19937":
            return None
        addr = bg.ctypes.state_address
        return addr if isinstance(addr, int) else None
    except Exception:
        return None


def sample_items(users, train_rowptr, train_col, n_items):
    """(pos, neg) int32 arrays for `users`, consuming np.random's GLOBAL legacy MT19937 stream exactly like
    Data.sample()'s per-user np.random.randint calls (utility/load_data.py:166-187)."""
    users = np.ascontiguousarray(users, dtype=np.int32)
    out_p = np.empty(users.shape[0], dtype=np.int32)
    out_n = np.empty(users.shape[0], dtype=np.int32)
    addr = _np_global_mt()
    if addr is not None:
        rc = N.lib().llmrec_host_sample_items(addr, addr + 624 * 4, users.ctypes.data, users.shape[0],
                                              train_rowptr.ctypes.data, train_col.ctypes.data, int(n_items),
                                              out_p.ctypes.data, out_n.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"native sampler failed (rc={rc}): a sampled user has no train items or no possible negative")
        return out_p, out_n
    name, key, pos, has_gauss, cached = np.random.get_state()
    if name != "MT19937":
        raise RuntimeError("np.random global state is not MT19937")
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    p = np.array([pos], dtype=np.int32)
    rc = N.lib().llmrec_host_sample_items(key.ctypes.data, p.ctypes.data, users.ctypes.data, users.shape[0],
                                          train_rowptr.ctypes.data, train_col.ctypes.data, int(n_items),
                                          out_p.ctypes.data, out_n.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"native sampler failed (rc={rc}): a sampled user has no train items or no possible negative")
    np.random.set_state((name, key, int(p[0]), has_gauss, cached))
    return out_p, out_n


def _sample_uses_pool(n, k):
    """Which branch of CPython's random.sample(population of n, k) runs (Lib/random.py: `n <= setsize`), computed with the
    interpreter's own float arithmetic so that it can never disagree with `random.sample` itself."""
    from math import ceil, log
    setsize = 21
    if k > 5:
        setsize += 4 ** ceil(log(k * 3, 4))
    return n <= setsize


class BatchSampler:
    """Whole training batches (users, positives, rejection-sampled negatives, augmented edges) drawn in one C call,
    consuming the GLOBAL `random` and `np.random` streams exactly like Data.sample() + main.py:213-224 do in Python.

    exist_users / train CSR: from `Data`; aug_pos / aug_neg: int32[n_users] tables of augmented_sample_dict
    (INT32_MIN where a uid is missing).  `draw(out, aug_rate)` fills a [3 x ld] int32 array (rows users, pos, neg --
    typically pinned staging memory) and returns B'."""

    MISSING = np.iinfo(np.int32).min

    def __init__(self, exist_users, train_rowptr, train_col, n_items, batch_size, aug_pos=None, aug_neg=None, aug_limit=None):
        self.exist = np.ascontiguousarray(exist_users, dtype=np.int32)
        self.rowptr = np.ascontiguousarray(train_rowptr, dtype=np.int32)
        self.col = np.ascontiguousarray(train_col, dtype=np.int32)
        self.n_items, self.batch = int(n_items), int(batch_size)
        self.aug_limit = int(n_items if aug_limit is None else aug_limit)
        n = self.exist.shape[0]
        self.users_pool = self.batch <= n and _sample_uses_pool(n, self.batch)
        self.stamp = np.zeros(max(n, 1), dtype=np.int32)
        self.epoch = 0
        self.pool = np.empty(max(n if self.users_pool else 0, 2 * self.batch) + 8, dtype=np.int32)
        self.set_aug(aug_pos, aug_neg)
        self._py_key = np.empty(624, dtype=np.uint32)
        self._pos = np.zeros(4, dtype=np.int32)                 # [py_pos, np_pos, n_out, -]

    def set_aug(self, aug_pos, aug_neg):
        self.aug_pos = None if aug_pos is None else np.ascontiguousarray(aug_pos, dtype=np.int32)
        self.aug_neg = None if aug_neg is None else np.ascontiguousarray(aug_neg, dtype=np.int32)

    @staticmethod
    def aug_tables(aug_dict, n_users, n_items=None):
        """{uid: {0: pos, 1: neg}} (main.py:216-220) -> two int32[n_users] tables.
        Upstream a NEGATIVE id passes the `< n_items` filter (main.py:219-221) and then wraps around under Python / torch negative
        indexing (row -1 = the last item).  The kernels take row ids literally, so with n_items given the wrap is applied here, once;
        ids below -n_items (an IndexError upstream) are marked missing."""
        pos = np.full(n_users, BatchSampler.MISSING, dtype=np.int32)
        neg = np.full(n_users, BatchSampler.MISSING, dtype=np.int32)

        def wrap(i):
            i = int(i)
            if i < 0 and n_items is not None:
                return i + n_items if i >= -n_items else BatchSampler.MISSING
            return i
        for u, pn in aug_dict.items():
            if 0 <= u < n_users:
                pos[u], neg[u] = wrap(pn[0]), wrap(pn[1])
        return pos, neg

    def draw(self, out, aug_rate=0.0):
        import random
        if out.dtype != np.int32 or out.ndim != 2 or out.shape[0] != 3 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous int32 [3 x ld] array")
        n_aug = int(self.batch * aug_rate) if self.aug_pos is not None else 0          # int(len(users) * rate), main.py:218
        if out.shape[1] < self.batch + n_aug:
            raise ValueError("staging buffer too small")
        version, py_state, gauss = random.getstate()
        if version != 3:
            raise RuntimeError("unexpected random.getstate() version")
        self._py_key[:] = py_state[:624]
        st = self._pos
        np_addr = _np_global_mt()
        if np_addr is not None:                                  # advance numpy's state in place
            np_key_ptr, np_pos_ptr = np_addr, np_addr + 624 * 4
        else:
            name, np_key, np_pos, has_gauss, cached = np.random.get_state()
            if name != "MT19937":
                raise RuntimeError("np.random global state is not MT19937")
            np_key = np.ascontiguousarray(np_key, dtype=np.uint32).copy()
            st[1] = np_pos
            np_key_ptr, np_pos_ptr = np_key.ctypes.data, st[1:].ctypes.data
        st[0] = py_state[624]
        self.epoch += 1
        if self.epoch >= 2 ** 31 - 1:
            self.stamp[:] = 0
            self.epoch = 1
        nil = 0
        rc = N.lib().llmrec_host_sample_batch(
            self._py_key.ctypes.data, st[0:].ctypes.data, np_key_ptr, np_pos_ptr,
            self.exist.ctypes.data, self.exist.shape[0], self.batch, 1 if self.users_pool else 0,
            self.rowptr.ctypes.data, self.col.ctypes.data, self.n_items,
            n_aug, 1 if (n_aug and _sample_uses_pool(self.batch, n_aug)) else 0,
            self.aug_pos.ctypes.data if n_aug else nil, self.aug_neg.ctypes.data if n_aug else nil,
            self.aug_pos.shape[0] if n_aug else 0, self.aug_limit,
            self.stamp.ctypes.data, self.epoch, self.pool.ctypes.data, out.ctypes.data, out.shape[1], st[2:].ctypes.data)
        if rc == 4:
            raise KeyError("a sampled user is missing from augmented_sample_dict")
        if rc != 0:
            raise RuntimeError(f"native batch sampler failed (rc={rc}): a sampled user has no train items or no possible negative")
        random.setstate((version, tuple(self._py_key.tolist()) + (int(st[0]),), gauss))
        if np_addr is None:
            np.random.set_state((name, np_key, int(st[1]), has_gauss, cached))
        return int(st[2])
