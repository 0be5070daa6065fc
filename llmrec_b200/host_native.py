"""Host-side native helpers (no GPU involved): the bit-identical C sampler of csrc/host_sampler.cu."""
import numpy as np

from . import _native as N


def sample_items(users, train_rowptr, train_col, n_items):
    """(pos, neg) int32 arrays for `users`, consuming np.random's GLOBAL legacy MT19937 stream exactly like
    Data.sample()'s per-user np.random.randint calls (utility/load_data.py:166-187)."""
    name, key, pos, has_gauss, cached = np.random.get_state()
    if name != "MT19937":
        raise RuntimeError("np.random global state is not MT19937")
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    p = np.array([pos], dtype=np.int32)
    users = np.ascontiguousarray(users, dtype=np.int32)
    out_p = np.empty(users.shape[0], dtype=np.int32)
    out_n = np.empty(users.shape[0], dtype=np.int32)
    rc = N.lib().llmrec_host_sample_items(key.ctypes.data, p.ctypes.data, users.ctypes.data, users.shape[0],
                                          train_rowptr.ctypes.data, train_col.ctypes.data, int(n_items),
                                          out_p.ctypes.data, out_n.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"native sampler failed (rc={rc}): a sampled user has no train items or no possible negative")
    np.random.set_state((name, key, int(p[0]), has_gauss, cached))
    return out_p, out_n
