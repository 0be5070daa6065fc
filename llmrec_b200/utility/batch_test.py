"""Full-catalog evaluation with the reference's `test_torch` signature (utility/batch_test.py:112-169).

The reference scores a block of users on the device, copies the whole [2048 x n_items] block to
the host and ranks every user in Python (set difference + heapq.nlargest, ~3.7 ms/user).  Here
scoring, train-item masking and top-max(Ks) selection (ties -> lowest item id) are ONE device op
(`llmrec_score_topk_f32`), the hit lookup another, and only the [n x max(Ks)] 0/1 hit matrix crosses
PCIe; metrics are the reference's float64 formulas applied to that matrix.
"""
import numpy as np
import torch

from .. import ops
from ..runtime import get_args
from . import metrics

data_generator = None          # module-global like the reference (batch_test.py:16); set by init()
Ks = [10, 20, 50]
BATCH_SIZE = 1024
_dev_cache = {}


def init(generator, args=None):
    global data_generator, Ks, BATCH_SIZE, USR_NUM, ITEM_NUM, N_TRAIN, N_TEST
    args = args or get_args()
    data_generator = generator
    Ks = eval(args.Ks)
    BATCH_SIZE = args.batch_size
    USR_NUM, ITEM_NUM = generator.n_users, generator.n_items
    N_TRAIN, N_TEST = generator.n_train, generator.n_test
    _dev_cache.clear()


def _device_csr(which, device):
    key = (which, str(device))
    if key not in _dev_cache:
        rp, col = data_generator.csr(which, sorted_rows=True)
        _dev_cache[key] = (torch.from_numpy(rp).to(device), torch.from_numpy(col).to(device))
    return _dev_cache[key]


def rank_block(ua_embeddings, ia_embeddings, user_batch, is_val, mode=None):
    """-> (top-K ids int32 [b x Kmax] on device, hits uint8 [b x Kmax] on device)."""
    dev = ua_embeddings.device
    args = get_args()
    mode = ops.SCORE_MODE.get(getattr(args, "proj_mode", "3xtf32"), 0) if mode is None else mode
    users = torch.as_tensor(np.asarray(user_batch, dtype=np.int32)).to(dev, non_blocking=True)
    mrp, mcol = _device_csr("train", dev)
    trp, tcol = _device_csr("val" if is_val else "test", dev)
    idx = ops.score_topk(ua_embeddings, ia_embeddings, users, mrp, mcol, max(Ks), mode=mode)
    hits = ops.topk_hits(idx, users, trp, tcol)
    return idx, hits


def test_torch(ua_embeddings, ia_embeddings, users_to_test, is_val, drop_flag=False, batch_test_flag=False):
    nK = len(Ks)
    result = {"precision": np.zeros(nK), "recall": np.zeros(nK), "ndcg": np.zeros(nK), "hit_ratio": np.zeros(nK), "auc": 0.0}
    full = get_args().test_flag != "part"            # 'full': the same hit vectors + per-user ROC-AUC over every candidate (batch_test.py:38-68)
    test_users = np.asarray(list(users_to_test) if not isinstance(users_to_test, np.ndarray) else users_to_test, dtype=np.int32)
    n_test_users = int(test_users.shape[0])
    u_batch_size = BATCH_SIZE * 2                                             # batch_test.py:117
    if ops.SCORE_MODE.get(getattr(get_args(), "proj_mode", "3xtf32"), 0) != 2:
        u_batch_size = max(u_batch_size, 32768)       # no score block is materialised: larger user blocks, same results
    truth = data_generator.val_set if is_val else data_generator.test_set
    ua = ua_embeddings.detach()
    ia = ia_embeddings.detach()
    count = 0
    pending = []
    for start in range(0, n_test_users, u_batch_size):
        user_batch = test_users[start:start + u_batch_size]
        _, hits = rank_block(ua, ia, user_batch, is_val)
        auc = None
        if full:
            dev = ua.device
            users_dev = torch.as_tensor(np.asarray(user_batch, dtype=np.int32)).to(dev)
            auc = ops.user_auc(ua, ia, users_dev, *_device_csr("train", dev), *_device_csr("val" if is_val else "test", dev))
        pending.append((user_batch, hits, auc))
    for user_batch, hits, auc in pending:                                     # one D2H per block, after all launches
        h = hits.cpu().numpy()
        if auc is not None:
            for a in auc.cpu().numpy().astype(np.float64):                    # `result['auc'] += re['auc'] / n_test_users` (:165)
                result["auc"] += a / n_test_users
        trp = data_generator.csr("val" if is_val else "test")[0]
        n_pos = (trp[user_batch + 1] - trp[user_batch]).astype(np.int64)       # len(truth[u]) per user
        rows, m = metrics.block_metrics_sparse(h, n_pos, Ks)
        for k in ("precision", "recall", "ndcg", "hit_ratio"):
            # sequential float64 accumulation in user order == the reference's `+= re[k] / n` loop (:160-165); users without a
            # hit contribute exactly +0.0, which leaves the accumulator's bits unchanged, so only the others are walked
            acc = np.cumsum(np.vstack([result[k][None, :], m[k] / n_test_users]), axis=0)
            result[k] = acc[-1]
        count += len(user_batch)
    assert count == n_test_users
    return result
