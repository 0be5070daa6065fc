"""Ranking metrics of the reference (utility/metrics.py), scalar API plus a vectorised block form.

The per-user functions keep the reference's names and float64 arithmetic.  `block_metrics` evaluates
a [n x K_max] 0/1 hit matrix for all users at once with the SAME numpy reductions per row, so the
per-user values are bit-identical to the scalar functions (checked in tests/test_metrics.py).
Note ndcg's ideal DCG is built from the hits inside the retrieved list (metrics.py:68-78).
"""
import numpy as np


def precision_at_k(r, k):
    assert k >= 1
    return np.mean(np.asarray(r)[:k])


def dcg_at_k(r, k, method=1):
    r = np.asarray(r, dtype=np.float64)[:k]
    if r.size:
        if method == 0:
            return r[0] + np.sum(r[1:] / np.log2(np.arange(2, r.size + 1)))
        if method == 1:
            return np.sum(r / np.log2(np.arange(2, r.size + 2)))
        raise ValueError("method must be 0 or 1.")
    return 0.0


def ndcg_at_k(r, k, method=1):
    best = dcg_at_k(sorted(r, reverse=True), k, method)
    if not best:
        return 0.0
    return dcg_at_k(r, k, method) / best


def recall_at_k(r, k, all_pos_num):
    if all_pos_num == 0:
        return 0
    return np.sum(np.asarray(r, dtype=np.float64)[:k]) / all_pos_num


def hit_at_k(r, k):
    return 1.0 if np.sum(np.array(r)[:k]) > 0 else 0.0


def F1(pre, rec):
    return (2.0 * pre * rec) / (pre + rec) if pre + rec > 0 else 0.0


def auc(ground_truth, prediction):
    try:
        from sklearn.metrics import roc_auc_score
        return roc_auc_score(y_true=ground_truth, y_score=prediction)
    except Exception:
        return 0.0


_IDEAL_CACHE = {}


def _ideal_dcg_table(K):
    """best[m] = dcg_at_k of m ones followed by zeros, evaluated with the SAME numpy expression the scalar path uses."""
    if K not in _IDEAL_CACHE:
        disc = np.log2(np.arange(2, K + 2))
        rows = (np.arange(K)[None, :] < np.arange(K + 1)[:, None]).astype(np.float64)
        _IDEAL_CACHE[K] = np.array([np.sum(rows[m] / disc) for m in range(K + 1)])
    return _IDEAL_CACHE[K]


def block_metrics_sparse(hits, n_pos, Ks):
    """hits: uint8 [n x K_max] (rank order), n_pos: int [n] = len(test_set[u]).
    -> (rows, m): rows = ascending indices of the users with at least one hit in the retrieved list, m = dict of float64
    [len(rows) x len(Ks)] arrays precision / recall / ndcg / hit_ratio for those users; every other user's metrics are
    exactly 0.0.  Per user bit-identical to precision_at_k / recall_at_k / ndcg_at_k / hit_at_k: hits are 0/1, so means and
    sums are exact integer ratios; DCG is the scalar expression evaluated on the user's row; the ideal DCG of "m hits inside
    the retrieved list" comes from a table built with the scalar expression.  Work beyond one word-wise scan of the hit matrix is proportional to the number of users WITH hits."""
    hits = np.ascontiguousarray(hits, dtype=np.uint8)
    n, kmax = hits.shape
    # users with any hit: OR-reduce each row as 64-bit words (rows padded to a multiple of 8 bytes)
    w = -(-kmax // 8) * 8
    if w != kmax:
        padded = np.zeros((n, w), dtype=np.uint8)
        padded[:, :kmax] = hits
    else:
        padded = hits
    rows = np.flatnonzero(np.bitwise_or.reduce(padded.view(np.uint64), axis=1)) if n else np.zeros(0, dtype=np.int64)
    out = {k: np.zeros((rows.size, len(Ks))) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    if rows.size == 0:
        return rows, out
    sub = hits[rows]                                                   # [len(rows) x K_max]
    total_hits = sub.sum(axis=1, dtype=np.int64)                       # hits inside the whole retrieved list
    npos = np.asarray(n_pos, dtype=np.float64)[rows]
    for j, K in enumerate(Ks):
        K = min(K, kmax)
        head = sub[:, :K]
        cnt = head.sum(axis=1, dtype=np.int64)
        cntf = cnt.astype(np.float64)
        out["precision"][:, j] = cntf / K
        with np.errstate(divide="ignore", invalid="ignore"):
            out["recall"][:, j] = np.where(npos == 0, 0.0, cntf / npos)
        out["hit_ratio"][:, j] = (cnt > 0).astype(np.float64)
        live = np.nonzero(cnt)[0]                                      # dcg is 0 (and ndcg 0) without a hit in the head
        if live.size:
            disc = np.log2(np.arange(2, K + 2))
            dcg = np.sum(np.ascontiguousarray(head[live]).astype(np.float64) / disc, axis=1)
            best = _ideal_dcg_table(K)[np.minimum(total_hits[live], K)]
            out["ndcg"][live, j] = dcg / best
    return rows, out


def block_metrics(hits, n_pos, Ks):
    """Dense form of block_metrics_sparse: dict of float64 [n x len(Ks)] arrays."""
    rows, m = block_metrics_sparse(hits, n_pos, Ks)
    n = np.asarray(hits).shape[0]
    out = {k: np.zeros((n, len(Ks))) for k in m}
    for k in m:
        out[k][rows] = m[k]
    return out
