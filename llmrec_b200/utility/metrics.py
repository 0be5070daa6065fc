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


def block_metrics(hits, n_pos, Ks):
    """hits: uint8 [n x K_max] (rank order), n_pos: int [n] = len(test_set[u]).
    Returns dict of float64 [n x len(Ks)] arrays: precision, recall, ndcg, hit_ratio -- per user bit-identical to
    precision_at_k / recall_at_k / ndcg_at_k / hit_at_k (hits are 0/1, so means and sums are exact integer ratios; the
    ideal DCG of "m hits inside the retrieved list" comes from a table built with the scalar expression)."""
    hits = np.ascontiguousarray(hits)
    n, kmax = hits.shape
    npos = np.asarray(n_pos, dtype=np.float64)
    total_hits = hits.sum(axis=1, dtype=np.int64)                      # hits inside the whole retrieved list
    out = {k: np.zeros((n, len(Ks))) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    for j, K in enumerate(Ks):
        K = min(K, kmax)
        head = hits[:, :K]
        cnt = head.sum(axis=1, dtype=np.int64)
        cntf = cnt.astype(np.float64)
        out["precision"][:, j] = cntf / K
        with np.errstate(divide="ignore", invalid="ignore"):
            out["recall"][:, j] = np.where(npos == 0, 0.0, cntf / npos)
        out["hit_ratio"][:, j] = (cnt > 0).astype(np.float64)
        rows = np.nonzero(cnt)[0]                                      # dcg is 0 (and ndcg 0) without a hit in the head
        if rows.size:
            disc = np.log2(np.arange(2, K + 2))
            dcg = np.sum(np.ascontiguousarray(head[rows]).astype(np.float64) / disc, axis=1)
            best = _ideal_dcg_table(K)[np.minimum(total_hits[rows], K)]
            out["ndcg"][rows, j] = dcg / best
    return out
