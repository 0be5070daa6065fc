"""Vectorised metrics == the reference's scalar functions, bit for bit (utility/metrics.py)."""
import numpy as np

from llmrec_b200.utility import metrics as M
from oracle import llmrec_oracle as O


def test_block_metrics_bit_identical_to_scalar_and_oracle():
    rng = np.random.default_rng(0)
    Ks = [10, 20, 50]
    for dens in (0.0, 0.02, 0.3, 0.9):
        hits = (rng.random((500, 50)) < dens).astype(np.uint8)
        npos = rng.integers(0, 6, 500)
        got = M.block_metrics(hits, npos, Ks)
        for i in range(500):
            r = hits[i].tolist()
            for j, K in enumerate(Ks):
                assert got["precision"][i, j] == M.precision_at_k(r, K)
                assert got["recall"][i, j] == M.recall_at_k(r, K, int(npos[i]))
                assert got["ndcg"][i, j] == M.ndcg_at_k(r, K)
                assert got["hit_ratio"][i, j] == M.hit_at_k(r, K)
            om = O.user_metrics(r, int(npos[i]), Ks)
            for k in ("precision", "recall", "ndcg", "hit_ratio"):
                assert (got[k][i] == om[k]).all()


def test_known_answers():
    r = [0, 1, 0, 0, 1] + [0] * 45
    assert M.ndcg_at_k(r, 10) == 0.6240505200038379 and M.precision_at_k(r, 10) == 0.2
    assert abs(M.recall_at_k(r, 10, 3) - 2 / 3) < 1e-15 and M.hit_at_k(r, 10) == 1.0
    r2 = [0] * 15 + [1] + [0] * 34
    assert M.ndcg_at_k(r2, 10) == 0.0 and M.ndcg_at_k(r2, 20) == 0.24465054211822604
