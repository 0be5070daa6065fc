"""Roofline denominators and the algorithmic byte counts of SURVEY.md 8(d) (fp32, int32 indices), shared by bench.py's legs."""
from __future__ import annotations

import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    """-> (HBM GB/s, dense bf16 TFLOP/s, source): the driver-measured MEASURED_PEAKS.json, else B200_PROFILING.md's fallback."""
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def spmm_bytes(nnz, M, N, d, segs=1):
    """Y[M x d] = diag(s) R X[N x d] for `segs` operands sharing the pattern: every operand touched once."""
    return 4 * nnz + 4 * (M + 1) + 4 * M + segs * (4 * d * N + 4 * d * M)


def proj_bytes(n, k, d):
    return 4 * n * k + 4 * k * d + 4 * n * d


def step_bytes(hp, nnz):
    """Algorithmic bytes per kernel family for one training step of engine.HotPath `hp`."""
    nu, ni, d, S, L = hp.nu, hp.ni, hp.d, hp.S, hp.L
    out = {}
    if hp.has_feats:
        f = hp.feats
        gemms = [(ni, f["image"].shape[1]), (ni, f["text"].shape[1])] + [(ni, v.shape[1]) for v in f["item"].values()] + [(nu, f["user"].shape[1])]
        out["proj_fwd"] = out["proj_wgrad"] = sum(proj_bytes(n, k, d) for n, k in gemms)
    sp = lambda M, N, segs: spmm_bytes(nnz, M, N, d, segs)
    out["spmm_fwd"] = sp(nu, ni, S + 1) + sp(ni, nu, S + 2) + sp(nu, ni, 2) + (sp(ni, nu, 1) if L >= 2 else 0)
    out["spmm_bwd"] = sp(ni, nu, 1) + sp(nu, ni, S + 2) + sp(ni, nu, S + 1) + (sp(nu, ni, 1) + sp(ni, nu, 1) if L >= 2 else 0)
    out["adamw"] = 28 * sum(p.numel() for p in hp.opt.params)
    T = (3 + len(hp.keys)) if hp.has_feats else 0
    fuse = 4 * d * (nu + ni) * ((L + 1) + T + 1)
    out["fuse_fwd"] = fuse
    out["fuse_bwd"] = fuse + 4 * d * (nu + ni) * T
    return out
