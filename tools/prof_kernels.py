"""Micro-driver for ncu: one NF-shaped call of each hot kernel family (projection fwd / wgrad groups, SpMM launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch
from llmrec_b200 import ops
from llmrec_b200.graph import BipartiteGraph

torch.manual_seed(0)
dev = "cuda"
nu, ni, d = 13187, 17366, int(os.environ.get("D", 64))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(os.environ.get("REPS", 3))
mode = int(os.environ.get("MODE", 0))
if which in ("all", "proj"):
    dims = [(ni, 1536)] * 5 + [(nu, 1536), (ni, 768), (ni, 512)]
    Xs = [torch.randn(n, k, device=dev) for n, k in dims]
    Ws = {k: torch.randn(d, k, device=dev) / k ** 0.5 for k in (1536, 768, 512)}
    Wu = torch.randn(d, 1536, device=dev) / 39.0
    b = torch.zeros(d, device=dev)
    outs = [torch.empty(n, d, device=dev) for n, _ in dims]
    fw = [(X, (Wu if i == 5 else Ws[X.shape[1]]), b, o) for i, (X, o) in enumerate(zip(Xs, outs))]
    dWs = [torch.empty(d, X.shape[1], device=dev) for X in Xs]
    dbs = [torch.empty(d, device=dev) for _ in Xs]
    wg = [(X, o, dW, db_, False) for X, o, dW, db_ in zip(Xs, outs, dWs, dbs)]
    for _ in range(reps):
        ops.proj_fwd_group(fw, d, mode)
        ops.proj_wgrad_group(wg, d, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("fwd", lambda: ops.proj_fwd_group(fw, d, mode)), ("wgrad", lambda: ops.proj_wgrad_group(wg, d, mode))):
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        byts = sum(4 * n * k + 4 * k * d + 4 * n * d for n, k in dims)
        ms = e0.elapsed_time(e1) / 10
        print(f"proj_{name}: {ms:.4f} ms  {byts / ms / 1e6:.1f} GB/s", flush=True)
if which in ("all", "spmm"):
    rng = np.random.default_rng(0)
    w = 1.0 / np.power(np.arange(ni) + 8.0, 0.8); w /= w.sum()
    rows = rng.integers(0, nu, 43000); cols = rng.choice(ni, size=43000, p=w)
    m = sp.csr_matrix((np.ones(43000, np.float32), (rows, cols)), shape=(nu, ni)); m.sum_duplicates(); m.data[:] = 1
    for tile in (0, 8, 16, 32):
        g = BipartiteGraph(m, dev, tile_nnz=tile)
        for S in (8, 1):
            Xi = torch.randn(ni, S * d, device=dev); Yu = torch.empty(nu, S * d, device=dev); Yi = torch.empty(ni, S * d, device=dev)
            su = [(Xi[:, s * d:(s + 1) * d], Yu[:, s * d:(s + 1) * d], None, False) for s in range(S)]
            si = [(Yu[:, s * d:(s + 1) * d], Yi[:, s * d:(s + 1) * d], None, False) for s in range(S)]
            for _ in range(reps):
                g.ui.apply(su); g.iu.apply(si)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for name, op, segs in (("ui", g.ui, su), ("iu", g.iu, si)):
                e0.record()
                for _ in range(20):
                    op.apply(segs)
                e1.record(); torch.cuda.synchronize()
                print(f"spmm tile={tile} S={S} {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us  max_deg={int((op.rowptr[1:] - op.rowptr[:-1]).max())}", flush=True)
