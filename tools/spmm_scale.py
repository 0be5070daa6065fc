"""SpMM at the large synthetic scale on one GPU: times both gather directions for the current LLMREC_SPMM_VARIANT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmrec_b200.dist import ShardedGraph, synthetic_shard
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
dev = torch.device("cuda")
nu, ni, ne, d = int(10_000_000 * scale), int(1_000_000 * scale), int(200_000_000 * scale), 128
ul, it, lo, hi = synthetic_shard(nu, ni, ne, 0, 1, dev)
g = ShardedGraph(ul, it, nu, ni, tile_nnz=int(os.environ.get("TILE", 0)))
del ul, it
Xi, Xu = torch.randn(ni, d, device=dev), torch.randn(nu, d, device=dev)
Yu, Yi = torch.empty(nu, d, device=dev), torch.empty(ni, d, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, op, seg in (("ui (gather items)", g.ui, [(Xi, Yu, None, False)]), ("iu (gather users)", g.iu_raw, [(Xu, Yi, None, False)]),
                      ("iuT (gather items, weighted)", g.iuT, [(Xi, Yu, None, False)]), ("uiT (gather users, weighted)", g.uiT_raw, [(Xu, Yi, None, False)])):
    op.apply(seg); torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        op.apply(seg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"variant={os.environ.get('LLMREC_SPMM_VARIANT','0')} tile={op.plan.tile_nnz} {name}: {ms:.3f} ms  gather {(g.nnz * (4 * d + 4)) / ms / 1e6:.0f} GB/s  nnz={g.nnz} tiles={op.plan.n_tiles} split={op.plan.n_split}", flush=True)

# Column windows: the item table ([ni x 128] fp32 = 512 MB at full scale) does not fit the 126 MB L2, a 32-column window of it
# (128 MB) nearly does.  Same product as W launches over column slices (views with the full leading dimension): more index
# traffic (col re-read per window) against a higher L2 hit rate on the gathered rows.  COLWIN="2,4" to try.
for W in [int(x) for x in os.environ.get("COLWIN", "").split(",") if x]:
    w = d // W
    for name, op, X, Y in (("ui (gather items)", g.ui, Xi, Yu), ("iu (gather users)", g.iu_raw, Xu, Yi)):
        ref = torch.empty_like(Y)
        op.apply([(X, ref, None, False)])
        run = lambda: [op.apply([(X[:, j * w:(j + 1) * w], Y[:, j * w:(j + 1) * w], None, False)]) for j in range(W)]
        run(); torch.cuda.synchronize()
        same = bool(torch.equal(Y, ref))
        e0.record()
        for _ in range(3):
            run()
        e1.record(); torch.cuda.synchronize()
        print(f"colwin={W} {name}: {e0.elapsed_time(e1) / 3:.3f} ms  identical={same}", flush=True)
