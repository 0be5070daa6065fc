"""Stand-alone probe of the tcgen05 scoring + fused top-K kernel against the exact SIMT path (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from llmrec_b200 import ops
torch.manual_seed(0)
dev = "cuda"
cases = [(300, 1000, 64, 50), (200, 17366, 64, 50), (4096, 17366, 64, 50), (64, 100000, 128, 20), (1000, 5000, 32, 10)]
if len(sys.argv) > 1:
    cases = [cases[int(sys.argv[1])]]
for (nb, ni, d, K) in cases:
    nu = nb + 7
    U = torch.randn(nu, d, device=dev); I = torch.randn(ni, d, device=dev)
    users = torch.randperm(nu, device=dev)[:nb].to(torch.int32)
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 9, nu); rowptr = np.zeros(nu + 1, np.int32); rowptr[1:] = np.cumsum(lens)
    col = np.concatenate([np.sort(rng.choice(ni, size=l, replace=False)) for l in lens] + [np.zeros(0, int)]).astype(np.int32)
    rp, cl = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
    i2, v2 = ops.score_topk(U, I, users, rp, cl, K, mode=2, want_vals=True)
    torch.cuda.synchronize()
    i0, v0 = ops.score_topk(U, I, users, rp, cl, K, mode=0, want_vals=True)
    torch.cuda.synchronize()
    same = (i0 == i2).all(1).float().mean().item()
    t = time.perf_counter()
    for _ in range(3):
        ops.score_topk(U, I, users, rp, cl, K, mode=0)
    torch.cuda.synchronize(); t0 = (time.perf_counter() - t) / 3
    t = time.perf_counter()
    for _ in range(3):
        ops.score_topk(U, I, users, rp, cl, K, mode=2)
    torch.cuda.synchronize(); t2 = (time.perf_counter() - t) / 3
    print(f"nb={nb} ni={ni} d={d} K={K}: identical rows {same:.4f}  vals max diff {(v0 - v2).abs().max().item():.2e}  tc {t0*1e3:.2f} ms  simt {t2*1e3:.2f} ms", flush=True)
print("probe done")
