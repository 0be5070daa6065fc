"""Stand-alone probe of the tcgen05 projection kernels (run under `timeout`): prints max errors vs fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmrec_b200 import ops

torch.manual_seed(0)
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for (n, k, d) in [(300, 64, 64), (1000, 512, 64), (777, 96, 64), (17366, 1536, 64), (300, 1536, 128), (4096, 768, 128), (500, 256, 32)]:
    X = torch.randn(n, k, device=dev); W = torch.randn(d, k, device=dev) / k ** 0.5; b = torch.randn(d, device=dev)
    ref = X.double() @ W.double().t() + b.double()
    for mode in (1, 0):
        if which in ("all", "fwd"):
            wide = torch.zeros(n, 2 * d, device=dev); Y = wide[:, d:]
            ops.proj_fwd(X, W, b, Y, mode)
            torch.cuda.synchronize()
            err = (Y.double() - ref).abs().max().item()
            print(f"fwd  n={n} k={k} d={d} mode={mode} max_abs_err={err:.3e} (ref max {ref.abs().max().item():.2f}) untouched_cols_ok={float(wide[:, :d].abs().max()) == 0.0}", flush=True)
        if which in ("all", "wgrad") and d % 32 == 0:
            dY = torch.randn(n, d, device=dev)
            dW = torch.full((d, k), 7.0, device=dev); db = torch.full((d,), 7.0, device=dev)
            ops.proj_wgrad(X, dY, dW, db, False, mode)
            torch.cuda.synchronize()
            rw = dY.double().t() @ X.double()
            print(f"wgrad n={n} k={k} d={d} mode={mode} max_abs_err={(dW.double() - rw).abs().max().item():.3e} (ref max {rw.abs().max().item():.1f}) "
                  f"db_err={(db.double() - dY.double().sum(0)).abs().max().item():.3e}", flush=True)
print("probe done")
