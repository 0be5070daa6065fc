"""-m gpu, opt-in: code paths that are built and wired but have not been through a GPU run yet (written after the round's
GPU budget was spent).  They run only with LLMREC_TEST_EXPERIMENTAL=1 so that an unvalidated path cannot mask the
validated suite; once green on a B200 they move into test_kernels_gpu.py / test_path_gpu.py.

    LLMREC_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -q -m gpu
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LLMREC_TEST_EXPERIMENTAL") != "1", reason="set LLMREC_TEST_EXPERIMENTAL=1")]
cuda = "cuda"


@pytest.mark.parametrize("n,k", [(1, 32), (333, 96), (4097, 544)])
def test_panelize_round_trip(n, k):
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(n)
    wide = torch.randn(n, k + 8, generator=g).to(cuda)
    X = wide[:, 4:4 + k]                                   # leading dimension != k
    P = ops.PanelFeat(X)
    n_pad = (n + 127) // 128 * 128
    assert P.shape == (n, k) and tuple(P.data.shape) == ((k // 32) * n_pad, 32)
    assert torch.equal(P.rows(), X.contiguous())
    assert float(P.data.view(k // 32, n_pad, 32)[:, n:].abs().sum()) == 0.0          # zero padding


# tails on purpose: n % 128 != 0 (fwd row tiles), n % 32 != 0 (wgrad row blocks), k % 128 != 0 (wgrad feature tiles)
@pytest.mark.parametrize("n,k,d", [(100, 32, 32), (1000, 512, 64), (4133, 544, 64), (2500, 768, 128), (17366, 1536, 64)])
@pytest.mark.parametrize("mode", [0, 1])
def test_panel_layout_equals_row_layout(n, k, d, mode):
    """Same MMAs in the same order on the same tile images: results must be bit-identical to the row-major path."""
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(k + d)
    X = torch.randn(n, k, generator=g).to(cuda)
    W = (torch.randn(d, k, generator=g) / k ** 0.5).to(cuda)
    b = torch.randn(d, generator=g).to(cuda)
    dY = torch.randn(n, d, generator=g).to(cuda)
    P = ops.PanelFeat(X)
    Y0, Y1 = torch.empty(n, d, device=cuda), torch.empty(n, d, device=cuda)
    ops.proj_fwd_group([(X, W, b, Y0)], d, mode)
    ops.proj_fwd_group([(P, W, b, Y1)], d, mode)
    assert torch.equal(Y0, Y1)
    out = []
    for Xop in (X, P):
        dW, db = torch.empty(d, k, device=cuda), torch.empty(d, device=cuda)
        ops.proj_wgrad_group([(Xop, dY, dW, db, False)], d, mode)
        ops.proj_wgrad_group([(Xop, dY, dW, db, True)], d, mode)
        out.append((dW, db))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    if mode == 0:
        torch.testing.assert_close(Y1.double(), X.double() @ W.double().t() + b.double(), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(out[1][0].double(), 2 * (dY.double().t() @ X.double()), rtol=1e-4, atol=2e-4 * n ** 0.5)


def test_panel_layout_rejected_off_the_tensor_core_path():
    from llmrec_b200 import ops
    X = torch.randn(64, 64, device=cuda)
    P = ops.PanelFeat(X)
    W, Y = torch.randn(32, 64, device=cuda), torch.empty(64, 32, device=cuda)
    with pytest.raises(RuntimeError):
        ops.proj_fwd_group([(P, W, None, Y)], 32, 2)
    with pytest.raises(ValueError):
        ops.PanelFeat(torch.randn(8, 40, device=cuda))


def test_training_with_panel_features_is_identical(tiny_root):
    from test_path_gpu import _trainer
    losses = []
    for layout in ("rows", "panels"):
        tr, gen, M = _trainer(tiny_root, ["--feat_layout", layout, "--cuda_graph", "0"])
        M.set_seed(7)
        ls = []
        for _ in range(3):
            users, pos, neg = tr.sample_batch()
            ls.append(float(tr.train_batch(users, pos, neg)))
        losses.append((ls, tr.model_mm.state_dict()["image_trans.weight"].clone()))
    assert losses[0][0] == losses[1][0]
    assert torch.equal(losses[0][1], losses[1][1])


def test_item_sharded_exchange_world2_equals_engine():
    """dist.ShardedHotPath(item_sharded=True): reduce-scatter / row-local / all-gather form == single-GPU engine."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, LLMREC_DIST_ITEM_SHARDED="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", os.path.join(here, "dist_gpu_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_feature_path_equals_engine(world):
    """dist_feat.ShardedFeatureHotPath (side features, users and item tables sharded) == engine.HotPath."""
    import subprocess
    import sys
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, os.path.join(here, "dist_feat_gpu_check.py")] if world == 1 else \
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
         "--master-port", "29523", os.path.join(here, "dist_feat_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_FEAT_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
