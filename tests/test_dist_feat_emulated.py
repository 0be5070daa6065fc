"""dist_feat.ShardedFeatureHotPath (full side-feature model, users AND item feature tables sharded) on CPU: world-size-1/2/3 gloo
runs with torch stand-ins for the kernels (tests/ops_emulator.py) must track the CPU oracle step for step.  Uneven item ranges
(world 3) take the zero-fill + all-reduce form of the item-row gather, even ones (world 2) the all-gather."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ddir, layers, out):
    sys.path.insert(0, HERE); sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import numpy as np
    import ops_emulator
    ops_emulator.install()
    from llmrec_b200.dist import ShardedGraph, shard_bounds
    from llmrec_b200.dist_feat import ShardedFeatureHotPath
    from llmrec_b200.engine import HotPathConfig, PARAM_ORDER
    from oracle import llmrec_oracle as O
    data = O.load_dataset(ddir)
    d = 32
    ocfg = O.OracleConfig(batch_size=128, embed_size=d, weight_size=(d,) * layers, lr=1e-3)
    O.set_seed(2022)
    otr = O.OracleTrainer(data, ocfg)
    nu, ni = data.n_users, data.n_items
    ub, ib = shard_bounds(nu, world), shard_bounds(ni, world)
    lo, hi, ilo, ihi = ub[rank], ub[rank + 1], ib[rank], ib[rank + 1]
    coo = data.train_mat.tocoo()
    keep = (coo.row >= lo) & (coo.row < hi)
    g = ShardedGraph(torch.from_numpy(coo.row[keep].astype(np.int64) - lo), torch.from_numpy(coo.col[keep].astype(np.int64)), hi - lo, ni)
    params = {k: otr.params[k].detach().clone() for k in PARAM_ORDER}
    params["user_id_embedding.weight"] = params["user_id_embedding.weight"][lo:hi].clone()
    feats = dict(image=otr.feats["image"][ilo:ihi].clone(), text=otr.feats["text"][ilo:ihi].clone(), user=otr.feats["user"][lo:hi].clone(),
                 item={k: v[ilo:ihi].clone() for k, v in otr.feats["item"].items()})
    cfg = HotPathConfig(embed_size=d, n_layers=layers, batch_size=128)
    hp = ShardedFeatureHotPath(g, params, feats, cfg, lo, ilo)
    hp.set_lr(1e-3)
    O.set_seed(7)
    ok = True
    t = lambda x: torch.tensor(x, dtype=torch.int32)
    for step in range(3):
        users, pos, neg = O.sample_batch(data, ocfg)                   # same seed on every rank -> same batch
        got = float(hp.train_step(t(users), t(pos), t(neg)))
        want, _ = otr.step(users, pos, neg)
        ok &= abs(got - want) < 2e-5 * max(1.0, abs(want))
    tol = dict(rtol=2e-4, atol=2e-6)
    for k in PARAM_ORDER:
        want = otr.params[k].detach()
        ok &= bool(torch.allclose(params[k], want[lo:hi] if k == "user_id_embedding.weight" else want, **tol))
    U, I = hp.forward()
    with torch.no_grad():
        o = otr.forward()
    ok &= bool(torch.allclose(U, o["U"][lo:hi], rtol=1e-4, atol=1e-6)) and bool(torch.allclose(I, o["I"], rtol=1e-4, atol=1e-6))
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,layers", [(1, 2), (2, 2), (2, 1), (3, 3)])
def test_sharded_feature_engine_tracks_the_oracle(tiny_root, world, layers):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), os.path.join(tiny_root, "netflix_valid_item"), layers, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}
