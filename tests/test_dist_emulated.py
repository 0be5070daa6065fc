"""The sharded engine's ORCHESTRATION on CPU: dist.ShardedHotPath (both exchange forms) runs world-size-1 and -2 under gloo
with torch stand-ins for the CUDA kernels (tests/ops_emulator.py) and must reproduce an independent autograd + torch.optim.AdamW
implementation of the ID-only model (Models.py:152-186 without side features; main.py:330-342,158-165 loss).
What this covers is everything between the kernels: shard construction, exchanges, row-sparse item gradients, buffer
ping-pong, optimizer sharding.  The kernels themselves are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
NU, NI, D, L, B, STEPS, LR = 203, 96, 32, 2, 64, 3, 1e-2


def _problem():
    rng = np.random.default_rng(3)
    e = np.unique(np.stack([np.concatenate([np.arange(NU), rng.integers(0, NU, 900)]), rng.integers(0, NI, NU + 900)], 1), axis=0)
    torch.manual_seed(0)
    Eu, Ei = torch.randn(NU, D) * 0.3, torch.randn(NI, D) * 0.3
    batches = [tuple(torch.from_numpy(rng.integers(0, hi, B).astype(np.int32)) for hi in (NU, NI, NI)) for _ in range(STEPS)]
    return e, Eu, Ei, batches


def _reference(e, Eu, Ei, batches, cfg):
    """Plain autograd model + torch.optim.AdamW: no shared code with the engines."""
    R = torch.zeros(NU, NI)
    R[e[:, 0], e[:, 1]] = 1.0
    ui = torch.pow(R.sum(1) + 1e-8, -0.5)[:, None] * R
    iu = torch.pow(R.sum(0) + 1e-8, -0.5)[:, None] * R.t()
    Eu, Ei = Eu.clone().requires_grad_(True), Ei.clone().requires_grad_(True)
    opt = torch.optim.AdamW([Eu, Ei], lr=LR)
    losses = []

    def forward():
        Ul, Il = [Eu], [Ei]
        for l in range(1, L + 1):
            u = ui @ Il[-1]
            u = torch.softmax(u, -1) if l == L else u
            i = iu @ u
            i = torch.softmax(i, -1) if l == L else i
            Ul.append(u); Il.append(i)
        return sum(Ul) / (L + 1), sum(Il) / (L + 1)

    for users, pos, neg in batches:
        U, I = forward()
        a, b, c = U[users.long()], I[pos.long()], I[neg.long()]
        maxi = torch.nn.functional.logsigmoid((a * b).sum(1) - (a * c).sum(1) + 1e-8)
        keep = torch.argsort(maxi.detach(), stable=True)[:int((1 - cfg.prune_loss_drop_rate) * B)]
        loss = -maxi[keep].mean() + cfg.regs0 / cfg.batch_size * (1 / (2 * a.pow(2).sum() + 1e-8) + 1 / (2 * b.pow(2).sum() + 1e-8) + 1 / (2 * c.pow(2).sum() + 1e-8))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    with torch.no_grad():
        U, I = forward()
    return Eu.detach(), Ei.detach(), U, I, losses


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, item_sharded, pieces, out, demand=False):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import ops_emulator
    ops_emulator.install()
    from llmrec_b200.dist import ShardedGraph, ShardedHotPath, shard_bounds
    from llmrec_b200.engine import HotPathConfig
    e, Eu, Ei, batches = _problem()
    cfg = HotPathConfig(embed_size=D, n_layers=L, batch_size=B)
    want_Eu, want_Ei, want_U, want_I, want_losses = _reference(e, Eu, Ei, batches, cfg)
    b = shard_bounds(NU, world)
    lo, hi = b[rank], b[rank + 1]
    mine = e[(e[:, 0] >= lo) & (e[:, 0] < hi)]
    g = ShardedGraph(torch.from_numpy(mine[:, 0] - lo), torch.from_numpy(mine[:, 1]), hi - lo, NI, pieces=pieces)
    sh = ShardedHotPath(g, Eu[lo:hi].clone(), Ei.clone(), cfg, lo, item_sharded=item_sharded, demand=demand)
    assert sh.item_sharded == (item_sharded and world > 1) and sh.demand == demand
    if demand:                                           # rows the demand mode never writes must not leak into results: poison them
        for t in sh.Ul[1:] + sh.Il[1:]:
            t.fill_(123.0)
        sh.g_Eu.fill_(float("nan")); sh.bufU.fill_(float("nan")); sh.tmpI.fill_(float("nan"))
    sh.set_lr(LR)
    tol = dict(rtol=2e-4, atol=2e-6)
    ok = True
    for (users, pos, neg), want in zip(batches, want_losses):
        got = float(sh.train_step(users, pos, neg))
        ok &= abs(got - want) < 1e-5 * max(1.0, abs(want))
    ok &= bool(torch.allclose(sh.E_u, want_Eu[lo:hi], **tol)) and bool(torch.allclose(sh.E_i, want_Ei, **tol))
    U, I = sh.forward()
    ok &= bool(torch.allclose(U, want_U[lo:hi], **tol)) and bool(torch.allclose(I, want_I, **tol))
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,item_sharded,pieces,demand", [(1, False, 1, False), (2, False, 1, False), (2, False, 2, False), (2, True, 1, False),
                                                               (1, False, 1, True), (2, False, 1, True), (2, False, 2, True)])
def test_sharded_engine_orchestration(world, item_sharded, pieces, demand):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), item_sharded, pieces, out, demand), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}
