"""Host-side logic of the sharded path on CPU: partition helpers, and a world_size-2 gloo run of the shard
construction (row shards tile the graph, item degrees are globally reduced, every batch row has exactly one owner)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llmrec_b200.dist import build_shard_csr, owner_local_index, shard_bounds


def test_shard_bounds_and_ownership():
    for n, w in ((10, 3), (1000, 8), (7, 8), (13187, 2)):
        b = shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and all(0 <= b[i + 1] - b[i] <= n // w + 1 for i in range(w))
    users = torch.tensor([0, 5, 9, 3, 5], dtype=torch.int32)
    b = shard_bounds(10, 3)
    loc = [owner_local_index(users, b[r], b[r + 1]) for r in range(3)]
    owned = torch.stack([(l >= 0) for l in loc]).sum(0)
    assert owned.tolist() == [1] * 5
    for r in range(3):
        m = loc[r] >= 0
        assert (loc[r][m].long() + b[r] == users[m].long()).all()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    rng = np.random.default_rng(0)
    nu, ni, ne = 211, 97, 1500
    e = np.unique(np.stack([rng.integers(0, nu, ne), rng.integers(0, ni, ne)], 1), axis=0)
    b = shard_bounds(nu, world)
    mine = e[(e[:, 0] >= b[rank]) & (e[:, 0] < b[rank + 1])]
    c = build_shard_csr(torch.from_numpy(mine[:, 0] - b[rank]), torch.from_numpy(mine[:, 1]), b[rank + 1] - b[rank], ni)
    deg_i = np.bincount(e[:, 1], minlength=ni)
    deg_u = np.bincount(e[:, 0], minlength=nu)[b[rank]:b[rank + 1]]
    ok = np.allclose(c["si"].numpy(), np.power(deg_i + 1e-8, -0.5).astype(np.float32))          # global item degrees
    ok &= np.allclose(c["su"].numpy(), np.power(deg_u + 1e-8, -0.5).astype(np.float32))         # local user degrees
    rows = np.repeat(np.arange(b[rank + 1] - b[rank]), np.diff(c["rowptr_u"].numpy()))
    ok &= bool((np.stack([rows + b[rank], c["col_u"].numpy()], 1) == mine[np.lexsort((mine[:, 1], mine[:, 0]))]).all())
    rows_t = np.repeat(np.arange(ni), np.diff(c["rowptr_i"].numpy()))
    tr = np.stack([c["col_i"].numpy() + b[rank], rows_t], 1)
    ok &= bool((tr[np.lexsort((tr[:, 1], tr[:, 0]))] == mine[np.lexsort((mine[:, 1], mine[:, 0]))]).all())
    # partial item-side products sum to the full product (the exchange step of dist.ShardedHotPath)
    torch.manual_seed(1)
    Y = torch.randn(nu, 8)
    R = torch.zeros(nu, ni); R[e[:, 0], e[:, 1]] = 1.0
    part = torch.zeros(ni, 8)
    part.index_add_(0, torch.from_numpy(mine[:, 1]), Y[torch.from_numpy(mine[:, 0])])
    dist.all_reduce(part)
    ok &= bool(torch.allclose(part, R.t() @ Y, atol=1e-5))
    nnz = torch.tensor([c["nnz"]]); dist.all_reduce(nnz)
    ok &= int(nnz) == len(e)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gloo_two_rank_shards():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
