"""-m gpu: every kernel behind the C ABI against a plain torch reference of the same op."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu

cuda = torch.device("cuda")


def _rand_csr(n_rows, n_cols, nnz, seed, power=False):
    rng = np.random.default_rng(seed)
    if power:
        w = 1.0 / (np.arange(n_rows) + 1.0)
        rows = rng.choice(n_rows, size=nnz, p=w / w.sum())
    else:
        rows = rng.integers(0, n_rows, nnz)
    cols = rng.integers(0, n_cols, nnz)
    m = sp.csr_matrix((np.ones(nnz, np.float32), (rows, cols)), shape=(n_rows, n_cols))
    m.sum_duplicates(); m.data[:] = 1.0; m.sort_indices()
    return m


def _op(m, vals=False, rs=False, cs=False, tile=0, seed=0):
    from llmrec_b200.ops import CsrOperator
    g = torch.Generator().manual_seed(seed)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(cuda)
    v = torch.rand(m.nnz, generator=g).to(cuda) if vals else None
    r = torch.rand(m.shape[0], generator=g).to(cuda) if rs else None
    c = torch.rand(m.shape[1], generator=g).to(cuda) if cs else None
    op = CsrOperator(t(m.indptr, np.int32), t(m.indices, np.int32), m.shape[0], m.shape[1], vals=v, rs=r, cs=c, tile_nnz=tile)
    dense = torch.from_numpy(m.toarray()).double()
    if vals:
        d2 = torch.zeros_like(dense)
        rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
        d2[rows, m.indices] = v.cpu().double()
        dense = d2
    if rs:
        dense = r.cpu().double()[:, None] * dense
    if cs:
        dense = dense * c.cpu().double()[None, :]
    return op, dense


@pytest.mark.parametrize("d,nseg", [(64, 1), (64, 8), (128, 1), (128, 9), (32, 1), (32, 2), (16, 1), (100, 3), (6, 2), (256, 5)])
@pytest.mark.parametrize("variant", ["rs", "cs", "vals+rs+cs", "plain"])
def test_spmm_matches_dense(d, nseg, variant):
    m = _rand_csr(301, 257, 2000, seed=d + nseg)
    op, dense = _op(m, vals="vals" in variant, rs="rs" in variant, cs="cs" in variant)
    g = torch.Generator().manual_seed(1)
    wide = torch.randn(257, nseg * d, generator=g).to(cuda)          # column-block views of one buffer
    out = torch.full((301, nseg * d), float("nan"), device=cuda)
    segs = [(wide[:, s * d:(s + 1) * d], out[:, s * d:(s + 1) * d], None, False) for s in range(nseg)]
    op.apply(segs)
    ref = (dense @ wide.cpu().double())
    torch.testing.assert_close(out.cpu().double(), ref, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("d", [32, 64, 128, 48])
def test_spmm_softmax_and_addend(d):
    m = _rand_csr(200, 180, 1500, seed=3)
    op, dense = _op(m, rs=True)
    g = torch.Generator().manual_seed(2)
    X1, X2 = torch.randn(180, d, generator=g).to(cuda), torch.randn(180, d, generator=g).to(cuda)
    Z = torch.randn(200, d, generator=g).to(cuda)
    Y1, Y2 = torch.empty(200, d, device=cuda), torch.empty(200, d, device=cuda)
    op.apply([(X1, Y1, None, True), (X2, Y2, Z, False)])
    r1 = torch.softmax(dense @ X1.cpu().double(), dim=-1)
    r2 = dense @ X2.cpu().double() + Z.cpu().double()
    torch.testing.assert_close(Y1.cpu().double(), r1, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(Y2.cpu().double(), r2, rtol=2e-6, atol=2e-6)
    # in-place accumulate (Z is Y)
    Y3 = Z.clone()
    op.apply([(X2, Y3, Y3, False)])
    torch.testing.assert_close(Y3.cpu().double(), r2, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("d,nseg", [(64, 1), (128, 2), (64, 7)])
def test_spmm_row_tiling_long_rows(d, nseg):
    m = _rand_csr(150, 4000, 30000, seed=5, power=True)       # a few rows with thousands of entries
    assert np.diff(m.indptr).max() > 1000
    op, dense = _op(m, cs=True, rs=True, tile=248)
    assert op.plan.n_split > 0
    g = torch.Generator().manual_seed(4)
    X = [torch.randn(4000, d, generator=g).to(cuda) for _ in range(nseg)]
    Y = [torch.empty(150, d, device=cuda) for _ in range(nseg)]
    op.apply([(x, y, None, s == 0) for s, (x, y) in enumerate(zip(X, Y))])
    for s in range(nseg):
        ref = dense @ X[s].cpu().double()
        if s == 0:
            ref = torch.softmax(ref, -1)
        torch.testing.assert_close(Y[s].cpu().double(), ref, rtol=1e-4, atol=1e-4)   # sums of thousands of terms
    # deterministic: bitwise equal across runs
    Y2 = [torch.empty(150, d, device=cuda) for _ in range(nseg)]
    op.apply([(x, y, None, s == 0) for s, (x, y) in enumerate(zip(X, Y2))])
    assert all(torch.equal(a, b) for a, b in zip(Y, Y2))


def test_spmm_empty_rows_and_empty_matrix():
    from llmrec_b200.ops import CsrOperator
    m = sp.csr_matrix((5, 7), dtype=np.float32)
    op, _ = _op(m)
    X = torch.randn(7, 64, device=cuda); Y = torch.full((5, 64), 3.0, device=cuda)
    op.apply([(X, Y, None, False)])
    assert torch.count_nonzero(Y) == 0


@pytest.mark.parametrize("d", [64, 128, 20])
def test_softmax_and_backward(d):
    from llmrec_b200 import ops
    X = torch.randn(333, d, device=cuda)
    S = ops.row_softmax(X)
    torch.testing.assert_close(S, torch.softmax(X, -1), rtol=1e-5, atol=1e-7)
    g = torch.randn(333, d, device=cuda)
    Xr = X.clone().requires_grad_(True)
    torch.softmax(Xr, -1).backward(g)
    torch.testing.assert_close(ops.row_softmax_bwd(S, g), Xr.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("n,k,d", [(1000, 512, 64), (777, 96, 64), (300, 1536, 128), (65, 33, 24)])
def test_projection_and_wgrad(n, k, d):
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(0)
    X = torch.randn(n, k, generator=g).to(cuda)
    W = (torch.randn(d, k, generator=g) / k ** 0.5).to(cuda)
    b = torch.randn(d, generator=g).to(cuda)
    wide = torch.empty(n, 3 * d, device=cuda)
    Y = wide[:, d:2 * d]
    for mode in (0, 2):
        ops.proj_fwd(X, W, b, Y, mode)
        ref = (X.double() @ W.double().t() + b.double())
        # mode 2: sequential fp32 FMA; mode 0: 3xTF32 products (2^-21) + tensor-core fp32 accumulation -> 1e-5-class
        tol = 1e-5 if mode == 2 else 1e-4
        torch.testing.assert_close(Y.double(), ref, rtol=tol, atol=tol)
        dY = torch.randn(n, 3 * d, generator=g).to(cuda)[:, d:2 * d]
        dW = torch.empty(d, k, device=cuda); db = torch.empty(d, device=cuda)
        ops.proj_wgrad(X, dY, dW, db, False, mode)
        torch.testing.assert_close(dW.double(), dY.double().t() @ X.double(), rtol=1e-4, atol=1e-4 * (n ** 0.5))
        torch.testing.assert_close(db.double(), dY.double().sum(0), rtol=1e-4, atol=1e-4)
        ops.proj_wgrad(X, dY, dW, db, True, mode)
        torch.testing.assert_close(dW.double(), 2 * (dY.double().t() @ X.double()), rtol=1e-4, atol=2e-4 * (n ** 0.5))


@pytest.mark.parametrize("d,L,T", [(64, 3, 8), (128, 4, 8), (32, 2, 0), (20, 3, 3)])
def test_fuse_forward_backward(d, L, T):
    from llmrec_b200 import ops
    n = 257
    g = torch.Generator().manual_seed(d)
    layers = [torch.randn(n, d, generator=g).to(cuda).requires_grad_(True) for _ in range(L)]
    wide = torch.randn(n, max(T, 1) * d, generator=g).to(cuda)
    wide[5] = 0.0                                               # zero row -> clamped-norm branch
    sides = [wide[:, t * d:(t + 1) * d].detach().clone().requires_grad_(True) for t in range(T)]
    sides_v = [wide[:, t * d:(t + 1) * d] for t in range(T)]
    coefs = [0.02, 0.02, 2.8, 0.005, 0.005, 0.005, 0.005, 0.005][:T]
    ref = torch.mean(torch.stack(layers), 0)
    for c, s in zip(coefs, sides):
        ref = ref + c * torch.nn.functional.normalize(s, p=2, dim=1)
    out = torch.empty(n, d, device=cuda)
    ops.fuse_fwd([l.detach() for l in layers], sides_v, coefs, out)
    torch.testing.assert_close(out, ref.detach(), rtol=1e-5, atol=1e-6)
    go = torch.randn(n, d, generator=g).to(cuda)
    ref.backward(go)
    dl = torch.empty(n, d, device=cuda)
    dwide = torch.ones(n, max(T, 1) * d, device=cuda)
    ds = [dwide[:, t * d:(t + 1) * d] for t in range(T)]
    ops.fuse_bwd(go, L, dl, sides_v, coefs, ds, True)
    torch.testing.assert_close(dl, layers[0].grad, rtol=1e-5, atol=1e-7)
    for t in range(T):
        want = sides[t].grad + 1.0
        torch.testing.assert_close(ds[t], want, rtol=1e-4, atol=1e-5)
    # row-list variant touches only the listed rows
    rows = torch.tensor([3, 5, 100], dtype=torch.int32, device=cuda)
    out2 = torch.zeros(n, d, device=cuda)
    ops.fuse_fwd([l.detach() for l in layers], sides_v, coefs, out2, rows=rows)
    torch.testing.assert_close(out2[rows.long()], ref.detach()[rows.long()], rtol=1e-5, atol=1e-6)
    assert torch.count_nonzero(out2) == torch.count_nonzero(out2[rows.long()])


@pytest.mark.parametrize("B,d", [(1126, 64), (140, 64), (2000, 128), (7, 20)])
def test_bpr_heads_vs_oracle(B, d):
    from llmrec_b200 import ops
    from oracle import llmrec_oracle as O
    nu, ni = 500, 700
    g = torch.Generator().manual_seed(B)
    cfg = O.OracleConfig(batch_size=128)
    XU = [torch.randn(nu, d, generator=g) * 0.3 for _ in range(2)]
    XI = [torch.randn(ni, d, generator=g) * 0.3 for _ in range(3)]
    users = torch.randint(0, nu, (B,), generator=g); pos = torch.randint(0, ni, (B,), generator=g); neg = torch.randint(0, ni, (B,), generator=g)
    spec = [(0, 0, 1.0, 1.0), (1, 1, 1e-4, 0.0), (1, 2, 0.012, 0.0)]       # (user matrix, item matrix, w_mf, w_emb)
    # oracle
    xu = [x.clone().requires_grad_(True) for x in XU]; xi = [x.clone().requires_grad_(True) for x in XI]
    tot, vals = 0, []
    for a, b, wm, we in spec:
        mf, emb = O.bpr_head(xu[a][users], xi[b][pos], xi[b][neg], cfg)
        tot = tot + wm * mf + we * emb
        vals.append((float(mf), float(emb)))
    tot.backward()
    # device
    dXU = [x.to(cuda) for x in XU]; dXI = [x.to(cuda) for x in XI]
    GU = [torch.zeros_like(x) for x in dXU]; GI = [torch.zeros_like(x) for x in dXI]
    heads = [(dXU[a], dXI[b], GU[a], GI[b], wm, we) for a, b, wm, we in spec]
    out = torch.zeros(len(spec) * 4, device=cuda); loss = torch.zeros(1, device=cuda)
    work = ops.bpr_work(len(spec), B, cuda)
    i32 = lambda t: t.to(torch.int32).to(cuda)
    n_keep = O.num_remember(B, cfg.prune_loss_drop_rate)
    for _ in range(2):        # second call checks the work/counter state is re-usable
        for t in GU + GI:
            t.zero_()
        loss.zero_()
        ops.bpr_heads(heads, i32(users), i32(pos), i32(neg), n_keep, cfg.regs0 / cfg.batch_size, out, loss, work)
    o = out.cpu().view(-1, 4)
    for h, (mf, emb) in enumerate(vals):
        assert abs(float(o[h, 0]) - mf) <= 2e-6 * max(1, abs(mf)), (h, float(o[h, 0]), mf)
        assert abs(float(o[h, 1]) - emb) <= 1e-5 * abs(emb) + 1e-12
    assert abs(float(loss) - float(tot)) <= 3e-6 * max(1.0, abs(float(tot)))
    for a in range(2):
        torch.testing.assert_close(GU[a].cpu(), xu[a].grad, rtol=2e-4, atol=1e-8)
    for b in range(3):
        torch.testing.assert_close(GI[b].cpu(), xi[b].grad, rtol=2e-4, atol=1e-8)


def test_bpr_prune_ties_keep_lowest_positions():
    from llmrec_b200 import ops
    d, B = 8, 10
    XU = torch.zeros(4, d); XU[:, 0] = 1.0
    XI = torch.zeros(6, d); XI[0, 0] = 1.0; XI[1, 0] = 3.0          # scores: item0 ->1, item1 ->3
    users = torch.zeros(B, dtype=torch.int32)
    pos = torch.tensor([0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=torch.int32)    # positions 2,6 have larger x
    neg = torch.full((B,), 2, dtype=torch.int32)
    GU, GI = torch.zeros(4, d, device=cuda), torch.zeros(6, d, device=cuda)
    out = torch.zeros(4, device=cuda); loss = torch.zeros(1, device=cuda)
    ops.bpr_heads([(XU.to(cuda), XI.to(cuda), GU, GI, 1.0, 0.0)], users.to(cuda), pos.to(cuda), neg.to(cuda), 3, 0.0, out, loss, ops.bpr_work(1, B, cuda))
    # the 3 smallest logsigmoid values are the x=1 entries at the LOWEST positions 0,1,3 -> item0 grad only
    ls1 = torch.nn.functional.logsigmoid(torch.tensor(1.0 + 1e-8))
    assert abs(float(out[0]) + float(ls1)) < 1e-6
    assert float(GI[1].abs().sum()) == 0.0 and float(GI[0, 0]) < 0


def test_sqnorm_grad():
    from llmrec_b200 import ops
    X = torch.randn(1000, 192, device=cuda)[:, 32:160]
    G = torch.ones(1000, 128, device=cuda)
    loss = torch.full((1,), 2.0, device=cuda)
    ops.sqnorm_grad(X, G, 3e-3, False, loss)
    torch.testing.assert_close(G, 3e-3 * X, rtol=1e-6, atol=1e-9)
    assert abs(float(loss) - (2.0 + 3e-3 * 0.5 * float((X.double() ** 2).sum()))) < 1e-3
    ops.sqnorm_grad(X, G, 3e-3, True, loss)
    torch.testing.assert_close(G, 6e-3 * X, rtol=1e-6, atol=1e-9)


def test_adamw_matches_torch():
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 512), (64,), (1000, 64), (13, 7)]
    ps = [torch.randn(*s, generator=g).to(cuda) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    topt = torch.optim.AdamW(ref, lr=1e-3)
    mine = ops.AdamW(ps, lr=1e-3)
    for step in range(5):
        grads = [torch.randn(*s, generator=g).to(cuda) for s in shapes]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        topt.step()
        mine.step(grads)
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p, r.detach(), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("nb,ni,d,K", [(300, 1000, 64, 50), (64, 17366, 64, 50), (33, 5000, 128, 20), (5, 70, 24, 64)])
def test_score_topk_vs_oracle(nb, ni, d, K):
    from llmrec_b200 import ops
    from oracle import llmrec_oracle as O
    g = torch.Generator().manual_seed(ni)
    nu = nb + 10
    U = torch.randn(nu, d, generator=g); I = torch.randn(ni, d, generator=g)
    users = torch.randperm(nu, generator=g)[:nb]
    rng = np.random.default_rng(0)
    train = [sorted(rng.choice(ni, size=rng.integers(0, 8), replace=False).tolist()) for _ in range(nu)]
    rowptr = np.zeros(nu + 1, np.int32); rowptr[1:] = np.cumsum([len(t) for t in train])
    col = np.array([c for t in train for c in t], dtype=np.int32)
    for mode in (0, 2):
        idx, val = ops.score_topk(U.to(cuda), I.to(cuda), users.to(torch.int32).to(cuda), torch.from_numpy(rowptr).to(cuda),
                                  torch.from_numpy(col).to(cuda), K, mode=mode, want_vals=True)
        scores = (U[users].double() @ I.double().t()).float().numpy()
        want = O.rank_users_numpy(scores, [train[u] for u in users.tolist()], K)
        got = idx.cpu().numpy().copy()
        assert not any(set(got[r]) & set(train[u]) for r, u in enumerate(users.tolist()))     # train items never ranked
        for r, u in enumerate(users.tolist()):                 # fewer than K candidates: tail is -1 (the reference returns a shorter list)
            n_cand = ni - len(train[u])
            if n_cand < K:
                assert (got[r, n_cand:] == -1).all()
                got[r, n_cand:] = want[r, n_cand:]
        same = (got == want).all(axis=1)
        # rows may differ only where fp32 summation order flips a near-tie: require the score gap to be at rounding level
        for r in np.nonzero(~same)[0]:
            a, b = scores[r, got[r]], scores[r, want[r]]
            assert np.allclose(a, b, rtol=0, atol=2e-5), (r, got[r], want[r])
        assert same.mean() > 0.97


def test_score_topk_exact_ties_lowest_id():
    from llmrec_b200 import ops
    ni, d, K = 600, 64, 50
    U = torch.zeros(3, d); U[:, 0] = 1.0
    I = torch.zeros(ni, d); I[:, 0] = torch.tensor([float(i % 7) for i in range(ni)])    # 7 score levels, massive ties
    rowptr = torch.tensor([0, 2, 2, 3], dtype=torch.int32); col = torch.tensor([6, 13, 20], dtype=torch.int32)
    for mode in (0, 2):
        idx = ops.score_topk(U.to(cuda), I.to(cuda), torch.arange(3, dtype=torch.int32, device=cuda), rowptr.to(cuda), col.to(cuda), K, mode=mode).cpu()
        for u, banned in ((0, {6, 13}), (1, set()), (2, {20})):
            cand = [i for i in range(ni) if i not in banned]
            want = sorted(cand, key=lambda i: (-(i % 7), i))[:K]
            assert idx[u].tolist() == want


def test_topk_hits():
    from llmrec_b200 import ops
    idx = torch.tensor([[5, 3, 9], [1, 2, -1]], dtype=torch.int32, device=cuda)
    users = torch.tensor([2, 0], dtype=torch.int32, device=cuda)
    rowptr = torch.tensor([0, 1, 1, 3], dtype=torch.int32, device=cuda); col = torch.tensor([2, 9, 5], dtype=torch.int32, device=cuda)
    assert ops.topk_hits(idx, users, rowptr, col).cpu().tolist() == [[1, 0, 1], [0, 1, 0]]


@pytest.mark.parametrize("B,n_keep", [(1, 0), (1, 1), (7, 7), (1126, 326), (5000, 1450), (40000, 11600)])
def test_bpr_select_large_batches_with_ties_and_device_meta(B, n_keep):
    """The per-head radix select against torch's stable argsort: heavy ties (scores drawn from 5 values), the kept set must be the
    n_keep smallest log-sigmoids with ties -> lower position; the same call again with capacity > B and {B, n_keep} read from
    the device (`meta`, the CUDA-graph path) must give the same kept set and loss."""
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(B)
    d, nu, ni = 16, 50, 60
    XU = torch.zeros(nu, d); XU[:, 0] = 1.0
    XI = torch.zeros(ni, d); XI[:, 0] = torch.randint(-2, 3, (ni,), generator=g).float()      # 5 distinct scores -> massive ties
    users = torch.randint(0, nu, (B,), generator=g, dtype=torch.int32)
    pos = torch.randint(0, ni, (B,), generator=g, dtype=torch.int32)
    neg = torch.randint(0, ni, (B,), generator=g, dtype=torch.int32)
    x = XI[pos.long(), 0] - XI[neg.long(), 0] + 1e-8
    maxi = torch.nn.functional.logsigmoid(x)
    keep = torch.argsort(maxi, stable=True)[:n_keep]
    want = -(maxi[keep].double().mean()) if n_keep else float("nan")
    for cap in (B, B + 37):
        pad = lambda t: torch.cat([t, torch.zeros(cap - B, dtype=torch.int32)]).to(cuda)
        meta = torch.tensor([B, n_keep], dtype=torch.int32, device=cuda) if cap != B else None
        GU, GI = torch.zeros(nu, d, device=cuda), torch.zeros(ni, d, device=cuda)
        out = torch.zeros(4, device=cuda); loss = torch.zeros(1, device=cuda)
        work = ops.bpr_work(1, cap, cuda)
        for _ in range(2):
            ops.bpr_heads([(XU.to(cuda), XI.to(cuda), GU, GI, 1.0, 0.0)], pad(users), pad(pos), pad(neg), n_keep, 0.0, out, loss, work, meta=meta)
        torch.cuda.synchronize()
        kept = work[32 + 6 * cap: 32 + 6 * cap + B].cpu()                 # keep flags of head 0 (work layout in csrc/bpr.cu)
        ref = torch.zeros(B); ref[keep] = 1.0
        assert torch.equal(kept, ref), (cap, int((kept != ref).sum()))
        if n_keep:
            assert abs(float(out[0]) - float(want)) <= 2e-6 * max(1.0, abs(float(want)))
        else:
            assert bool(torch.isnan(out[0]))


def test_grad_init_regions():
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(0)
    wide = torch.randn(777, 7 * 64, generator=g).to(cuda)
    G = torch.full((777, 7 * 64), 5.0, device=cuda)
    odd = torch.full((50, 13), 3.0, device=cuda)                            # width % 4 != 0: scalar path of the same kernel
    oddX = torch.randn(50, 13, generator=g).to(cuda)
    loss = torch.full((1,), 123.0, device=cuda)
    c = 2.5e-3
    for _ in range(2):                                                       # second call: the ticket was left at zero
        ops.grad_init([(G[:, :128], wide[:, :128], c), (G[:, 128:], None, 0.0), (odd, oddX, 0.5)], loss)
    torch.testing.assert_close(G[:, :128], c * wide[:, :128], rtol=1e-6, atol=1e-9)
    assert float(G[:, 128:].abs().sum()) == 0.0
    torch.testing.assert_close(odd, 0.5 * oddX, rtol=1e-6, atol=1e-9)
    want = c * 0.5 * float((wide[:, :128].double() ** 2).sum()) + 0.5 * 0.5 * float((oddX.double() ** 2).sum())
    assert abs(float(loss) - want) <= 1e-5 * want                            # overwritten, not accumulated


def test_hoist_helpers_rank1_colsum_gram():
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(1)
    n, d, k = 700, 64, 96
    T = torch.randn(n, k + 32, generator=g).to(cuda)                       # table with a scale column at k
    Y = torch.randn(n, 3 * d, generator=g).to(cuda)
    b = torch.randn(d, generator=g).to(cuda)
    want = Y.clone(); want[:, d:2 * d] += T[:, k][:, None] * b[None, :]
    ops.rank1_add([(Y[:, d:2 * d], T[:, k], b)])
    torch.testing.assert_close(Y, want, rtol=1e-6, atol=1e-6)
    G1, G2 = torch.randn(n, d, generator=g).to(cuda), torch.randn(2 * n, 2 * d, generator=g).to(cuda)
    s2 = torch.randn(2 * n, generator=g).to(cuda)
    out = torch.full((d,), 7.0, device=cuda)
    for _ in range(2):                                                       # ticket reusable
        ops.scaled_colsum([(G1, T[:, k]), (G2[:, d:], s2), (G1, None)], out, accumulate=False)
    ref = (G1.double() * T[:, k].double()[:, None]).sum(0) + (G2[:, d:].double() * s2.double()[:, None]).sum(0) + G1.double().sum(0)
    torch.testing.assert_close(out.double(), ref, rtol=1e-5, atol=1e-4)
    ops.scaled_colsum([(G1, None)], out, accumulate=True)
    torch.testing.assert_close(out.double(), ref + G1.double().sum(0), rtol=1e-5, atol=1e-4)
    # feat_reg through the Gram matrix == the dense definition
    X = torch.randn(900, k, generator=g).to(cuda); s = torch.rand(900, generator=g).to(cuda)
    W = (torch.randn(d, k, generator=g) / k ** 0.5).to(cuda).requires_grad_(True)
    bb = torch.randn(d, generator=g).to(cuda).requires_grad_(True)
    c = 3e-4
    F = X.double() @ W.double().t() + s.double()[:, None] * bb.double()[None, :]
    L = c * 0.5 * (F ** 2).sum()
    L.backward()
    Gm = (X.double().t() @ X.double()).float(); h = (X.double().t() @ s.double()).float(); n2 = float((s.double() ** 2).sum())
    dW = torch.ones(d, k, device=cuda); db = torch.ones(d, device=cuda); loss = torch.full((1,), 2.0, device=cuda)
    for _ in range(1):
        ops.feat_reg_gram(W.detach(), bb.detach(), Gm, h, n2, c, dW, db, loss)
    torch.testing.assert_close(dW.double() - 1.0, W.grad.double(), rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(db.double() - 1.0, bb.grad.double(), rtol=2e-4, atol=1e-6)
    assert abs(float(loss) - 2.0 - float(L)) <= 2e-4 * float(L) + 1e-6


def test_fuse_fwd_compact_rows():
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(2)
    n, d, B = 500, 64, 77
    layers = [torch.randn(n, d, generator=g).to(cuda) for _ in range(3)]
    sides = [torch.randn(B, d, generator=g).to(cuda) for _ in range(2)]
    rows = torch.randint(0, n, (B,), generator=g, dtype=torch.int32).to(cuda)
    out = torch.empty(B, d, device=cuda)
    ops.fuse_fwd(layers, sides, [0.02, 2.8], out, rows=rows, compact=True)
    ref = sum(l[rows.long()] for l in layers) / 3 + 0.02 * torch.nn.functional.normalize(sides[0]) + 2.8 * torch.nn.functional.normalize(sides[1])
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("d", [32, 64, 128])
def test_spmm_row_list_and_source_mask_match_the_full_product(d):
    """llmrec_spmm_rows_f32 and the src_mask of llmrec_spmm_csr_f32 (the demand-driven step of dist.py): listed rows equal the
    full product, unlisted rows are untouched, masked-out source rows are never read (they hold NaN), softmax / addend fused."""
    import scipy.sparse as sp
    from llmrec_b200 import ops
    from llmrec_b200.graph import BipartiteGraph
    rng = np.random.default_rng(d)
    nu, ni, ne = 4001, 1500, 60000
    w = 1.0 / (np.arange(ni) + 4.0); w /= w.sum()
    R = sp.csr_matrix((np.ones(ne, np.float32), (rng.integers(0, nu, ne), rng.choice(ni, size=ne, p=w))), shape=(nu, ni))
    R.sum_duplicates(); R.data[:] = 1.0
    g = BipartiteGraph(R, cuda, tile_nnz=32)
    gen = torch.Generator().manual_seed(d)
    X = torch.randn(ni, d, generator=gen).to(cuda)
    Z = torch.randn(nu, d, generator=gen).to(cuda)
    full = torch.empty(nu, d, device=cuda)
    for sm in (False, True):
        g.ui.apply([(X, full, Z, sm)])
        rs = ops.RowSet(nu, cuda)
        items = torch.tensor(rng.integers(0, ni, 40).astype(np.int32)).to(cuda)
        rs.clear(); rs.add_neighbors(g.rowptr_i, g.col_i, items); rs.add_ids(torch.tensor([5, -1, 7], dtype=torch.int32, device=cuda)); rs.compact()
        n = int(rs.count[0])
        want = set(np.concatenate([R[:, items.cpu().numpy()].nonzero()[0], [5, 7]]).tolist())
        assert n == len(want) and set(rs.list[:n].cpu().tolist()) == want
        out = torch.full((nu, d), 777.0, device=cuda)
        g.ui.apply_rows((X, out, Z, sm), rs.list, rs.count)
        listed = torch.zeros(nu, dtype=torch.bool, device=cuda); listed[rs.list[:n].long()] = True
        torch.testing.assert_close(out[listed], full[listed], rtol=1e-5, atol=1e-6)
        assert bool((out[~listed] == 777.0).all())
    # one CTA per listed row (short list, long rows: the hub items of a batch), softmax + addend fused, duplicates in the list
    Xu = torch.randn(nu, d, generator=gen).to(cuda); Zi = torch.randn(ni, d, generator=gen).to(cuda)
    hubs = torch.tensor([0, 1, 2, 0, 777, -1, 1499], dtype=torch.int32, device=cuda)            # items 0.. are the most popular ones
    cnt = torch.tensor([hubs.numel()], dtype=torch.int32, device=cuda)
    for sm in (False, True):
        fi = torch.empty(ni, d, device=cuda); g.iu.apply([(Xu, fi, Zi, sm)])
        oi = torch.full((ni, d), 555.0, device=cuda)
        g.iu.apply_rows((Xu, oi, Zi, sm), hubs, cnt, cta_per_row=True)
        sel = hubs[hubs >= 0].long()
        torch.testing.assert_close(oi[sel], fi[sel], rtol=2e-5, atol=2e-6)
        rest = torch.ones(ni, dtype=torch.bool, device=cuda); rest[sel] = False
        assert bool((oi[rest] == 555.0).all())
    # source mask: rows of the operand outside the set hold NaN and must never be fetched
    Y = torch.randn(nu, d, generator=gen).to(cuda)
    keep = torch.zeros(nu, dtype=torch.bool, device=cuda); keep[rs.list[:n].long()] = True
    Yz = torch.where(keep[:, None], Y, torch.zeros_like(Y))
    Yn = torch.where(keep[:, None], Y, torch.full_like(Y, float("nan")))
    ref, got = torch.empty(ni, d, device=cuda), torch.empty(ni, d, device=cuda)
    g.uiT.apply([(Yz, ref, None, False)])
    g.uiT.apply([(Yn, got, None, False)], src_mask=rs.mask)
    assert bool(torch.isfinite(got).all())
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    # row-list softmax backward, zero / assign rows
    S = torch.softmax(torch.randn(nu, d, generator=gen).to(cuda), -1); dS = torch.randn(nu, d, generator=gen).to(cuda)
    o = torch.full((nu, d), 3.0, device=cuda)
    ops.row_softmax_bwd_rows(S, dS, o, rs.list, rs.count)
    torch.testing.assert_close(o[keep], (S * (dS - (S * dS).sum(-1, keepdim=True)))[keep], rtol=1e-5, atol=1e-6)
    assert bool((o[~keep] == 3.0).all())
    idx = torch.tensor([3, -1, 9, 3], dtype=torch.int32, device=cuda)
    ops.zero_rows(o, idx)
    assert float(o[3].abs().sum()) == 0.0 and float(o[9].abs().sum()) == 0.0 and float(o[4].abs().sum()) > 0
    G = torch.arange(4 * d, dtype=torch.float32, device=cuda).view(4, d); G[3] = G[0]
    ops.assign_rows(G, idx, o)
    assert torch.equal(o[3], G[0]) and torch.equal(o[9], G[2])


def test_adamw_row_sparse_gradient_matches_dense():
    from llmrec_b200 import ops
    g = torch.Generator().manual_seed(3)
    n, w = 1000, 64
    p0 = torch.randn(n, w, generator=g).to(cuda)
    small = torch.randn(17, 8, generator=g).to(cuda)
    rows = torch.tensor([0, 5, 999, 31, 32, 63, 64], dtype=torch.int32, device=cuda)
    rs = ops.RowSet(n, cuda)
    a, b = ops.AdamW([p0.clone(), small.clone()], lr=1e-3), ops.AdamW([p0.clone(), small.clone()], lr=1e-3)
    for step in range(4):
        gd = torch.zeros(n, w, device=cuda); gd[rows.long()] = torch.randn(rows.numel(), w, generator=g).to(cuda)
        gs = torch.randn(17, 8, generator=g).to(cuda)
        gsparse = torch.where((gd != 0).any(1, keepdim=True), gd, torch.full_like(gd, float("nan")))   # rows without a gradient are never read
        rs.clear(); rs.add_ids(rows)
        a.step([gd, gs])
        b.step([gsparse, gs], row_masks=[rs.mask, None])
    for x, y in zip(a.params + a.m + a.v, b.params + b.m + b.v):
        assert torch.equal(x, y)
