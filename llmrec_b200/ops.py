"""Thin tensor-level wrappers over the C ABI (include/llmrec_b200.h).

PyTorch supplies device memory and the current CUDA stream only; every op below is one or more
launches of the hand-written sm_100a kernels.  2-D operands must be row-major views
(stride(1) == 1); column slices of wider buffers are fine (the leading dimension is passed).
"""
from __future__ import annotations

import copy
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N


STATS = {"launches": 0}     # kernels of libllmrec_b200 launched through this module (bench.py reports it)


def _count(n=1):
    STATS["launches"] += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _mat(t: torch.Tensor, name="operand"):
    if t.dim() != 2 or t.dtype != torch.float32 or not t.is_cuda or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name}: need a CUDA fp32 row-major 2-D tensor, got {tuple(t.shape)} {t.dtype} {t.device} strides {t.stride()}")
    return t


def _ld(t):
    return int(t.stride(0)) if t.shape[0] > 1 else int(max(t.stride(0), t.shape[1]))


def _i32(t, name="index"):
    if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: need a contiguous CUDA int32 tensor")
    return t


DEFAULT_TILE_NNZ = int(os.environ.get("LLMREC_SPMM_TILE", "0"))      # 0 = size tiles from the graph


def auto_tile_nnz(nnz):
    """16 non-zeros per tile on small graphs (parallelism and short dependent-load chains; measured best at netflix
    scale: 8 cuts ordinary rows into pieces, 32/64 are 3-6 % slower per step), growing to 248 on large ones (amortised
    index reads, few long-row pieces)."""
    t = nnz // 65536
    return int(min(248, max(16, (t // 8) * 8)))


class TilePlan:
    """nnz-bounded work decomposition of a CSR pattern (llmrec_spmm_plan_tiles); shared by the operators
    that use the same pattern (forward of one direction, backward of the other)."""

    def __init__(self, rowptr_dev, n_rows, tile_nnz=0, max_rows=15):
        rp = np.ascontiguousarray(rowptr_dev.cpu().numpy().astype(np.int32))
        tile_nnz = int(tile_nnz) or DEFAULT_TILE_NNZ or auto_tile_nnz(int(rp[-1]))
        lib = N.lib()
        counts = np.zeros(3, dtype=np.int32)
        N.check(lib.llmrec_spmm_plan_tiles(rp.ctypes.data, n_rows, tile_nnz, max_rows, None, None, None, counts.ctypes.data), "spmm_plan")
        tiles = np.zeros((max(int(counts[0]), 1), 8), dtype=np.int32)
        srow = np.zeros(max(int(counts[1]), 1), dtype=np.int32)
        sfirst = np.zeros(int(counts[1]) + 1, dtype=np.int32)
        N.check(lib.llmrec_spmm_plan_tiles(rp.ctypes.data, n_rows, tile_nnz, max_rows, tiles.ctypes.data, srow.ctypes.data,
                                           sfirst.ctypes.data, counts.ctypes.data), "spmm_plan")
        dev = rowptr_dev.device
        self.tiles, self.split_row, self.split_first = (torch.from_numpy(a).to(dev) for a in (tiles, srow, sfirst))
        self.n_tiles, self.n_split, self.n_split_tiles = (int(c) for c in counts)
        self.tile_nnz, self.scratch = tile_nnz, None
        # per (long row, column window) tickets: the last piece to finish reduces the row inside the same launch (self-resetting)
        self.tickets = torch.zeros(max(self.n_split, 1) * 64, dtype=torch.int32, device=dev) if self.n_split else None


class CsrOperator:
    """One sparse operator  Y = diag(rs) . P(vals) . diag(cs) . X  over a CSR pattern P."""

    def __init__(self, rowptr, col, n_rows, n_cols, vals=None, rs=None, cs=None, tile_nnz=0, plan=None):
        self.rowptr, self.col = _i32(rowptr, "rowptr"), _i32(col, "col")
        self.vals, self.rs, self.cs = vals, rs, cs
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.nnz = int(col.numel())
        self.plan = plan if plan is not None else TilePlan(self.rowptr, self.n_rows, tile_nnz)

    def branch(self):
        """The same operator with its own long-row scratch and tickets, for launches that may overlap launches of `self` (or of an
        operator sharing its plan) on another stream."""
        o, k = copy.copy(self), copy.copy(self.plan)
        k.scratch = None
        k.tickets = torch.zeros_like(self.plan.tickets) if self.plan.tickets is not None else None
        o.plan = k
        return o

    def _tiling_struct(self, width):
        k = self.plan
        need = k.n_split_tiles * width
        if need and (k.scratch is None or k.scratch.numel() < need):
            if k.scratch is not None:
                _retired.append(k.scratch)
            k.scratch = torch.empty(need, dtype=torch.float32, device=self.rowptr.device)
        return N.SpmmTiling(_p(k.tiles), _p(k.split_row), _p(k.split_first), _p(k.scratch), k.n_tiles, k.n_split, k.n_split_tiles, 0, _p(k.tickets), None)

    def apply_rows(self, seg, rows, count, max_rows=None, src_mask=None, cta_per_row=False):
        """Row-list form (llmrec_spmm_rows_f32): only rows[0 .. count[0]) are computed and written.  seg = (X, Y, Z|None, softmax);
        rows: int32 CUDA list, count: int32[1] CUDA (device-side length), src_mask: optional uint32/int32 bitmask over source rows."""
        X, Y, Z, sm = seg
        _mat(X, "spmm X"); _mat(Y, "spmm Y")
        d = int(X.shape[1])
        if X.shape[0] != self.n_cols or Y.shape[0] != self.n_rows or Y.shape[1] != d:
            raise ValueError("spmm_rows: shape mismatch")
        sg = N.SpmmSeg(_p(X), _p(Y), _p(Z) if Z is not None else None, _ld(X), _ld(Y), _ld(Z) if Z is not None else 0, N.SPMM_SOFTMAX if sm else 0, 0)
        mx = int(rows.numel()) if max_rows is None else int(max_rows)
        N.check(N.lib().llmrec_spmm_rows_f32(_p(self.rowptr), _p(self.col), _p(self.vals), _p(self.rs), _p(self.cs), d, C.byref(sg),
                                              _p(_i32(rows)), _p(count), mx, _p(src_mask), 1 if cta_per_row else 0, _stream()), "spmm_rows")
        _count()

    def apply(self, segs, src_mask=None):
        """segs: list of (X, Y, Z_or_None, softmax: bool); all share d = X.shape[1].  src_mask: optional bitmask over the SOURCE rows
        (rows of X): a clear bit promises an all-zero row, whose fetch is skipped."""
        if not segs:
            return
        d = int(segs[0][0].shape[1])
        arr = (N.SpmmSeg * len(segs))()
        for i, (X, Y, Z, sm) in enumerate(segs):
            _mat(X, "spmm X"); _mat(Y, "spmm Y")
            if X.shape[0] != self.n_cols or Y.shape[0] != self.n_rows or X.shape[1] != d or Y.shape[1] != d:
                raise ValueError(f"spmm: shape mismatch X{tuple(X.shape)} Y{tuple(Y.shape)} for operator {self.n_rows}x{self.n_cols}")
            arr[i] = N.SpmmSeg(_p(X), _p(Y), _p(Z) if Z is not None else None, _ld(X), _ld(Y), _ld(Z) if Z is not None else 0,
                               N.SPMM_SOFTMAX if sm else 0, 0)
        til = self._tiling_struct(d * min(len(segs), N.MAX_SEG))
        til.src_mask = src_mask.data_ptr() if src_mask is not None else None
        N.check(N.lib().llmrec_spmm_csr_f32(_p(self.rowptr), _p(self.col), _p(self.vals), _p(self.rs), _p(self.cs),
                                             self.n_rows, self.n_cols, d, arr, len(segs), C.byref(til), _stream()), "spmm")
        _count(-(-len(segs) // N.MAX_SEG))


def row_softmax(X, out=None):
    out = torch.empty_like(X) if out is None else out
    N.check(N.lib().llmrec_row_softmax_f32(_p(_mat(X)), _ld(X), _p(_mat(out)), _ld(out), X.shape[0], X.shape[1], _stream()), "row_softmax")
    _count()
    return out


def row_softmax_bwd(S, dS, out=None):
    out = torch.empty((S.shape[0], S.shape[1]), dtype=torch.float32, device=S.device) if out is None else out
    N.check(N.lib().llmrec_row_softmax_bwd_f32(_p(_mat(S)), _ld(S), _p(_mat(dS)), _ld(dS), _p(_mat(out)), _ld(out),
                                                S.shape[0], S.shape[1], _stream()), "row_softmax_bwd")
    _count()
    return out


def row_softmax_bwd_rows(S, dS, out, rows, count, max_rows=None):
    mx = int(rows.numel()) if max_rows is None else int(max_rows)
    N.check(N.lib().llmrec_row_softmax_bwd_rows_f32(_p(_mat(S)), _ld(S), _p(_mat(dS)), _ld(dS), _p(_mat(out)), _ld(out), _p(_i32(rows)), _p(count), mx,
                                                     S.shape[1], _stream()), "row_softmax_bwd_rows")
    _count()
    return out


class RowSet:
    """A set of row ids living on the device: uint32 bitmask + compacted id list + its length, rebuilt every step without a host sync."""

    def __init__(self, n, device):
        self.n = int(n)
        self.mask = torch.zeros((self.n + 31) // 32 + 1, dtype=torch.int32, device=device)
        self.list = torch.zeros(max(self.n, 1), dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)

    def clear(self):
        N.check(N.lib().llmrec_fill_f32(_p(self.mask), self.mask.numel(), 0.0, _stream()), "fill")
        N.check(N.lib().llmrec_fill_f32(_p(self.count), 1, 0.0, _stream()), "fill")
        _count(2)

    def add_neighbors(self, rowptr, col, rows):
        """every column id of the CSR rows named in `rows` (int32 CUDA list; entries < 0 skipped)"""
        N.check(N.lib().llmrec_mark_neighbors(_p(_i32(rowptr)), _p(_i32(col)), _p(_i32(rows)), rows.numel(), _p(self.mask), _stream()), "mark_neighbors")
        _count()

    def add_ids(self, ids):
        N.check(N.lib().llmrec_mark_ids(_p(_i32(ids)), ids.numel(), _p(self.mask), _stream()), "mark_ids")
        _count()

    def compact(self):
        N.check(N.lib().llmrec_compact_mask(_p(self.mask), self.n, _p(self.list), _p(self.count), _stream()), "compact_mask")
        _count()


def zero_rows(Y, idx):
    N.check(N.lib().llmrec_zero_rows_f32(_p(_mat(Y)), _ld(Y), _p(_i32(idx)), idx.numel(), Y.shape[1], _stream()), "zero_rows")
    _count()


def assign_rows(G, idx, Y):
    N.check(N.lib().llmrec_assign_rows_f32(_p(_mat(G)), _ld(G), _p(_i32(idx)), idx.numel(), G.shape[1], _p(_mat(Y)), _ld(Y), _stream()), "assign_rows")
    _count()


PROJ_MODE = {"3xtf32": 0, "tf32": 1, "fp32": 2}


_scratch = {}


_retired = []     # outgrown scratch buffers stay allocated: a captured CUDA graph may still hold their addresses


def _get_scratch(key, n, device, zero=False):
    t = _scratch.get(key)
    if t is None or t.numel() < n:
        if t is not None:
            _retired.append(t)
        t = _scratch[key] = (torch.zeros if zero else torch.empty)(max(int(n), 1), dtype=torch.float32, device=device)
    return t


def proj_fwd_group(problems, d, mode=0):
    """problems: list of (X[n x k], W[d x k], bias[d]|None, out[n x d]).  One grouped launch (tcgen05) --
    the 8 nn.Linear calls of Models.py:145-150.  Problems sharing W share the hi/lo split buffer."""
    arr = (N.ProjFwdProblem * len(problems))()
    for i, (X, W, b, out) in enumerate(problems):
        _mat(out)
        n, k = _mat(X).shape
        if not W.is_contiguous() or tuple(W.shape) != (d, k) or tuple(out.shape) != (n, d):
            raise ValueError("proj_fwd: bad shapes")
        ws = _get_scratch(("wsplit", W.data_ptr()), 2 * d * k, X.device) if mode == 0 else None
        arr[i] = N.ProjFwdProblem(_p(X), _p(W), _p(b), _p(out), _p(ws), _ld(X), _ld(out), n, k, 0)
    N.check(N.lib().llmrec_proj_fwd_group_f32(arr, len(problems), d, mode, _stream()), "proj_fwd_group")
    _count((2 if mode == 0 else 1) * -(-len(problems) // 8))


def proj_fwd(X, W, b, out, mode=0):
    """out[n x d] = X[n x k] W[d x k]^T + b  (nn.Linear; Models.py:145-150)."""
    proj_fwd_group([(X, W, b, out)], W.shape[0], mode)
    return out


def proj_wgrad_group(problems, d, mode=0):
    """problems: list of (X[n x k], dY[n x d], dW[d x k], db[d]|None, accumulate).  dW (+)= dY^T X ; db (+)= colsum(dY)."""
    arr = (N.ProjWgradProblem * len(problems))()
    for i, (X, dY, dW, db, acc) in enumerate(problems):
        _mat(dY)
        n, k = _mat(X).shape
        if not dW.is_contiguous() or tuple(dW.shape) != (d, k) or tuple(dY.shape) != (n, d):
            raise ValueError("proj_wgrad: bad shapes")
        arr[i] = N.ProjWgradProblem(_p(X), _p(dY), _p(dW), _p(db), _ld(X), _ld(dY), n, k, 1 if acc else 0)
    need = int(N.lib().llmrec_proj_wgrad_group_scratch(arr, len(problems), d, mode))
    scratch = _get_scratch(("wgrad", problems[0][0].device.index), need, problems[0][0].device, zero=True) if need else None     # holds a ticket word
    N.check(N.lib().llmrec_proj_wgrad_group_f32(arr, len(problems), d, mode, _p(scratch), need, _stream()), "proj_wgrad_group")
    _count(3 if need else len(problems))


def proj_wgrad(X, dY, dW, db, accumulate=False, mode=0):
    """dW[d x k] (+)= dY^T X ; db[d] (+)= colsum(dY)."""
    proj_wgrad_group([(X, dY, dW, db, accumulate)], dY.shape[1], mode)


def _ptr_table(tensors):
    arr = (C.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def _ld_table(tensors):
    arr = (C.c_int64 * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = _ld(t) if t is not None else 0
    return arr


def fuse_fwd(layers, sides, coefs, out, rows=None, compact=False):
    """out = mean(layers) + sum_t coefs[t] * normalize(sides[t])   (Models.py:185-197).
    rows: only these rows; compact=True: layers are read at rows[b], sides and out are [len(rows) x d] blocks indexed by b."""
    for t in list(layers) + list(sides) + [out]:
        _mat(t)
    n = out.shape[0] if rows is None else rows.numel()
    if compact:
        if rows is None:
            raise ValueError("fuse_fwd: compact form needs a row list")
        n = -n
    cf = (C.c_float * max(1, len(coefs)))(*[float(c) for c in coefs])
    N.check(N.lib().llmrec_fuse_fwd_f32(_ptr_table(layers), _ld_table(layers), len(layers), _ptr_table(sides), _ld_table(sides), cf,
                                         len(sides), _p(out), _ld(out), _p(rows), n, out.shape[1], _stream()), "fuse_fwd")
    _count()
    return out


def fuse_bwd(g, n_layers, d_layer, sides, coefs, d_sides, accumulate, rows=None):
    _mat(g)
    n = g.shape[0] if rows is None else rows.numel()
    cf = (C.c_float * max(1, len(coefs)))(*[float(c) for c in coefs])
    N.check(N.lib().llmrec_fuse_bwd_f32(_p(g), _ld(g), n_layers, _p(d_layer), _ld(d_layer) if d_layer is not None else 0,
                                         _ptr_table(sides), _ld_table(sides), cf, _ptr_table(d_sides), _ld_table(d_sides), len(sides),
                                         1 if accumulate else 0, _p(rows), n, g.shape[1], _stream()), "fuse_bwd")
    _count()


def bpr_work(n_heads, B, device):
    return torch.zeros(int(N.lib().llmrec_bpr_work_elems(n_heads, B)), dtype=torch.float32, device=device)


def bpr_heads(heads, users, pos, neg, n_keep, regs0_over_bs, out, loss, work, meta=None):
    """heads: list of (XU, XI, GU|None, GI|None, w_mf, w_emb).  See include/llmrec_b200.h.
    meta: optional int32 CUDA tensor {live B', n_keep}; users/pos/neg/work are then sized for the capacity B."""
    arr = (N.BprHead * len(heads))()
    d = int(heads[0][0].shape[1])
    for i, (XU, XI, GU, GI, wmf, wemb) in enumerate(heads):
        _mat(XU); _mat(XI)
        arr[i] = N.BprHead(_p(XU), _p(XI), _p(GU), _p(GI), _ld(XU), _ld(XI), _ld(GU) if GU is not None else 0,
                           _ld(GI) if GI is not None else 0, float(wmf), float(wemb))
    B = int(users.numel())
    if work.numel() < N.lib().llmrec_bpr_work_elems(len(heads), B):
        raise ValueError("bpr_heads: work buffer too small for this capacity")
    N.check(N.lib().llmrec_bpr_heads_f32(arr, len(heads), _p(_i32(users)), _p(_i32(pos)), _p(_i32(neg)), B, int(n_keep), _p(meta),
                                          float(regs0_over_bs), d, _p(out), _p(loss), _p(work), _stream()), "bpr_heads")
    _count(2)


_ginit = {}


def grad_init(regions, loss):
    """regions: list of (G, X|None, c): G = c*X (or 0) written once; loss = sum 0.5*c*|X|^2 (overwritten).  One launch."""
    arr = (N.GradRegion * len(regions))()
    for i, (G, X, c) in enumerate(regions):
        _mat(G)
        if X is not None and tuple(_mat(X).shape) != tuple(G.shape):
            raise ValueError("grad_init: X and G shapes differ")
        arr[i] = N.GradRegion(_p(G), _p(X), _ld(G), _ld(X) if X is not None else 0, G.shape[0], G.shape[1], float(c))
    key = loss.device.index
    sc = _ginit.get(key)
    if sc is None:
        sc = _ginit[key] = torch.zeros(int(N.lib().llmrec_grad_init_scratch()), dtype=torch.float32, device=loss.device)
    N.check(N.lib().llmrec_grad_init_f32(arr, len(regions), _p(loss), _p(sc), _stream()), "grad_init")
    _count()


_partial = {}


def sqnorm_grad(X, G, c, accumulate, loss):
    """loss += c*0.5*sum(X^2); G = (G if accumulate else 0) + c*X   (feat_reg, main.py:151-156)."""
    _mat(X)
    part = _partial.get(X.device.index)
    if part is None:
        part = _partial[X.device.index] = torch.empty(1024, dtype=torch.float32, device=X.device)
    N.check(N.lib().llmrec_sqnorm_grad_f32(_p(X), _ld(X), _p(G), _ld(G) if G is not None else 0, X.shape[0], X.shape[1], float(c),
                                            1 if accumulate else 0, _p(loss), _p(part), _stream()), "sqnorm_grad")
    _count(2)


class AdamW:
    """Dense fused AdamW over a fixed list of parameter tensors (torch.optim.AdamW defaults)."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        dev = self.params[0].device
        self.m = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.v = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.state = torch.zeros(4, dtype=torch.float64, device=dev)
        self._n = (C.c_int64 * len(self.params))(*[p.numel() for p in self.params])

    def advance(self):
        """bump the device-side step counter / bias corrections once per optimizer step (before any step_tensor)"""
        N.check(N.lib().llmrec_adamw_advance(_p(self.state), self.lr, self.betas[0], self.betas[1], _stream()), "adamw_advance")
        _count()

    def step_tensor(self, i, grad, row_mask=None):
        """update parameter i alone (after advance()): lets independent tables be updated at different points of a step, e.g. the
        user table while an item-side exchange is in flight"""
        lib, p = N.lib(), self.params[i]
        if row_mask is not None:
            N.check(lib.llmrec_adamw_step_rows_f32(_p(p), _p(grad), _p(self.m[i]), _p(self.v[i]), p.shape[0], p.shape[1], _p(row_mask), _p(self.state),
                                                    self.lr, self.betas[0], self.betas[1], self.eps, self.wd, _stream()), "adamw_step_rows")
        else:
            n = (C.c_int64 * 1)(p.numel())
            N.check(lib.llmrec_adamw_step_f32(_ptr_table([p.data]), _ptr_table([grad]), _ptr_table([self.m[i]]), _ptr_table([self.v[i]]), n, 1, _p(self.state),
                                               self.lr, self.betas[0], self.betas[1], self.eps, self.wd, _stream()), "adamw_step")
        _count()

    def step(self, grads, row_masks=None):
        """row_masks: optional list (one entry per parameter) of RowSet-style bitmasks or None: a masked [n x w] parameter reads its
        gradient only on the flagged rows and takes the g = 0 update elsewhere (row-sparse gradients of a dense AdamW)."""
        lib = N.lib()
        N.check(lib.llmrec_adamw_advance(_p(self.state), self.lr, self.betas[0], self.betas[1], _stream()), "adamw_advance")
        dense = [i for i in range(len(self.params)) if not (row_masks and row_masks[i] is not None)]
        for i in range(len(self.params)):
            if i in dense:
                continue
            p = self.params[i]
            N.check(lib.llmrec_adamw_step_rows_f32(_p(p), _p(grads[i]), _p(self.m[i]), _p(self.v[i]), p.shape[0], p.shape[1], _p(row_masks[i]), _p(self.state),
                                                    self.lr, self.betas[0], self.betas[1], self.eps, self.wd, _stream()), "adamw_step_rows")
            _count()
        if dense:
            ps, gs, ms, vs = ([x[i] for i in dense] for x in ([p.data for p in self.params], grads, self.m, self.v))
            n = (C.c_int64 * len(dense))(*[self.params[i].numel() for i in dense])
            N.check(lib.llmrec_adamw_step_f32(_ptr_table(ps), _ptr_table(gs), _ptr_table(ms), _ptr_table(vs), n, len(dense), _p(self.state),
                                               self.lr, self.betas[0], self.betas[1], self.eps, self.wd, _stream()), "adamw_step")
        _count(1 + -(-len(dense) // 16))


SCORE_MODE = {"3xtf32": 0, "fp32": 2}
_score_scratch = {}


def score_topk(U, I, users, mask_rowptr, mask_col, K, mode=0, want_vals=False):
    """Top-K item ids per user among items not in the user's mask row; ties -> lowest id."""
    _mat(U); _mat(I)
    nb, ni, d = int(users.numel()), int(I.shape[0]), int(I.shape[1])
    idx = torch.empty((nb, K), dtype=torch.int32, device=U.device)
    val = torch.empty((nb, K), dtype=torch.float32, device=U.device) if want_vals else None
    need = int(N.lib().llmrec_score_topk_scratch(nb, ni, d, K, mode))
    key = U.device.index
    scratch = _score_scratch.get(key)
    if need and (scratch is None or scratch.numel() < need):
        if scratch is not None:
            _retired.append(scratch)
        scratch = _score_scratch[key] = torch.empty(need, dtype=torch.float32, device=U.device)
    N.check(N.lib().llmrec_score_topk_f32(_p(U), _ld(U), _p(I), _ld(I), _p(_i32(users)), nb, ni, d, _p(mask_rowptr), _p(mask_col), K,
                                           _p(idx), _p(val), mode, _p(scratch), scratch.numel() if scratch is not None else 0,
                                           _stream()), "score_topk")
    _count(3)
    return (idx, val) if want_vals else idx


def topk_hits(idx, users, truth_rowptr, truth_col):
    hits = torch.empty(idx.shape, dtype=torch.uint8, device=idx.device)
    N.check(N.lib().llmrec_topk_hits(_p(_i32(idx)), idx.shape[0], idx.shape[1], _p(_i32(users)), _p(_i32(truth_rowptr)), _p(_i32(truth_col)),
                                      _p(hits), _stream()), "topk_hits")
    _count()
    return hits


def user_auc(U, I, users, mask_rowptr, mask_col, truth_rowptr, truth_col):
    """fp32[n] per-user ROC-AUC over the candidates (test_flag='full', batch_test.py:38-68)."""
    _mat(U); _mat(I)
    out = torch.empty(users.numel(), dtype=torch.float32, device=U.device)
    N.check(N.lib().llmrec_user_auc_f32(_p(U), _ld(U), _p(I), _ld(I), _p(_i32(users)), users.numel(), I.shape[0], I.shape[1], _p(mask_rowptr), _p(mask_col),
                                         _p(_i32(truth_rowptr)), _p(_i32(truth_col)), _p(out), _stream()), "user_auc")
    _count()
    return out


def row_scale_softmax(X, scale, out, softmax):
    N.check(N.lib().llmrec_row_scale_softmax_f32(_p(_mat(X)), _ld(X), _p(scale), _p(_mat(out)), _ld(out), X.shape[0], X.shape[1],
                                                  1 if softmax else 0, _stream()), "row_scale_softmax")
    _count()
    return out


def gather_rows(X, idx, out):
    N.check(N.lib().llmrec_gather_rows_f32(_p(_mat(X)), _ld(X), _p(_i32(idx)), idx.numel(), X.shape[1], _p(_mat(out)), _ld(out), _stream()), "gather_rows")
    _count()
    return out


def scatter_add_rows(G, idx, Y):
    N.check(N.lib().llmrec_scatter_add_rows_f32(_p(_mat(G)), _ld(G), _p(_i32(idx)), idx.numel(), G.shape[1], _p(_mat(Y)), _ld(Y), _stream()), "scatter_add_rows")
    _count()


def _scale_col(t):
    """1-D fp32 CUDA view (a column of a row-major table) -> (pointer, element stride)."""
    if t is None:
        return C.c_void_p(0), 0
    if t.dim() != 1 or t.dtype != torch.float32 or not t.is_cuda:
        raise ValueError("scale: need a 1-D fp32 CUDA tensor (a column view is fine)")
    return C.c_void_p(t.data_ptr()), int(t.stride(0)) if t.numel() > 1 else 1


def rank1_add(blocks):
    """blocks: list of (Y[n x w], scale[n] (1-D view), bias[w]):  Y += scale (x) bias.  One launch."""
    arr = (N.Rank1Block * len(blocks))()
    for i, (Y, sc, b) in enumerate(blocks):
        _mat(Y)
        ptr, lds = _scale_col(sc)
        arr[i] = N.Rank1Block(_p(Y), ptr, _p(b), _ld(Y), lds, Y.shape[0], Y.shape[1], 0)
    N.check(N.lib().llmrec_rank1_add_f32(arr, len(blocks), _stream()), "rank1_add")
    _count()


def scaled_colsum(terms, out, accumulate=False):
    """out[c] (+)= sum over terms (G[n x w], scale[n] | None) of sum_r scale[r] * G[r, c].  One launch, deterministic."""
    w = int(out.numel())
    arr = (N.ColsumTerm * len(terms))()
    for i, (G, sc) in enumerate(terms):
        _mat(G)
        if G.shape[1] != w:
            raise ValueError("scaled_colsum: width mismatch")
        ptr, lds = _scale_col(sc)
        arr[i] = N.ColsumTerm(_p(G), ptr, _ld(G), lds, G.shape[0])
    key = ("colsum", out.device.index, w)
    sc_ = _scratch.get(key)
    if sc_ is None:
        sc_ = _scratch[key] = torch.zeros(int(N.lib().llmrec_scaled_colsum_scratch(w)), dtype=torch.float32, device=out.device)
    N.check(N.lib().llmrec_scaled_colsum_f32(arr, len(terms), w, _p(out), 1 if accumulate else 0, _p(sc_), _stream()), "scaled_colsum")
    _count()


def feat_reg_gram(W, b, G, h, n2, c, dW, db, loss):
    """feat_reg over all rows through the Gram matrix G[k x k] of a propagated table (include/llmrec_b200.h)."""
    d, k = W.shape
    key = ("gram", W.device.index, d, k)
    sc_ = _scratch.get(key)
    if sc_ is None:
        sc_ = _scratch[key] = torch.zeros(int(N.lib().llmrec_feat_reg_gram_scratch(d, k)), dtype=torch.float32, device=W.device)
    N.check(N.lib().llmrec_feat_reg_gram_f32(_p(W), _p(b), _p(G), _p(h), float(n2), d, k, float(c), _p(dW), _p(db), _p(loss), _p(sc_), _stream()), "feat_reg_gram")
    _count(2)


def fill(t, v):
    N.check(N.lib().llmrec_fill_f32(_p(t), t.numel(), float(v), _stream()), "fill")
