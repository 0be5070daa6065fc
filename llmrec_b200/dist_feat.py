"""Multi-GPU form of the FULL hot path (side-feature projections + propagation + fusion + the 8 loss heads)  --  SURVEY.md 8e.

Same sharding as dist.ShardedHotPath (users in contiguous row ranges, item-sized activations replicated), plus:
    item feature tables   sharded by contiguous ITEM ranges: rank r stores and projects rows [ilo, ihi) only (the projection is
                          row-local, Models.py:145-150); the projected block Pi [ni x S*d] is all-gathered;
    user feature table    sharded with the users; P_usr stays local;
    linear layers         replicated; every rank forms the weight gradient of its own rows, ONE small all-reduce per tensor.
Exchanges per step (L layers, S = 2 + #attribute tables feature blocks of d columns):
    forward   all-gather Pi | sum_r R_r^T Fu_r [ni x S*d] | sum_r R_r^T P_usr_r [ni x d] | sum_r R_r^T U_l,r [ni x d] per layer
              | the batch's user rows [B' x 4d]
    backward  sum_r R_r^T (su . Gprof_u_r) [ni x d] | sum_r R_r^T (su . gU_l,r) [ni x d] per layer | sum_r R_r^T (su . GFu_r) [ni x S*d]
              | the 8 weight / bias gradients
Item-side gradients that come straight from the loss (heads, feat_reg, fusion) are identical on every rank and are added ONCE,
after the cross-rank sum.  The schedule mirrors engine.HotPath line by line; world size 1 reproduces it.

Status: the orchestration is verified on CPU (tests/test_dist_feat_emulated.py: world-size-1/2 gloo, torch stand-ins for the
kernels, against the oracle); it has not run on GPUs yet and no benchmark uses it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .dist import ShardedGraph, owner_local_index, shard_bounds
from .engine import HotPathConfig, PARAM_ORDER


class ShardedFeatureHotPath:
    """params: dict name -> fp32 tensor; `user_id_embedding.weight` holds THIS RANK's user rows, everything else is replicated.
    feats_local: dict(image, text, item={key: ...}) with this rank's ITEM rows [item_lo, item_hi) and user = this rank's USER rows."""

    def __init__(self, graph: ShardedGraph, params, feats_local, cfg: HotPathConfig, user_lo: int, item_lo: int, group=None):
        self.g, self.cfg, self.group, self.p, self.f = graph, cfg, group, params, feats_local
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.E_u, self.E_i = params["user_id_embedding.weight"], params["item_id_embedding.weight"]
        nu, ni, d, L = self.E_u.shape[0], self.E_i.shape[0], cfg.embed_size, cfg.n_layers
        self.nu, self.ni, self.d, self.L = nu, ni, d, L
        self.lo, self.hi = int(user_lo), int(user_lo) + nu
        self.keys = list(feats_local["item"].keys())
        S = self.S = 2 + len(self.keys)
        self.ilo, self.ihi = int(item_lo), int(item_lo) + feats_local["image"].shape[0]
        self.even_items = self.world > 1 and ni % self.world == 0 and (self.ihi - self.ilo) * self.world == ni
        dev = self.E_i.device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.Pi, self.Fu, self.Fi = new(ni, S * d), new(nu, S * d), new(ni, S * d)
        self.P_usr, self.prof_i, self.prof_u = new(nu, d), new(ni, d), new(nu, d)
        self.GFu, self.GFi, self.GP_usr = new(nu, S * d), new(ni, S * d), new(nu, d)
        self.Gprof_i, self.Gprof_u = new(ni, d), new(nu, d)
        self.Ul = [self.E_u] + [new(nu, d) for _ in range(L)]
        self.Il = [self.E_i] + [new(ni, d) for _ in range(L)]
        self.U, self.I = new(nu, d), new(ni, d)
        self.part, self.part_w = new(ni, d), new(ni, S * d)
        self.parts = [new(ni, d), new(ni, d)]
        self.gU, self.gI, self.dIl = new(nu, d), new(ni, d), new(ni, d)
        self.bufU, self.tmpI = new(nu, d), new(ni, d)
        self.grads = {k: torch.zeros_like(v) for k, v in params.items()}
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_local = torch.zeros(1, dtype=torch.float32, device=dev)
        self.n_heads = 3 + len(self.keys)
        self.head_out = torch.zeros(self.n_heads * 4, dtype=torch.float32, device=dev)
        self._B = None
        self.opt = ops.AdamW([params[k] for k in PARAM_ORDER], lr=1e-4)
        self.comm_bytes = 0

    def set_lr(self, lr):
        self.opt.lr = lr

    def blk(self, buf, s):
        return buf[:, s * self.d:(s + 1) * self.d]

    # -- collectives ---------------------------------------------------------------------------------------------------------
    def _allreduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
            self.comm_bytes += t.numel() * 4

    def _gather_item_rows(self, full):
        """rows [ilo, ihi) of `full` are valid on this rank -> all rows valid everywhere."""
        if self.world == 1:
            return
        if self.even_items:
            dist.all_gather_into_tensor(full, full[self.ilo:self.ihi], group=self.group)
        else:                                                         # uneven item ranges: zero the foreign rows and sum
            full[:self.ilo].zero_(); full[self.ihi:].zero_()
            dist.all_reduce(full, group=self.group)
        self.comm_bytes += full.numel() * 4

    def _item_sum(self, op, segs_src, buf, out, softmax=False):
        """out = [softmax]( si (.) sum over ranks of op . src ): per-rank partial into `buf`, one all-reduce, row scale."""
        op.apply([(x, y, None, False) for x, y in segs_src])
        self._allreduce(buf)
        ops.row_scale_softmax(buf, self.g.si, out, softmax)

    # -- forward (Models.py:145-197) --------------------------------------------------------------------------------------------
    def forward(self):
        g, p, f, d, S, L, m = self.g, self.p, self.f, self.d, self.S, self.L, self.cfg.proj_mode
        own = self.Pi[self.ilo:self.ihi]
        probs = [(f["image"], p["image_trans.weight"], p["image_trans.bias"], self.blk(own, 0)),
                 (f["text"], p["text_trans.weight"], p["text_trans.bias"], self.blk(own, 1))]
        probs += [(f["item"][k], p["item_trans.weight"], p["item_trans.bias"], self.blk(own, 2 + j)) for j, k in enumerate(self.keys)]
        probs.append((f["user"], p["user_trans.weight"], p["user_trans.bias"], self.P_usr))
        probs = [t for t in probs if t[0].shape[0] > 0]
        probs.sort(key=lambda t: -t[0].shape[1])
        ops.proj_fwd_group(probs, d, m)                                                                   # :145-150, this rank's rows
        self._gather_item_rows(self.Pi)
        g.ui.apply([(self.blk(self.Pi, s), self.blk(self.Fu, s), None, False) for s in range(S)] + [(self.Il[0], self.Ul[1], None, L == 1)])
        self._item_sum(g.iu_raw, [(self.blk(self.Fu, s), self.blk(self.part_w, s)) for s in range(S)], self.part_w, self.Fi)   # :154,157,163
        self._item_sum(g.iu_raw, [(self.P_usr, self.part)], self.part, self.prof_i)                        # :166
        self._item_sum(g.iu_raw, [(self.Ul[1], self.part)], self.part, self.Il[1], L == 1)                 # :175,180
        segs = [(self.prof_i, self.prof_u, None, False)]                                                   # :167
        if L >= 2:
            segs.append((self.Il[1], self.Ul[2], None, L == 2))
        g.ui.apply(segs)
        for l in range(2, L + 1):
            self._item_sum(g.iu_raw, [(self.Ul[l], self.part)], self.part, self.Il[l], l == L)
            if l < L:
                g.ui.apply([(self.Il[l], self.Ul[l + 1], None, l + 1 == L)])
        c = self.cfg
        coefs = [c.model_cat_rate, c.model_cat_rate, c.user_cat_rate] + [c.item_cat_rate] * len(self.keys)
        su = [self.blk(self.Fu, 0), self.blk(self.Fu, 1), self.prof_u] + [self.blk(self.Fu, 2 + j) for j in range(len(self.keys))]
        si = [self.blk(self.Fi, 0), self.blk(self.Fi, 1), self.prof_i] + [self.blk(self.Fi, 2 + j) for j in range(len(self.keys))]
        ops.fuse_fwd(self.Ul, su, coefs, self.U)                                                           # :185-197
        ops.fuse_fwd(self.Il, si, coefs, self.I)
        self._fuse_args = (coefs, su, si)
        return self.U, self.I

    # -- losses + output gradients (main.py:232-256,330-342,151-165) --------------------------------------------------------------------
    def loss_and_output_grads(self, users, pos, neg):
        c, d = self.cfg, self.d
        B = int(users.numel())
        if self._B != B:
            dev = users.device
            self._B = B
            self.Ub4, self.gUb4 = torch.empty(B, 4 * d, device=dev), torch.empty(B, 4 * d, device=dev)
            self.arange = torch.arange(B, dtype=torch.int32, device=dev)
            self.work = ops.bpr_work(self.n_heads, B, dev)
        n_keep = int((1 - c.prune_loss_drop_rate) * B)
        local = owner_local_index(users, self.lo, self.hi)
        ub = lambda s: self.Ub4[:, s * d:(s + 1) * d]
        gb = lambda s: self.gUb4[:, s * d:(s + 1) * d]
        for s, src in enumerate((self.U, self.blk(self.Fu, 0), self.blk(self.Fu, 1), self.prof_u)):       # owners fill, others zero
            ops.gather_rows(src, local, ub(s))
        self._allreduce(self.Ub4)
        for t in (self.loss, self.loss_local, self.gUb4, self.gU, self.gI, self.GFu, self.GFi, self.Gprof_u, self.Gprof_i):
            t.zero_()
        creg = c.feat_reg_decay / self.ni
        d2 = 2 * d
        ops.sqnorm_grad(self.Fu[:, :d2], self.GFu[:, :d2], creg, False, self.loss_local)                   # this rank's users only
        ops.sqnorm_grad(self.Fi[:, :d2], self.GFi[:, :d2], creg, False, self.loss)                         # replicated
        heads = [(ub(0), self.I, gb(0), self.gI, 1.0, 1.0),
                 (ub(1), self.blk(self.Fi, 0), gb(1), self.blk(self.GFi, 0), c.mm_mf_rate, 0.0),
                 (ub(2), self.blk(self.Fi, 1), gb(2), self.blk(self.GFi, 1), c.mm_mf_rate, 0.0)]
        for j in range(len(self.keys)):
            heads.append((ub(3), self.blk(self.Fi, 2 + j), gb(3), self.blk(self.GFi, 2 + j), c.aug_mf_rate, 0.0))
        ops.bpr_heads(heads, self.arange, pos, neg, n_keep, c.regs0 / c.batch_size, self.head_out, self.loss, self.work)
        for s, dst in enumerate((self.gU, self.blk(self.GFu, 0), self.blk(self.GFu, 1), self.Gprof_u)):   # user-row grads to their owners
            ops.scatter_add_rows(gb(s), local, dst)
        self._allreduce(self.loss_local)
        self.loss += self.loss_local
        return self.loss

    # -- backward ---------------------------------------------------------------------------------------------------------------
    def backward(self):
        g, d, S, L, m = self.g, self.d, self.S, self.L, self.cfg.proj_mode
        G, f = self.grads, self.f
        coefs, su, si = self._fuse_args
        nk = len(self.keys)
        dsu = [self.blk(self.GFu, 0), self.blk(self.GFu, 1), self.Gprof_u] + [self.blk(self.GFu, 2 + j) for j in range(nk)]
        dsi = [self.blk(self.GFi, 0), self.blk(self.GFi, 1), self.Gprof_i] + [self.blk(self.GFi, 2 + j) for j in range(nk)]
        dUl = G["user_id_embedding.weight"]
        ops.fuse_bwd(self.gU, L + 1, dUl, su, coefs, dsu, True)
        ops.fuse_bwd(self.gI, L + 1, self.dIl, si, coefs, dsi, True)
        # prof_u = ui . prof_i  ->  Gprof_i += sum_r R_r^T (su . Gprof_u_r)
        g.uiT_raw.apply([(self.Gprof_u, self.part, None, False)])
        self._allreduce(self.part)
        self.Gprof_i += self.part
        g_cur = self.dIl
        for l in range(L, 0, -1):
            src = ops.row_softmax_bwd(self.Il[l], g_cur, out=self.tmpI) if l == L else g_cur
            segs = [(src, self.bufU, dUl, False)]                                                          # gU_l = dUl + iu^T src (local rows)
            if l == L:
                segs += [(self.blk(self.GFi, s), self.blk(self.GFu, s), self.blk(self.GFu, s), False) for s in range(S)]
                segs.append((self.Gprof_i, self.GP_usr, None, False))
            g.iuT.apply(segs)
            if l == L:
                ops.row_softmax_bwd(self.Ul[l], self.bufU, out=self.bufU)
            dst = self.parts[l & 1]
            g.uiT_raw.apply([(self.bufU, dst, None, False)])                                               # sum_r R_r^T (su . gU_l)
            self._allreduce(dst)
            dst += self.dIl                                                                                # replicated direct part, once
            if l == L:
                g.uiT_raw.apply([(self.blk(self.GFu, s), self.blk(self.part_w, s), None, False) for s in range(S)])
                self._allreduce(self.part_w)                                                               # = GPi (all item rows)
            g_cur = dst
        G["item_id_embedding.weight"].copy_(g_cur)
        # weight gradients: this rank's item rows / user rows, then one all-reduce per tensor
        own = self.part_w[self.ilo:self.ihi]
        probs = [(f["item"][k], self.blk(own, 2 + j), G["item_trans.weight"], G["item_trans.bias"], j > 0) for j, k in enumerate(self.keys)]
        probs.append((f["user"], self.GP_usr, G["user_trans.weight"], G["user_trans.bias"], False))
        probs.append((f["text"], self.blk(own, 1), G["text_trans.weight"], G["text_trans.bias"], False))
        probs.append((f["image"], self.blk(own, 0), G["image_trans.weight"], G["image_trans.bias"], False))
        live = [t for t in probs if t[0].shape[0] > 0]
        for name in ("image_trans", "text_trans", "user_trans", "item_trans"):
            if not any(t[2] is G[name + ".weight"] for t in live):                                        # a rank without rows of that table
                G[name + ".weight"].zero_(); G[name + ".bias"].zero_()
        if live:
            ops.proj_wgrad_group(self._fix_accumulate(live), d, m)
        for name in ("image_trans", "text_trans", "user_trans", "item_trans"):
            self._allreduce(G[name + ".weight"])
            self._allreduce(G[name + ".bias"])
        return G

    @staticmethod
    def _fix_accumulate(probs):
        """accumulate flags must be False for the first problem writing a given dW and True for the later ones."""
        seen, out = set(), []
        for X, dY, dW, db, _ in probs:
            out.append((X, dY, dW, db, id(dW) in seen))
            seen.add(id(dW))
        return out

    def train_step(self, users, pos, neg):
        self.forward()
        self.loss_and_output_grads(users, pos, neg)
        self.backward()
        self.opt.step([self.grads[k] for k in PARAM_ORDER])
        return self.loss


def item_shard_bounds(n_items: int, world: int):
    """Item ranges of the feature tables (same contiguous balanced partition as the users')."""
    return shard_bounds(n_items, world)
