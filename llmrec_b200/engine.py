"""The fused hot path: forward, losses, backward and AdamW of one LLMRec training step,
scheduled by hand over the sm_100a kernels (no autograd, no host synchronisation).

Follows (reference file:line):  MM_Model.forward  Models.py:127-199 ; bpr_loss / prune_loss /
feat_reg / loss assembly  main.py:330-342,158-165,151-156,273 ; AdamW  main.py:100-104,278.
Default flags (mask off, dropout p = 0); the mask / MAE branch drives the same pieces eagerly from main.Trainer._train_batch_masked.

HBM layout (fp32, row-major):
  Pi  [ni x S*d]  side-feature projections, column blocks  img | txt | att_0..att_4      (S = 2 + #keys)
  Fu  [nu x S*d]  ui . Pi      (img_u | txt_u | att_u)        Fi [ni x S*d]  iu . Fu  (img_i | txt_i | att_i)
  P_usr [nu x d], prof_i [ni x d] = iu . P_usr, prof_u [nu x d] = ui . prof_i
  Ul[l] [nu x d], Il[l] [ni x d]  ID layers (l = 0 is the embedding table itself)
  U [nu x d], I [ni x d]          fused outputs
Every operand that shares a sparsity pattern rides in ONE SpMM launch (segments), so the reference's
20 forward SpMMs are 2L launches (4 at L = 2) and the 20 backward ones 2L + 1.

Schedule: a step is a small DAG, not a chain.  `_fork` / `_join` put independent launches on side streams (event fork / join):
the ID layers beside the projection kernel and the side-feature products, the first touch of the gradient buffers beside the
tail of the forward pass, user- and item-side fusion side by side, the ID backward chain beside the side-feature chain and the
weight-gradient kernel.  Captured, that is ONE CUDA graph with parallel branches (0.84 -> 0.74 ms per step at the netflix shape);
on CPU stand-ins and under the span timer the same launches run in program order.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass

import torch

from . import ops


@dataclass
class HotPathConfig:
    embed_size: int = 64
    n_layers: int = 2                 # len(weight_size)
    model_cat_rate: float = 0.02
    user_cat_rate: float = 2.8
    item_cat_rate: float = 0.005
    aug_mf_rate: float = 0.012
    mm_mf_rate: float = 1e-4
    prune_loss_drop_rate: float = 0.71
    feat_reg_decay: float = 1e-5
    regs0: float = 1e-5
    batch_size: int = 1024
    aug_sample_rate: float = 0.1      # main.py:218: a batch grows by at most int(batch_size * rate) augmented triplets
    proj_mode: int = 0                # ops.PROJ_MODE


PARAM_ORDER = ("image_trans.weight", "image_trans.bias", "text_trans.weight", "text_trans.bias",
               "user_trans.weight", "user_trans.bias", "item_trans.weight", "item_trans.bias",
               "user_id_embedding.weight", "item_id_embedding.weight")


class KernelTimer:
    """CUDA-event timer per kernel family on the current stream (bench.py's roofline leg)."""

    def __init__(self):
        self.spans = []

    @contextlib.contextmanager
    def span(self, name):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.spans.append((name, a, b))

    def totals(self):
        """name -> (total ms, count) ; call after a synchronize."""
        out = {}
        for name, a, b in self.spans:
            t, c = out.get(name, (0.0, 0))
            out[name] = (t + a.elapsed_time(b), c + 1)
        return out


class HotPath:
    """params: dict name -> fp32 CUDA tensor (the live parameter storage, updated in place).
    feats: None (ID-only, the large synthetic config) or dict(image, text, user, item={key: tensor})."""

    def __init__(self, operators, params, feats, cfg: HotPathConfig):
        self.ui, self.iu, self.uiT, self.iuT = operators
        self.cfg = cfg
        self.p = params
        self.feats = feats
        d, L = cfg.embed_size, cfg.n_layers
        self.E_u, self.E_i = params["user_id_embedding.weight"], params["item_id_embedding.weight"]
        nu, ni = self.E_u.shape[0], self.E_i.shape[0]
        self.nu, self.ni, self.d, self.L = nu, ni, d, L
        dev = self.E_u.device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.has_feats = feats is not None
        self.keys = list(feats["item"].keys()) if self.has_feats else []
        S = self.S = (2 + len(self.keys)) if self.has_feats else 0
        self.fx = feats                                                    # what the projection kernels read
        if self.has_feats:
            self.Pi, self.Fu, self.Fi = new(ni, S * d), new(nu, S * d), new(ni, S * d)
            self.P_usr, self.prof_i, self.prof_u = new(nu, d), new(ni, d), new(nu, d)
            self.GPi, self.GFu, self.GFi = new(ni, S * d), new(nu, S * d), new(ni, S * d)
            self.GP_usr, self.Gprof_i, self.Gprof_u = new(nu, d), new(ni, d), new(nu, d)
            self.names = ["img", "txt"] + ["att:" + k for k in self.keys]
        self.Ul = [self.E_u] + [new(nu, d) for _ in range(L)]
        self.Il = [self.E_i] + [new(ni, d) for _ in range(L)]
        self.U, self.I = new(nu, d), new(ni, d)
        # gradients
        self.gU, self.gI = new(nu, d), new(ni, d)
        self.dIl = new(ni, d)
        self.bufU, self.bufI, self.tmpI = new(nu, d), new(ni, d), new(ni, d)
        self.grads = {k: torch.zeros_like(v) for k, v in params.items()}
        self.dUl = self.grads["user_id_embedding.weight"]
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.n_heads = (3 + len(self.keys)) if self.has_feats else 1
        self.head_out = torch.zeros(self.n_heads * 4, dtype=torch.float32, device=dev)
        self._bpr_work, self._cap, self._graph = None, 0, None
        self.pre_step = None              # optional launches replayed in front of every staged step (device-side batch sampler)
        self.pre_step_undo = None         # undoes the side effect of ONE pre_step (the warm-up step before a capture must not consume a batch)
        self.opt = None
        self.timer = None
        # independent launches of a step run as BRANCHES: a side stream forked from / joined into the current one with events, so the
        # captured step is a graph with parallel nodes (LLMREC_BRANCHES=0: one chain; off on CPU stand-ins and under the span timer)
        self.branches = dev.type == "cuda" and os.environ.get("LLMREC_BRANCHES", "1") != "0"
        self._side, self._forked = {}, set()
        self.force_split = False          # tests: run the branch SCHEDULE (split launches) even where nothing can overlap (CPU stand-ins)

    def _t(self, name):
        return self.timer.span(name) if self.timer is not None else contextlib.nullcontext()

    # ---- branches -------------------------------------------------------------------------------
    def _fork(self, thunk, lane=0):
        """Run `thunk` on side stream `lane`, ordered after everything enqueued on the current stream so far; `_join` orders the current
        stream behind every lane used since the last join.  Every buffer a branch touches is a persistent engine buffer, so no allocator
        bookkeeping is needed."""
        if not self.branches or self.timer is not None:
            thunk()
            return
        st = self._side.get(lane)
        if st is None:
            st = self._side[lane] = torch.cuda.Stream(device=self.E_u.device)
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            thunk()
        self._forked.add(lane)

    def _opset(self, k):
        """(ui, iu, uiT, iuT) with their own long-row scratch / tickets: launches through set k may overlap launches of the primary set
        and of the other sets."""
        sets = self.__dict__.setdefault("_opsets", {})
        if k not in sets:
            sets[k] = tuple(o.branch() for o in (self.ui, self.iu, self.uiT, self.iuT))
        return sets[k]

    def _join(self, lanes=None):
        for lane in sorted(self._forked if lanes is None else self._forked & set(lanes)):
            torch.cuda.current_stream().wait_stream(self._side[lane])
            self._forked.discard(lane)

    # ---- column-block views ------------------------------------------------------------------
    def blk(self, buf, s):
        d = self.d
        return buf[:, s * d:(s + 1) * d]

    def side_views(self):
        """name -> tensor views in the order of the reference's return tuple (Models.py:199)."""
        if not self.has_feats:
            return {}
        v = dict(img_i=self.blk(self.Fi, 0), txt_i=self.blk(self.Fi, 1), img_u=self.blk(self.Fu, 0), txt_u=self.blk(self.Fu, 1),
                 p_usr=self.P_usr, prof_u=self.prof_u, prof_i=self.prof_i,
                 att_i={k: self.blk(self.Fi, 2 + j) for j, k in enumerate(self.keys)},
                 att_u={k: self.blk(self.Fu, 2 + j) for j, k in enumerate(self.keys)})
        return v

    # ---- forward -------------------------------------------------------------------------------
    def forward(self):
        self._proj_fwd()
        self._prop_fwd()
        self._fuse_fwd()
        return self.U, self.I

    def _proj_fwd(self):
        d, m = self.d, self.cfg.proj_mode
        p, f = self.p, self.fx
        if self.has_feats:
            with self._t("proj_fwd"):                                                                                # Models.py:145-150
                probs = [(f["image"], p["image_trans.weight"], p["image_trans.bias"], self.blk(self.Pi, 0)),
                         (f["text"], p["text_trans.weight"], p["text_trans.bias"], self.blk(self.Pi, 1))]
                probs += [(f["item"][k], p["item_trans.weight"], p["item_trans.bias"], self.blk(self.Pi, 2 + j)) for j, k in enumerate(self.keys)]
                probs.append((f["user"], p["user_trans.weight"], p["user_trans.bias"], self.P_usr))
                probs.sort(key=lambda t: -t[0].shape[1])                       # long-K tiles first
                ops.proj_fwd_group(probs, d, m)

    def _prop_fwd(self, with_feats=None, after_sides=None, with_ids=True, opset=None):
        """with_feats=False: the ID layers only (the hoisted mode propagates no side-feature operand); with_ids=False: the side-feature
        operands only (train_step runs the ID layers as a branch beside the projections).  opset: (ui, iu) to launch through.
        after_sides: called once Fu and Fi exist (after the second product; at once without side features) -- train_step forks the
        first touch of the gradient buffers there."""
        L, S = self.L, self.S
        wf = self.has_feats if with_feats is None else with_feats
        ui, iu = (self.ui, self.iu) if opset is None else opset
        if after_sides is not None and not wf:
            after_sides()
        # step t even: ui (I_{t/2} -> U_{t/2+1});  t odd: iu (U_{(t+1)/2} -> I_{(t+1)/2}); softmax on the last layer
        n_steps = max(2 * L if with_ids else 0, 3 if wf else 0)
        for t in range(n_steps):
            segs = []
            if t % 2 == 0:
                l = t // 2 + 1
                if wf and t == 0:
                    segs += [(self.blk(self.Pi, s), self.blk(self.Fu, s), None, False) for s in range(S)]                # :153,156,162
                if wf and t == 2:
                    segs.append((self.prof_i, self.prof_u, None, False))                                               # :167
                if with_ids and l <= L:
                    segs.append((self.Il[l - 1], self.Ul[l], None, l == L))                                            # :174,178
                with self._t("spmm_fwd"):
                    ui.apply(segs)
            else:
                l = (t + 1) // 2
                if wf and t == 1:
                    segs += [(self.blk(self.Fu, s), self.blk(self.Fi, s), None, False) for s in range(S)]                # :154,157,163
                    segs.append((self.P_usr, self.prof_i, None, False))                                                # :166
                if with_ids and l <= L:
                    segs.append((self.Ul[l], self.Il[l], None, l == L))                                                # :175,180
                with self._t("spmm_fwd"):
                    iu.apply(segs)
                if after_sides is not None and wf and t == 1:
                    after_sides()

    def _fuse_fwd(self):
        c = self.cfg
        if self.has_feats:
            coefs = [c.model_cat_rate, c.model_cat_rate, c.user_cat_rate] + [c.item_cat_rate] * len(self.keys)
            su = [self.blk(self.Fu, 0), self.blk(self.Fu, 1), self.prof_u] + [self.blk(self.Fu, 2 + j) for j in range(len(self.keys))]
            si = [self.blk(self.Fi, 0), self.blk(self.Fi, 1), self.prof_i] + [self.blk(self.Fi, 2 + j) for j in range(len(self.keys))]
        else:
            coefs, su, si = [], [], []
        with self._t("fuse_fwd"):
            self._fork(lambda: ops.fuse_fwd(self.Ul, su, coefs, self.U))                                               # :185-197
            ops.fuse_fwd(self.Il, si, coefs, self.I)
            self._join()
        self._fuse_args = (coefs, su, si)

    # ---- backward: expects gU, gI and (GFu, GFi, Gprof_u, Gprof_i, GP_usr_direct) filled ---------------
    def backward(self, gp_usr_direct=None, gpi_direct=None):
        self._fuse_bwd()
        self._chain_bwd(gp_usr_direct, gpi_direct)
        self._wgrad()
        return self.grads

    def _fuse_bwd(self):
        L = self.L
        coefs, su, si = self._fuse_args
        if self.has_feats:
            dsu = [self.blk(self.GFu, 0), self.blk(self.GFu, 1), self.Gprof_u] + [self.blk(self.GFu, 2 + j) for j in range(len(self.keys))]
            dsi = [self.blk(self.GFi, 0), self.blk(self.GFi, 1), self.Gprof_i] + [self.blk(self.GFi, 2 + j) for j in range(len(self.keys))]
        else:
            dsu, dsi = [], []
        with self._t("fuse_bwd"):
            self._fork(lambda: ops.fuse_bwd(self.gU, L + 1, self.dUl, su, coefs, dsu, True))
            ops.fuse_bwd(self.gI, L + 1, self.dIl, si, coefs, dsi, True)
            self._join()

    def _chain_bwd(self, gp_usr_direct=None, gpi_direct=None, with_feats=None, with_ids=True, opset=None):
        """with_feats=False: the ID chain only; with_ids=False: the side-feature operands only; opset: (uiT, iuT) to launch through."""
        L, S = self.L, self.S
        wf = self.has_feats if with_feats is None else with_feats
        uiT, iuT = (self.uiT, self.iuT) if opset is None else opset
        if wf:
            # prof_u = ui . prof_i  ->  Gprof_i += ui^T Gprof_u
            with self._t("spmm_bwd"):
                uiT.apply([(self.Gprof_u, self.Gprof_i, self.Gprof_i, False)])
        gE_i = self.grads["item_id_embedding.weight"]
        g_cur_I = self.dIl
        for l in range(L, 0, -1):
            if not with_ids and l < L:
                break
            # I_l = [softmax] iu . U_l
            segs = []
            if with_ids:
                if l == L:
                    with self._t("softmax_bwd"):
                        src = ops.row_softmax_bwd(self.Il[l], g_cur_I, out=self.tmpI)
                else:
                    src = g_cur_I
                segs.append((src, self.bufU, self.dUl, False))
            if wf and l == L:
                segs += [(self.blk(self.GFi, s), self.blk(self.GFu, s), self.blk(self.GFu, s), False) for s in range(S)]
                segs.append((self.Gprof_i, self.GP_usr, gp_usr_direct, False))
            with self._t("spmm_bwd"):
                iuT.apply(segs)
            # U_l = [softmax] ui . I_{l-1}
            segs = []
            dst = gE_i if l == 1 else self.bufI
            if with_ids:
                if l == L:
                    with self._t("softmax_bwd"):
                        ops.row_softmax_bwd(self.Ul[l], self.bufU, out=self.bufU)
                segs.append((self.bufU, dst, self.dIl, False))
            if wf and l == L:
                segs += [(self.blk(self.GFu, s), self.blk(self.GPi, s), self.blk(gpi_direct, s) if gpi_direct is not None else None, False)
                         for s in range(S)]
            with self._t("spmm_bwd"):
                uiT.apply(segs)
            g_cur_I = dst

    def _wgrad(self):
        d, m = self.d, self.cfg.proj_mode
        if self.has_feats:
            f, g = self.fx, self.grads
            with self._t("proj_wgrad"):
                probs = [(f["item"][k], self.blk(self.GPi, 2 + j), g["item_trans.weight"], g["item_trans.bias"], j > 0) for j, k in enumerate(self.keys)]
                probs.append((f["user"], self.GP_usr, g["user_trans.weight"], g["user_trans.bias"], False))
                probs.append((f["text"], self.blk(self.GPi, 1), g["text_trans.weight"], g["text_trans.bias"], False))
                probs.append((f["image"], self.blk(self.GPi, 0), g["image_trans.weight"], g["image_trans.bias"], False))
                ops.proj_wgrad_group(probs, d, m)

    # ---- losses + their gradients w.r.t. the forward outputs ---------------------------------------------
    def batch_capacity(self):
        """Largest B' a step of this configuration can see: batch_size sampled + int(batch_size * aug_sample_rate) augmented
        triplets (main.py:217-224), rounded up to a multiple of 8."""
        c = self.cfg
        return (c.batch_size + int(c.batch_size * c.aug_sample_rate) + 7) // 8 * 8

    def ensure_capacity(self, cap):
        """Index buffer [4 x cap] (rows users, pos, neg, meta = {B', n_keep}), the per-B' meta table and the BPR work block,
        sized ONCE for a batch capacity: every captured graph reads these addresses, so they are only ever replaced together
        with the graphs (a per-B' work buffer freed under a live graph was a use-after-free)."""
        if getattr(self, "_cap", 0) >= cap:
            return
        dev = self.E_u.device
        if getattr(self, "_graph", None) is not None:
            torch.cuda.synchronize()
        self._graph = None
        self._cap = int(cap)
        self._gidx = torch.zeros((4, self._cap), dtype=torch.int32, device=dev)
        keep = [int((1 - self.cfg.prune_loss_drop_rate) * b) for b in range(self._cap + 1)]      # main.py:161-162 (double arithmetic)
        self._meta_table = torch.tensor([[b, k] for b, k in enumerate(keep)], dtype=torch.int32).to(dev)
        self._bpr_work = ops.bpr_work(self.n_heads, self._cap, dev)

    def loss_and_output_grads(self, users, pos, neg, meta=None, init_done=False):
        """users/pos/neg: int32 CUDA tensors of equal length B' (sampled + augmented triplets) -- or, with `meta` (int32 CUDA
        {B', n_keep}), capacity-sized buffers whose first B' entries are live (the CUDA-graph path).
        init_done: `_grad_init` already ran for this step (train_step forks it beside the tail of the forward pass)."""
        c = self.cfg
        B = int(users.numel())
        self.ensure_capacity(max(B, self.batch_capacity()))
        n_keep = int((1 - c.prune_loss_drop_rate) * B)                     # main.py:161-162 (double arithmetic)
        heads = [(self.U, self.I, self.gU, self.gI, 1.0, 1.0)]                                                        # main.py:232-235
        if self.has_feats:
            heads.append((self.blk(self.Fu, 0), self.blk(self.Fi, 0), self.blk(self.GFu, 0), self.blk(self.GFi, 0), c.mm_mf_rate, 0.0))  # :238-241
            heads.append((self.blk(self.Fu, 1), self.blk(self.Fi, 1), self.blk(self.GFu, 1), self.blk(self.GFi, 1), c.mm_mf_rate, 0.0))  # :242-246
            for j in range(len(self.keys)):                                                                            # :248-254
                heads.append((self.prof_u, self.blk(self.Fi, 2 + j), self.Gprof_u, self.blk(self.GFi, 2 + j), c.aug_mf_rate, 0.0))
        if not init_done:
            self._grad_init()
        with self._t("bpr"):
            ops.bpr_heads(heads, users, pos, neg, n_keep, c.regs0 / c.batch_size, self.head_out, self.loss, self._bpr_work, meta=meta)
        return self.loss

    def _grad_init(self):
        """First touch of every gradient buffer the loss heads accumulate into: the feat_reg gradient c*X on the image/text blocks (its
        value starts the loss, main.py:151-156), zeros elsewhere.  Needs Fu / Fi only, not the fused outputs."""
        c = self.cfg
        regions = [(self.gU, None, 0.0), (self.gI, None, 0.0)]
        if self.has_feats:
            creg = c.feat_reg_decay / self.ni
            d2 = 2 * self.d
            regions += [(self.GFu[:, :d2], self.Fu[:, :d2], creg), (self.GFi[:, :d2], self.Fi[:, :d2], creg),
                        (self.Gprof_u, None, 0.0), (self.Gprof_i, None, 0.0)]
            if self.S > 2:
                regions += [(self.GFu[:, d2:], None, 0.0), (self.GFi[:, d2:], None, 0.0)]
        with self._t("grad_init"):
            ops.grad_init(regions, self.loss)

    def train_step(self, users, pos, neg, meta=None):
        """forward + losses + backward + AdamW; everything stays on the current stream."""
        if self.opt is None:
            raise RuntimeError("attach an optimizer with set_optimizer() first")
        split = self.has_feats and self.timer is None and (self.branches or self.force_split)
        if split:
            # the ID layers do not depend on the projections: they run as a branch (through operators with their own long-row scratch)
            # beside the projection kernel and the side-feature products; the first touch of the gradient buffers follows on the branch.
            # (Giving the two single-operand user-profile products a lane of their own was measured: no gain, removed.)
            ids = self._opset(0)
            if getattr(self, "_ev_ids", None) is None and self.branches:
                self._ev_ids = torch.cuda.Event()

            def id_layers():
                self._prop_fwd(with_feats=False, opset=ids[:2])
                if self.branches:
                    self._ev_ids.record()                                    # on the branch

            self._fork(id_layers, lane=0)
            self._proj_fwd()
            self._prop_fwd(with_ids=False, after_sides=lambda: self._fork(self._grad_init, lane=0))
            if self.branches:
                torch.cuda.current_stream().wait_event(self._ev_ids)         # the item-side fusion below reads Il; grad_init may still run
        else:
            self._proj_fwd()
            self._prop_fwd(after_sides=lambda: self._fork(self._grad_init))  # branch: runs beside the remaining products and the fusion
        self._fuse_fwd()                                                     # joins
        self._join()
        self.loss_and_output_grads(users, pos, neg, meta, init_done=True)
        if split:
            self._fuse_bwd()
            self._fork(lambda: self._chain_bwd(with_feats=False, opset=ids[2:]), lane=0)    # ID chain
            self._chain_bwd(with_ids=False)                                                  # side-feature operands
            self._wgrad()
            self._join()
        else:
            self.backward()
        with self._t("adamw"):
            self.opt.step([self.grads[k] for k in self._opt_names])
        return self.loss

    def families(self, users, pos, neg):
        """name -> thunk launching one kernel family of the step (bench.py times each from its own CUDA graph)."""
        f = {"spmm_fwd": self._prop_fwd, "fuse_fwd": self._fuse_fwd, "loss_heads": lambda: self.loss_and_output_grads(users, pos, neg),
             "fuse_bwd": self._fuse_bwd, "spmm_bwd": self._chain_bwd, "adamw": lambda: self.opt.step([self.grads[k] for k in self._opt_names])}
        if self.has_feats:
            f.update(proj_fwd=self._proj_fwd, proj_wgrad=self._wgrad)
        return f

    # ---- CUDA-graph replay of the whole step -----------------------------------------------------------
    def index_buffer(self, need):
        """The static [4 x cap] int32 buffer the captured step reads: rows users / pos / neg, row 3 = {B', n_keep, ...}."""
        self.ensure_capacity(max(int(need), self.batch_capacity()))
        return self._gidx

    def meta_row(self, B):
        """(B', n_keep) for the host side of a staging buffer (same double arithmetic as main.py:161-162)."""
        return int(B), int((1 - self.cfg.prune_loss_drop_rate) * B)

    def replay_staged(self):
        """Replay the captured step on whatever the index buffer holds (Trainer copies a pinned staging slot straight into it).
        ONE graph serves every batch length: the kernels take B' and n_keep from row 3 of the buffer."""
        if getattr(self, "_graph", None) is None:
            cap = self._cap
            u, p, n, meta = self._gidx[0], self._gidx[1], self._gidx[2], self._gidx[3]
            torch.cuda.synchronize()
            held = self._gidx.clone()
            if not getattr(self, "_warm", False):
                # one eager step at full capacity sizes every lazily allocated scratch buffer; its parameter update is undone
                snap = self._snapshot_state()
                self._gidx[3, :2].copy_(self._meta_table[cap])
                if self.pre_step is not None:
                    self.pre_step()
                self.train_step(u, p, n, meta)
                self._restore_state(snap)
                if self.pre_step is not None and self.pre_step_undo is not None:
                    self.pre_step_undo()
                self._gidx.copy_(held)
                self._warm = True
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if self.pre_step is not None:
                    self.pre_step()
                self.train_step(u, p, n, meta)
            self._graph = g
        self._graph.replay()
        return self.loss

    def train_step_graphed(self, users, pos, neg):
        """Same as train_step, replayed from the CUDA graph (the ~40 launches of a step cost more host time than device time at
        netflix scale).  users/pos/neg: int32 CUDA tensors of length B'; they are copied into the static index buffer."""
        B = int(users.numel())
        gi = self.index_buffer(B)
        gi[0, :B].copy_(users, non_blocking=True)
        gi[1, :B].copy_(pos, non_blocking=True)
        gi[2, :B].copy_(neg, non_blocking=True)
        gi[3, :2].copy_(self._meta_table[B], non_blocking=True)
        return self.replay_staged()

    def _snapshot_state(self):
        o = self.opt
        return ([p.clone() for p in o.params], [m.clone() for m in o.m], [v.clone() for v in o.v], o.state.clone())

    def _restore_state(self, snap):
        o = self.opt
        for dst, src in zip(o.params, snap[0]):
            dst.copy_(src)
        for dst, src in zip(o.m, snap[1]):
            dst.copy_(src)
        for dst, src in zip(o.v, snap[2]):
            dst.copy_(src)
        o.state.copy_(snap[3])

    def set_optimizer(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        names = [k for k in PARAM_ORDER if k in self.p and (self.has_feats or k.endswith("embedding.weight"))]
        self._opt_names = names
        self.opt = ops.AdamW([self.p[k] for k in names], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        return self.opt
