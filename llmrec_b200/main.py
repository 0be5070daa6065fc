"""Trainer / CLI with the reference's interface (main.py:37-369) on the B200 hot path.

    python main.py --dataset netflix [all flags of utility/parser.py]

Kept: Trainer(data_config), .train(), .test(users, is_val), bpr_loss / prune_loss /
feat_reg_loss_calculation / csr_norm / matrix_to_tensor, the epoch log lines, NaN exit, early stopping.
Changed on purpose (same results, fewer host round trips):
  * the whole step (forward, 8 BPR heads + prune, backward, AdamW) is engine.HotPath.train_step;
    no per-head D2H argsort (main.py:159), no float(loss) syncs per step (:280-283) -- losses are
    accumulated on the device and read once per epoch;
  * augmented_sample_dict is unpickled once, not every batch (main.py:216; the file never changes);
  * the derived `*_final` / `augmented_total_embed_dict` files are NOT written back into the data
    directory (main.py:66,78 side effects);
  * clip_grad_norm_ before zero_grad (main.py:274) is a no-op upstream and is omitted.
"""
from __future__ import annotations

import math
import os
import pickle
import random
import sys
from datetime import datetime
from time import time

import numpy as np
import scipy.sparse as sp
import torch

from .Models import Decoder, MM_Model
from .graph import BipartiteGraph
from .runtime import get_args, set_args
from .utility import batch_test
from .utility.load_data import Data
from .utility.logging import Logger
from .utility.parser import parse_args, resolve_dataset_dir


def set_seed(seed):
    """main.py:355-359"""
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class _StagingSlot:
    """One pinned [4 x cap] int32 index buffer (rows users / pos / neg / meta = {B', n_keep}) + the event recorded after its
    H2D copy.  A slot is rewritten only after that copy has completed, so the host may run ahead of the device by at most the
    ring length."""

    def __init__(self, cap):
        self.host = torch.zeros((4, cap), dtype=torch.int32).pin_memory()
        self.np = self.host.numpy()
        self.event = torch.cuda.Event()


def _stack_rows(obj):
    """indexable[n] -> ndarray [n x dim] (main.py:61-65, 73-77)."""
    if isinstance(obj, np.ndarray):
        return obj
    return np.array([obj[i] for i in range(len(obj))])


class Trainer(object):
    def __init__(self, data_config=None, data_generator=None, device="cuda"):
        args = self.args = get_args()
        if not torch.cuda.is_available():
            raise RuntimeError("llmrec_b200.Trainer needs a CUDA (B200) device: there is no CPU fallback")
        self.device = torch.device(device)
        self.task_name = "%s_%s_%s" % (datetime.now().strftime("%Y-%m-%d %H:%M:%S"), args.dataset, args.cf_model)
        self.logger = Logger(filename=self.task_name, is_debug=args.debug)
        self.logger.logging("PID: %d" % os.getpid())
        self.logger.logging(str(args))
        self.mess_dropout = eval(args.mess_dropout)
        self.lr, self.emb_dim, self.batch_size = args.lr, args.embed_size, args.batch_size
        self.weight_size = eval(args.weight_size)
        self.n_layers = len(self.weight_size)
        self.regs = eval(args.regs)
        self.decay = self.regs[0]

        ddir = self.data_dir = resolve_dataset_dir(args.data_path, args.dataset)
        if data_generator is None:
            data_generator = batch_test.data_generator
        if data_generator is None:
            data_generator = Data(path=ddir, batch_size=args.batch_size, sampler=getattr(args, "host_sampler", "python"))
        self.data_generator = data_generator
        batch_test.init(data_generator, args)

        rd = lambda name: pickle.load(open(os.path.join(ddir, name), "rb"))
        self.image_feats = np.load(os.path.join(ddir, "image_feat.npy"))                     # main.py:54-55
        self.text_feats = np.load(os.path.join(ddir, "text_feat.npy"))
        self.image_feat_dim, self.text_feat_dim = self.image_feats.shape[-1], self.text_feats.shape[-1]
        if os.path.exists(os.path.join(ddir, "train_mat")):
            self.ui_graph_raw = rd("train_mat")                                              # :59
        else:                                                                                # CSR store (utility/csr_store.py): same matrix, no pickle
            rowptr, col = data_generator.csr("train")
            self.ui_graph_raw = sp.csr_matrix((np.ones(col.shape[0], dtype=np.float32), col, rowptr),
                                              shape=(data_generator.n_users, data_generator.n_items))
        self.user_init_embedding = _stack_rows(rd("augmented_user_init_embedding"))          # :61-67
        raw_att = rd("augmented_atttribute_embedding_dict")                                  # :69-79
        self.item_attribute_embedding = {k: _stack_rows(raw_att[k]) for k in raw_att}
        self.augmented_sample_dict = rd("augmented_sample_dict")                             # :216 (loaded once)

        self.n_users, self.n_items = self.ui_graph_raw.shape                                 # :84-85
        self.graph = BipartiteGraph(self.ui_graph_raw, self.device)
        self.ui_graph, self.iu_graph = self.graph.coo_tensors()                              # :88-91
        self.image_ui_graph = self.text_ui_graph = self.ui_graph                             # :92-93
        self.image_iu_graph = self.text_iu_graph = self.iu_graph

        self.model_mm = MM_Model(self.n_users, self.n_items, self.emb_dim, self.weight_size, self.mess_dropout, self.image_feats,
                                 self.text_feats, self.user_init_embedding, self.item_attribute_embedding)   # built on CPU: RNG parity
        self.model_mm = self.model_mm.to(self.device)
        # main.py:97: the Decoder is built right after the model (its two nn.Linear inits consume the CPU generator before the first
        # torch.randperm of the mask branch); its optimizer exists upstream but is never stepped (main.py:106-110)
        self.decoder = Decoder(self.user_init_embedding.shape[1]).to(self.device)
        self.masked_mode = bool(args.mask) or args.mask_rate > 0 or args.drop_rate > 0
        # --hoist_side 1: only sound while dropout is the identity and the mask branch is off
        self.hoisted = bool(getattr(args, "hoist_side", 0)) and not (args.mask or args.mask_rate > 0 or args.drop_rate > 0)
        if self.hoisted:
            self.hot = self.model_mm.hot_path(self.ui_graph, self.iu_graph, hoisted=True, graph_scalars=self.graph.ones_propagated())
        else:
            self.hot = self.model_mm.hot_path(self.ui_graph, self.iu_graph)
        # torch.optim.AdamW defaults: betas (0.9, 0.999), eps 1e-8, weight_decay 0.01 (main.py:100-104)
        self.optimizer = self.hot.set_optimizer(lr=self.lr)
        self._slots, self._slot_i, self._idx_dev = [], 0, None
        # whole batches (users, items, augmented edges) from one C call, same `random` / `np.random` streams (host_native.BatchSampler)
        self._batch_sampler = None
        if getattr(data_generator, "_sampler", "python") == "native":
            from .host_native import BatchSampler
            rowptr, col = data_generator.csr("train")
            aug_pos, aug_neg = BatchSampler.aug_tables(self.augmented_sample_dict, data_generator.n_users, self.n_items)
            self._batch_sampler = BatchSampler(data_generator.exist_users, rowptr, col, data_generator.n_items, data_generator.batch_size,
                                               aug_pos, aug_neg, aug_limit=self.n_items)
            self._batch_np = np.empty((3, 2 * data_generator.batch_size + 8), dtype=np.int32)
        self.use_graph = bool(getattr(args, "cuda_graph", 1))
        self._epoch_stats = torch.zeros(4, dtype=torch.float32, device=self.device)          # total, mf, emb, interactions (device-side sampler)
        # --device_sampler 1: batches are drawn ON the GPU into the engine's index buffer, in front of every step
        self.device_sampler = None
        if getattr(args, "device_sampler", 0) and not self.masked_mode:
            from .device_sampler import DeviceSampler
            from .host_native import BatchSampler
            rowptr, col = data_generator.csr("train", sorted_rows=True)
            aug_pos, aug_neg = BatchSampler.aug_tables(self.augmented_sample_dict, data_generator.n_users, self.n_items)
            self.device_sampler = DeviceSampler(data_generator.exist_users, rowptr, col, data_generator.n_items, data_generator.batch_size,
                                                aug_pos, aug_neg, self.n_items, args.aug_sample_rate, self.device, seed=args.seed)
            gi = self.hot.index_buffer(self.hot.batch_capacity())
            self.hot.pre_step = lambda: self.device_sampler.fill(self.hot._gidx, self.hot._meta_table)
            self.hot.pre_step_undo = lambda: self.device_sampler.state[1:2].sub_(1)

    # ---- reference helper API (same names / returns) ----------------------------------------------
    def csr_norm(self, csr_mat, mean_flag=False):
        """main.py:114-126"""
        rowsum = np.array(csr_mat.sum(1))
        rowsum = np.power(rowsum + 1e-8, -0.5).flatten()
        rowsum[np.isinf(rowsum)] = 0.0
        left = sp.diags(rowsum) * csr_mat
        if mean_flag:
            return left
        colsum = np.array(csr_mat.sum(0))
        colsum = np.power(colsum + 1e-8, -0.5).flatten()
        colsum[np.isinf(colsum)] = 0.0
        return left * sp.diags(colsum)

    def matrix_to_tensor(self, cur_matrix):
        """main.py:128-134"""
        coo = cur_matrix.tocoo()
        idx = torch.from_numpy(np.vstack((coo.row, coo.col)).astype(np.int64))
        return torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data), coo.shape).to(torch.float32).to(self.device)

    def prune_loss(self, pred, drop_rate):
        """Mean of the int((1-drop_rate)*n) smallest entries (main.py:158-165), selected on the device."""
        n_keep = int((1 - drop_rate) * len(pred))
        order = torch.argsort(pred.detach(), stable=True)
        return pred[order[:n_keep]].mean()

    def bpr_loss(self, users, pos_items, neg_items):
        """(mf_loss, emb_loss, reg_loss) of main.py:330-342 for pre-gathered rows (torch autograd API)."""
        pos_scores = torch.sum(users * pos_items, dim=1)
        neg_scores = torch.sum(users * neg_items, dim=1)
        regularizer = 1.0 / (2 * (users ** 2).sum() + 1e-8) + 1.0 / (2 * (pos_items ** 2).sum() + 1e-8) + 1.0 / (2 * (neg_items ** 2).sum() + 1e-8)
        regularizer = regularizer / self.batch_size
        maxi = torch.nn.functional.logsigmoid(pos_scores - neg_scores + 1e-8)
        mf_loss = -self.prune_loss(maxi, self.args.prune_loss_drop_rate)
        return mf_loss, self.decay * regularizer, 0.0

    def feat_reg_loss_calculation(self, g_item_image, g_item_text, g_user_image, g_user_text):
        """main.py:151-156"""
        feat_reg = 0.5 * (g_item_image ** 2).sum() + 0.5 * (g_item_text ** 2).sum() + 0.5 * (g_user_image ** 2).sum() + 0.5 * (g_user_text ** 2).sum()
        return self.args.feat_reg_decay * (feat_reg / self.n_items)

    # ---- evaluation ----------------------------------------------------------------------------------
    def test(self, users_to_test, is_val):
        """main.py:182-187: full forward in eval mode, then test_torch."""
        self.model_mm.eval()
        with torch.no_grad():
            if self.masked_mode:
                self._mask_features()                 # Models.py:131-142 runs in eval mode too (and keeps mutating the feature buffers)
            ua_embeddings, ia_embeddings = self.hot.forward()
        return batch_test.test_torch(ua_embeddings, ia_embeddings, users_to_test, is_val)

    # ---- optional branch: feature mask / dropout / attribute restoration (Models.py:131-150, main.py:258-271) ---------
    def _mask_features(self):
        """Models.py:131-142: random rows of the attribute / profile tables are overwritten IN PLACE (persistently) with the column mean.
        torch.randperm draws from the CPU generator in the reference's order: items first (only with --mask), then users (always)."""
        args, m = self.args, self.model_mm
        i_mask = None
        if args.mask:
            i_mask = torch.randperm(self.n_items)[:int(args.mask_rate * self.n_items)].to(self.device)
            for k in m._item_keys:
                f = getattr(m, "item_feat__" + k)
                f[i_mask] = f.mean(0)
        u_mask = torch.randperm(self.n_users)[:int(args.mask_rate * self.n_users)].to(self.device)
        m.user_feats[u_mask] = m.user_feats.mean(0)
        return i_mask, u_mask

    @staticmethod
    def sce_criterion(x, y, alpha=1):
        """main.py:175-180"""
        x = torch.nn.functional.normalize(x, p=2, dim=-1)
        y = torch.nn.functional.normalize(y, p=2, dim=-1)
        return (1 - (x * y).sum(dim=-1)).pow(alpha).mean()

    @staticmethod
    def mse_criterion(x, y, alpha=3):
        """main.py:167-173 (the cosine term is computed and discarded upstream; the result is the MSE of the normalised rows)"""
        x = torch.nn.functional.normalize(x, p=2, dim=-1)
        y = torch.nn.functional.normalize(y, p=2, dim=-1)
        return torch.nn.functional.mse_loss(x, y)

    def _train_batch_masked(self, u, p, n):
        """One training step with --mask / --mask_rate / --drop_rate: the same kernels, launched eagerly, with the branch-specific pieces
        (feature masking, dropout masks, Decoder + restoration loss) as plain torch ops between them.  Off by default upstream."""
        args, hp, m = self.args, self.hot, self.model_mm
        i_mask, u_mask = self._mask_features()
        hp._proj_fwd()
        drop = None
        if args.drop_rate > 0:
            # nn.Dropout in training mode on each projection, reference order image, text, user, item keys (Models.py:145-150); the mask of a
            # contiguous [n x d] tensor is drawn from the CUDA generator exactly as nn.Dropout would on the projection itself
            d = hp.d
            blocks = [hp.blk(hp.Pi, 0), hp.blk(hp.Pi, 1), hp.P_usr] + [hp.blk(hp.Pi, 2 + j) for j in range(len(hp.keys))]
            drop = [torch.nn.functional.dropout(torch.ones(b.shape[0], d, device=self.device), p=args.drop_rate, training=True) for b in blocks]
            for b, mk in zip(blocks, drop):
                b.mul_(mk)
        hp._prop_fwd()
        hp._fuse_fwd()
        hp.loss_and_output_grads(u, p, n)
        if args.mask and args.att_re_rate != 0:
            # main.py:258-271: decoder on the DETACHED masked profile rows (torch.tensor(...) upstream copies) and on the masked rows of the
            # propagated attribute features (these keep their graph: the gradient flows back into GFi)
            v = hp.side_views()
            keys = hp.keys
            leaf = {k: v["att_i"][k][i_mask].detach().clone().requires_grad_(True) for k in keys}
            dec_u, dec_i = self.decoder(v["prof_u"][u_mask].detach(), leaf)
            crit = self.mse_criterion if args.feat_loss_type == "mse" else self.sce_criterion
            raw_u = torch.as_tensor(self.user_init_embedding[u_mask.cpu().numpy()], device=self.device).float()
            att = crit(dec_u, raw_u, alpha=args.alpha_l)
            for j, k in enumerate(keys):
                raw_i = torch.as_tensor(self.item_attribute_embedding[k][i_mask.cpu().numpy()], device=self.device).float()
                att = att + crit(dec_i[j], raw_i, alpha=args.alpha_l)
            (args.att_re_rate * att).backward()
            self.decoder.zero_grad(set_to_none=True)                      # de_optimizer is never stepped upstream
            for j, k in enumerate(keys):
                hp.blk(hp.GFi, 2 + j).index_add_(0, i_mask, leaf[k].grad)
            hp.loss.add_(args.att_re_rate * att.detach())
        hp._fuse_bwd()
        hp._chain_bwd()
        if drop is not None:                                              # backward of the dropout: the same masks on the projection gradients
            gblocks = [hp.blk(hp.GPi, 0), hp.blk(hp.GPi, 1), hp.GP_usr] + [hp.blk(hp.GPi, 2 + j) for j in range(len(hp.keys))]
            for b, mk in zip(gblocks, drop):
                b.mul_(mk)
        hp._wgrad()
        hp.opt.step([hp.grads[k] for k in hp._opt_names])
        self._last_dropout_masks = drop
        return hp.loss

    # ---- one batch ---------------------------------------------------------------------------------------
    def sample_batch(self):
        """Data.sample() + augmented edges (main.py:213-224); host side, reference RNG order.  -> three lists."""
        if self._batch_sampler is not None:
            o = self._batch_np
            B = self._batch_sampler.draw(o, self.args.aug_sample_rate)
            self.new_batch_size = B - self._batch_sampler.batch
            return o[0, :B].tolist(), o[1, :B].tolist(), o[2, :B].tolist()
        users, pos_items, neg_items = self.data_generator.sample()
        aug = self.augmented_sample_dict
        ni = self.n_items
        users_aug = random.sample(users, int(len(users) * self.args.aug_sample_rate))
        keep = [u for u in users_aug if (aug[u][0] < ni and aug[u][1] < ni)]
        self.new_batch_size = len(keep)
        users = users + keep
        # a negative augmented id passes upstream's filter and wraps under Python indexing (row -1 = last item): same row here
        pos_items = pos_items + [aug[u][0] % ni for u in keep]
        neg_items = neg_items + [aug[u][1] % ni for u in keep]
        return users, pos_items, neg_items

    def _next_slot(self, need):
        """Next pinned staging slot of the ring (4 slots), free to be rewritten.  The device side of the copy is the engine's
        static index buffer (what the captured CUDA graph reads), so a batch crosses PCIe exactly once."""
        self._idx_dev = self.hot.index_buffer(need)
        cap = self._idx_dev.shape[1]
        if not self._slots or self._slots[0].host.shape[1] != cap:
            if self._slots:
                torch.cuda.synchronize()                                  # copies out of the old ring may still be in flight
            self._slots = [_StagingSlot(cap) for _ in range(4)]
            self._slot_i = 0
        slot = self._slots[self._slot_i]
        self._slot_i = (self._slot_i + 1) % len(self._slots)
        slot.event.synchronize()                                          # its previous H2D copy has completed
        return slot

    def _push(self, slot, B):
        slot.np[3, 0:2] = self.hot.meta_row(B)                            # {B', n_keep}: the graph's kernels read them from the device
        w = max(B, 2)
        self._idx_dev[:, :w].copy_(slot.host[:, :w], non_blocking=True)
        slot.event.record()
        self.last_h2d_bytes = 4 * 4 * w
        d = self._idx_dev
        return d[0, :B], d[1, :B], d[2, :B]

    def upload_batch(self, users, pos_items, neg_items):
        """[3 x B'] int32 through pinned memory; returns three device views."""
        B = len(users)
        slot = self._next_slot(B)
        slot.np[0, :B] = users
        slot.np[1, :B] = pos_items
        slot.np[2, :B] = neg_items
        return self._push(slot, B)

    def stage_batch(self):
        """sample_batch + upload_batch without the Python lists in between: the C sampler writes the batch straight into
        a pinned staging slot.  -> three device views (users, pos, neg)."""
        if self._batch_sampler is None:
            return self.upload_batch(*self.sample_batch())
        slot = self._next_slot(self.hot.batch_capacity())
        B = self._batch_sampler.draw(slot.np[:3], self.args.aug_sample_rate)
        self.new_batch_size = B - self._batch_sampler.batch
        return self._push(slot, B)

    def _step(self, u, p, n):
        # u, p, n are views of the engine's index buffer (see _push): the graph replays on what was just staged
        if self.masked_mode:
            loss = self._train_batch_masked(u, p, n)
        else:
            loss = self.hot.replay_staged() if self.use_graph else self.hot.train_step(u, p, n)
        # device-side epoch accumulators: [total, mf(main), emb(main)]
        self._epoch_stats[0:1] += loss
        self._epoch_stats[1:3] += self.hot.head_out[0:2]
        return loss

    def train_batch(self, users, pos_items, neg_items):
        return self._step(*self.upload_batch(users, pos_items, neg_items))

    def train_next_batch(self):
        """One iteration of the training loop (main.py:213-278): draw the next batch, run the step.  -> (loss tensor, B').
        With the device-side sampler B' is only known on the device (-1 here; Trainer.train reads the epoch total once)."""
        if self.device_sampler is not None:
            hp = self.hot
            if self.use_graph:
                loss = hp.replay_staged()
            else:
                hp.pre_step()
                gi = hp._gidx
                loss = hp.train_step(gi[0], gi[1], gi[2], gi[3])
            self._epoch_stats[0:1] += loss
            self._epoch_stats[1:3] += hp.head_out[0:2]
            self._epoch_stats[3:4] += hp._gidx[3, 0:1].float()
            return loss, -1
        u, p, n = self.stage_batch()
        return self._step(u, p, n), int(u.numel())

    # ---- training loop (main.py:189-326) -----------------------------------------------------------------
    def train(self):
        args, dg = self.args, self.data_generator
        run_time = datetime.strftime(datetime.now(), "%Y_%m_%d__%H_%M_%S")
        training_time_list = []
        stopping_step, best_recall, test_ret = 0, 0, None
        for epoch in range(args.epoch):
            t1 = time()
            n_batch = dg.n_train // args.batch_size + 1
            self._epoch_stats.zero_()
            self.n_interactions = 0
            self.model_mm.train()
            for _ in range(n_batch):
                self.n_interactions += max(self.train_next_batch()[1], 0)
            loss, mf_loss, emb_loss, n_dev = (float(x) for x in self._epoch_stats.tolist())      # the one sync per epoch
            self.n_interactions += int(n_dev)
            reg_loss, contrastive_loss = 0.0, 0.0
            if math.isnan(loss):
                self.logger.logging("ERROR: loss is nan.")
                sys.exit()
            if (epoch + 1) % args.verbose != 0:
                self.logger.logging("Epoch %d [%.1fs]: train==[%.5f=%.5f + %.5f + %.5f  + %.5f]" % (
                    epoch, time() - t1, loss, mf_loss, emb_loss, reg_loss, contrastive_loss))
                training_time_list.append(time() - t1)
            t2 = time()
            users_to_test = list(dg.test_set.keys())
            ret = self.test(users_to_test, is_val=False)
            training_time_list.append(t2 - t1)
            t3 = time()
            self.last_epoch_times = (t2 - t1, t3 - t2)
            if args.verbose > 0:
                r, p, h, n = ret["recall"], ret["precision"], ret["hit_ratio"], ret["ndcg"]
                self.logger.logging(
                    "Epoch %d [%.1fs + %.1fs]: train==[%.5f=%.5f + %.5f + %.5f], recall=[%.5f, %.5f, %.5f, %.5f], "
                    "precision=[%.5f, %.5f, %.5f, %.5f], hit=[%.5f, %.5f, %.5f, %.5f], ndcg=[%.5f, %.5f, %.5f, %.5f]" % (
                        epoch, t2 - t1, t3 - t2, loss, mf_loss, emb_loss, reg_loss, r[0], r[1], r[2], r[-1], p[0], p[1], p[2], p[-1],
                        h[0], h[1], h[2], h[-1], n[0], n[1], n[2], n[-1]))
            if ret["recall"][1] > best_recall:
                best_recall = ret["recall"][1]
                test_ret = self.test(users_to_test, is_val=False)
                self.logger.logging("Test_Recall@%d: %.5f,  precision=[%.5f], ndcg=[%.5f]" % (
                    eval(args.Ks)[1], test_ret["recall"][1], test_ret["precision"][1], test_ret["ndcg"][1]))
                stopping_step = 0
            elif stopping_step < args.early_stopping_patience:
                stopping_step += 1
                self.logger.logging("#####Early stopping steps: %d #####" % stopping_step)
            else:
                self.logger.logging("#####Early stop! #####")
                break
        self.logger.logging(str(test_ret))
        return best_recall, run_time


def main(argv=None):
    args = set_args(parse_args(argv))
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(args.gpu_id))
    set_seed(args.seed)
    ddir = resolve_dataset_dir(args.data_path, args.dataset)
    gen = Data(path=ddir, batch_size=args.batch_size, sampler=args.host_sampler)
    batch_test.init(gen, args)
    config = dict(n_users=gen.n_users, n_items=gen.n_items)
    trainer = Trainer(data_config=config, data_generator=gen)
    return trainer.train()


if __name__ == "__main__":
    main()
