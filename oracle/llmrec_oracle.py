"""CPU oracle for the LLMRec training-and-eval hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU / numpy restatement of the reference's algorithm for the
path named in BASELINE.json.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it; the product package ``llmrec_b200`` never does (it fails loudly when the CUDA
library is missing instead of falling back here).

Parity pinning: the reference ships no tests or golden vectors, so
this restatement is pinned against the *unmodified reference itself*, imported from
/root/reference under the three compat shims of ``oracle/ref_shim.py``; the outputs of that
run are committed under ``tests/golden/`` by ``tests/golden/make_golden.py`` and checked by
``tests/test_oracle_golden.py`` (CPU) and the ``-m gpu`` suite (CUDA path vs the same
vectors).  Known-answer constants of SURVEY.md section 8a are checked as well.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import heapq
import json
import os
import pickle
import random as _pyrandom
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# configuration (defaults of utility/parser.py:7-54)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    embed_size: int = 64
    weight_size: tuple = (64, 64)
    batch_size: int = 1024
    lr: float = 1e-4
    regs0: float = 1e-5
    layers: int = 1
    model_cat_rate: float = 0.02
    user_cat_rate: float = 2.8
    item_cat_rate: float = 0.005
    aug_mf_rate: float = 0.012
    mm_mf_rate: float = 1e-4
    prune_loss_drop_rate: float = 0.71
    aug_sample_rate: float = 0.1
    feat_reg_decay: float = 1e-5
    Ks: tuple = (10, 20, 50)
    seed: int = 2022

    @property
    def n_ui_layers(self):
        return len(self.weight_size)


# ----------------------------------------------------------------------------------------
# data (utility/load_data.py:11-92, main.py:54-93)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleData:
    n_users: int
    n_items: int
    n_train: int
    exist_users: list
    train_items: dict
    test_set: dict
    val_set: dict
    image_feats: np.ndarray = None
    text_feats: np.ndarray = None
    user_feats: np.ndarray = None
    item_feats: dict = field(default_factory=dict)
    aug_samples: dict = None
    train_mat: sp.spmatrix = None


def load_dataset(path: str) -> OracleData:
    """JSON interaction dicts + feature files, as the reference's Data/Trainer read them.

    n_users = max train uid + 1 (load_data.py:35,55); n_items = rows of text_feat.npy
    (load_data.py:57-58); users with an empty train list are skipped (load_data.py:31-32).
    """
    def _js(name):
        with open(os.path.join(path, name)) as f:
            return json.load(f)

    tr, te, va = _js("train.json"), _js("test.json"), _js("val.json")
    exist, train_items, n_train, max_uid = [], {}, 0, 0
    for k, items in tr.items():
        if not items:
            continue
        u = int(k)
        exist.append(u)
        train_items[u] = items
        n_train += len(items)
        max_uid = max(max_uid, u)
    text = np.load(os.path.join(path, "text_feat.npy"))
    image = np.load(os.path.join(path, "image_feat.npy"))
    test_set = {int(k): v for k, v in te.items() if v}
    val_set = {int(k): v for k, v in va.items() if v}
    with open(os.path.join(path, "train_mat"), "rb") as f:
        train_mat = pickle.load(f)
    with open(os.path.join(path, "augmented_user_init_embedding"), "rb") as f:
        raw_u = pickle.load(f)
    user_feats = np.array([raw_u[i] for i in range(len(raw_u))])          # main.py:61-65
    with open(os.path.join(path, "augmented_atttribute_embedding_dict"), "rb") as f:
        raw_a = pickle.load(f)
    item_feats = {k: np.array([raw_a[k][i] for i in range(len(raw_a[k]))]) for k in raw_a}  # main.py:73-77
    with open(os.path.join(path, "augmented_sample_dict"), "rb") as f:
        aug = pickle.load(f)
    return OracleData(n_users=max_uid + 1, n_items=text.shape[0], n_train=n_train,
                      exist_users=exist, train_items=train_items, test_set=test_set,
                      val_set=val_set, image_feats=image, text_feats=text,
                      user_feats=user_feats, item_feats=item_feats, aug_samples=aug,
                      train_mat=train_mat)


def row_normalise(mat: sp.spmatrix) -> sp.spmatrix:
    """diag((rowsum + 1e-8)^-1/2) . A   -- csr_norm(mean_flag=True), main.py:114-126."""
    deg = np.asarray(mat.sum(1)).reshape(-1)
    s = np.power(deg + 1e-8, -0.5)
    s[np.isinf(s)] = 0.0
    return sp.diags(s) * mat


def to_torch_coo(mat: sp.spmatrix) -> torch.Tensor:
    """float64 scipy -> fp32 COO tensor with int64 indices (main.py:128-134)."""
    coo = mat.tocoo()
    idx = torch.from_numpy(np.vstack((coo.row, coo.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data), coo.shape).to(torch.float32)


def build_graphs(train_mat):
    """ui = norm(R), iu = norm(R^T)  (main.py:84-91)."""
    return to_torch_coo(row_normalise(train_mat)), to_torch_coo(row_normalise(train_mat.T))


# ----------------------------------------------------------------------------------------
# parameters (Models.py:20-55 construction / RNG order)
# ----------------------------------------------------------------------------------------
PARAM_NAMES = ("image_trans.weight", "image_trans.bias", "text_trans.weight", "text_trans.bias",
               "user_trans.weight", "user_trans.bias", "item_trans.weight", "item_trans.bias",
               "user_id_embedding.weight", "item_id_embedding.weight")


def init_params(cfg: OracleConfig, data: OracleData) -> dict:
    """Draw parameters from the CPU generator in the reference's construction order:
    4 x nn.Linear default init (image, text, user, item; Models.py:30-33), xavier_uniform_ on
    the 4 weights (:34-37), 2 x nn.Embedding normal init (:39-40), xavier_uniform_ x2 (:41-42).
    Caller seeds torch first (main.py:355-359)."""
    d = cfg.embed_size
    first_key = "title"                                                   # Models.py:33
    lins = [torch.nn.Linear(data.image_feats.shape[1], d), torch.nn.Linear(data.text_feats.shape[1], d),
            torch.nn.Linear(data.user_feats.shape[1], d), torch.nn.Linear(data.item_feats[first_key].shape[1], d)]
    for lin in lins:
        torch.nn.init.xavier_uniform_(lin.weight)
    eu = torch.nn.Embedding(data.n_users, d)
    ei = torch.nn.Embedding(data.n_items, d)
    torch.nn.init.xavier_uniform_(eu.weight)
    torch.nn.init.xavier_uniform_(ei.weight)
    vals = []
    for lin in lins:
        vals += [lin.weight, lin.bias]
    vals += [eu.weight, ei.weight]
    return {n: v.detach().clone().requires_grad_(True) for n, v in zip(PARAM_NAMES, vals)}


# ----------------------------------------------------------------------------------------
# forward (Models.py:127-199, default flags: mask off, dropout p=0)
# ----------------------------------------------------------------------------------------
def mask_features(feats: dict, n_users: int, n_items: int, mask: bool, mask_rate: float):
    """Models.py:131-142: in-place, persistent; torch.randperm on the CPU generator, items first (only with --mask), users always."""
    i_mask = None
    if mask:
        i_mask = torch.randperm(n_items)[:int(mask_rate * n_items)]
        for k in feats["item"]:
            feats["item"][k][i_mask] = feats["item"][k].mean(0)
    u_mask = torch.randperm(n_users)[:int(mask_rate * n_users)]
    feats["user"][u_mask] = feats["user"].mean(0)
    return i_mask, u_mask


def restoration_loss(out: dict, dec: dict, raw_user, raw_items: dict, i_mask, u_mask, alpha=2, kind="sce"):
    """main.py:258-271 + Models.py:203-225: decoder = Linear + LeakyReLU(negative_slope=True == 1.0, an identity) per side; the profile
    input is detached upstream (torch.tensor(...) copy), the attribute inputs keep their graph."""
    def crit(x, y):
        x, y = F.normalize(x, p=2, dim=-1), F.normalize(y, p=2, dim=-1)
        return F.mse_loss(x, y) if kind == "mse" else (1 - (x * y).sum(dim=-1)).pow(alpha).mean()
    dec_u = F.linear(out["prof_u"][u_mask].detach(), dec["u_w"], dec["u_b"])
    loss = crit(dec_u, raw_user[u_mask])
    for k in out["att_i"]:
        loss = loss + crit(F.linear(out["att_i"][k][i_mask], dec["i_w"], dec["i_b"]), raw_items[k][i_mask])
    return loss


def forward(params: dict, feats: dict, ui: torch.Tensor, iu: torch.Tensor, cfg: OracleConfig, drop=None) -> dict:
    """feats: {'image','text','user': tensors, 'item': {key: tensor}} fp32.
    drop: optional list of dropout masks (already scaled by 1/(1-p)) in the reference's order image, text, user, item keys
    (Models.py:145-150 with --drop_rate > 0); None = Dropout(p=0), the default."""
    def lin(x, name):
        return F.linear(x, params[name + ".weight"], params[name + ".bias"])

    p_img = lin(feats["image"], "image_trans")                            # Models.py:145
    p_txt = lin(feats["text"], "text_trans")                              # :146
    p_usr = lin(feats["user"], "user_trans")                              # :147
    p_att = {k: lin(v, "item_trans") for k, v in feats["item"].items()}   # :148-150
    if drop is not None:
        p_img, p_txt, p_usr = p_img * drop[0], p_txt * drop[1], p_usr * drop[2]
        p_att = {k: v * drop[3 + j] for j, (k, v) in enumerate(p_att.items())}

    spmm = torch.sparse.mm
    img_u = spmm(ui, p_img); img_i = spmm(iu, img_u)                      # :152-157 (layers>=1: same result)
    txt_u = spmm(ui, p_txt); txt_i = spmm(iu, txt_u)
    att_u, att_i = {}, {}
    for k in p_att:                                                       # :160-163
        att_u[k] = spmm(ui, p_att[k])
        att_i[k] = spmm(iu, att_u[k])
    prof_i = spmm(iu, p_usr)                                              # :166
    prof_u = spmm(ui, prof_i)                                             # :167

    e_u = params["user_id_embedding.weight"]
    e_i = params["item_id_embedding.weight"]
    us, its = [e_u], [e_i]
    L = cfg.n_ui_layers
    for l in range(L):                                                    # :172-183, sequential u -> i
        e_u = torch.mm(ui, e_i)
        if l == L - 1:
            e_u = torch.softmax(e_u, dim=-1)
        e_i = torch.mm(iu, e_u)
        if l == L - 1:
            e_i = torch.softmax(e_i, dim=-1)
        us.append(e_u); its.append(e_i)
    U = torch.mean(torch.stack(us), dim=0)                                # :185-186
    I = torch.mean(torch.stack(its), dim=0)

    n = lambda x: F.normalize(x, p=2, dim=1)
    U = U + cfg.model_cat_rate * n(img_u) + cfg.model_cat_rate * n(txt_u)  # :188
    I = I + cfg.model_cat_rate * n(img_i) + cfg.model_cat_rate * n(txt_i)  # :189
    U = U + cfg.user_cat_rate * n(prof_u)                                 # :191
    I = I + cfg.user_cat_rate * n(prof_i)                                 # :192
    for k in p_att:                                                       # :195-197
        U = U + cfg.item_cat_rate * n(att_u[k])
        I = I + cfg.item_cat_rate * n(att_i[k])
    return dict(U=U, I=I, img_i=img_i, txt_i=txt_i, img_u=img_u, txt_u=txt_u, p_usr=p_usr,
                att_i=att_i, prof_u=prof_u, prof_i=prof_i, att_u=att_u)


# ----------------------------------------------------------------------------------------
# losses (main.py:330-342, 158-165, 151-156, 273)
# ----------------------------------------------------------------------------------------
def num_remember(n: int, drop_rate: float) -> int:
    """int((1 - drop_rate) * n) in Python double arithmetic (main.py:161-162)."""
    return int((1 - drop_rate) * n)


def prune_mean(pred: torch.Tensor, drop_rate: float) -> torch.Tensor:
    """Mean of the num_remember smallest entries (ascending argsort; main.py:158-165).  The reference sorts on the HOST
    (`np.argsort(pred.cpu().data).cuda()`, main.py:159): with device tensors that is a D2H copy + CPU sort + H2D per head,
    reproduced here so the torch-on-GPU baseline leg pays what the reference pays."""
    order = torch.argsort(pred.detach().cpu(), stable=True).to(pred.device)
    keep = order[:num_remember(pred.shape[0], drop_rate)]
    return pred[keep].mean()


def bpr_head(u, p, n, cfg: OracleConfig):
    """-> (mf_loss, emb_loss).  regulariser is the RECIPROCAL of 2*sum-of-squares (+1e-8),
    divided by the flag batch_size, not the grown batch (main.py:334-341)."""
    pos = (u * p).sum(1)
    neg = (u * n).sum(1)
    reg = 1.0 / (2 * (u ** 2).sum() + 1e-8) + 1.0 / (2 * (p ** 2).sum() + 1e-8) + 1.0 / (2 * (n ** 2).sum() + 1e-8)
    reg = reg / cfg.batch_size
    maxi = F.logsigmoid(pos - neg + 1e-8)
    return -prune_mean(maxi, cfg.prune_loss_drop_rate), cfg.regs0 * reg


def batch_loss(out: dict, users, pos, neg, n_items: int, cfg: OracleConfig):
    """Loss assembly of main.py:232-256,273 (att_re term is 0 with default flags)."""
    mf, emb = bpr_head(out["U"][users], out["I"][pos], out["I"][neg], cfg)
    mf_img, _ = bpr_head(out["img_u"][users], out["img_i"][pos], out["img_i"][neg], cfg)
    mf_txt, _ = bpr_head(out["txt_u"][users], out["txt_i"][pos], out["txt_i"][neg], cfg)
    mf_aug = 0
    for k in out["att_i"]:
        t, _ = bpr_head(out["prof_u"][users], out["att_i"][k][pos], out["att_i"][k][neg], cfg)
        mf_aug = mf_aug + t
    sq = lambda x: 0.5 * (x ** 2).sum()
    feat_reg = (sq(out["img_i"]) + sq(out["txt_i"]) + sq(out["img_u"]) + sq(out["txt_u"])) / n_items
    feat = cfg.feat_reg_decay * feat_reg
    total = mf + emb + 0.0 + feat + cfg.aug_mf_rate * mf_aug + cfg.mm_mf_rate * (mf_img + mf_txt)
    return total, dict(mf=mf, emb=emb, feat=feat, mf_aug=mf_aug, mf_img=mf_img, mf_txt=mf_txt)


# ----------------------------------------------------------------------------------------
# sampling (utility/load_data.py:157-195 + main.py:216-224).  Uses the GLOBAL python `random`
# and `np.random` generators exactly like the reference, so seeding them reproduces its batches.
# ----------------------------------------------------------------------------------------
def sample_batch(data: OracleData, cfg: OracleConfig):
    if cfg.batch_size <= data.n_users:
        users = _pyrandom.sample(data.exist_users, cfg.batch_size)
    else:
        users = [_pyrandom.choice(data.exist_users) for _ in range(cfg.batch_size)]
    pos, neg = [], []
    for u in users:
        mine = data.train_items[u]
        pos.append(mine[np.random.randint(low=0, high=len(mine), size=1)[0]])
        while True:
            c = np.random.randint(low=0, high=data.n_items, size=1)[0]
            if c not in mine:
                neg.append(c)
                break
    ni = data.n_items
    picked = _pyrandom.sample(users, int(len(users) * cfg.aug_sample_rate))
    ok = [u for u in picked if data.aug_samples[u][0] < ni and data.aug_samples[u][1] < ni]
    users = users + ok
    pos = pos + [data.aug_samples[u][0] for u in ok]
    neg = neg + [data.aug_samples[u][1] for u in ok]
    return users, pos, neg


# ----------------------------------------------------------------------------------------
# evaluation (utility/batch_test.py:21-36,70-109,112-169; utility/metrics.py)
# ----------------------------------------------------------------------------------------
def dcg(r, k):
    r = np.asarray(r, dtype=np.float64)[:k]
    if r.size:
        return np.sum(r / np.log2(np.arange(2, r.size + 2)))
    return 0.0


def user_metrics(r, n_pos, Ks):
    """precision/recall/ndcg/hit at each K from the hit vector of the top-max(Ks) list.
    NDCG's ideal ordering is built from the hits INSIDE the retrieved list (metrics.py:68-78)."""
    prec, rec, nd, hit = [], [], [], []
    ideal = sorted(r, reverse=True)
    for K in Ks:
        head = np.asarray(r)[:K]
        prec.append(np.mean(head))
        rec.append(np.sum(np.asarray(r, dtype=np.float64)[:K]) / n_pos if n_pos else 0)
        best = dcg(ideal, K)
        nd.append(dcg(r, K) / best if best else 0.0)
        hit.append(1.0 if np.sum(head) > 0 else 0.0)
    return dict(precision=np.array(prec), recall=np.array(rec), ndcg=np.array(nd), hit_ratio=np.array(hit))


def rank_user_heapq(scores_row, train_items, n_items, kmax):
    """Top-kmax item ids among items not in train_items; ties -> lowest id (set-difference
    iterates ascending, heapq.nlargest is stable; batch_test.py:21-36,100-102)."""
    cand = list(set(range(n_items)) - set(train_items))
    table = {i: scores_row[i] for i in cand}
    return heapq.nlargest(kmax, table, key=table.get)


def rank_users_numpy(scores, train_lists, kmax):
    """Vectorised equivalent of rank_user_heapq for a block of users (checker for large cases)."""
    s = np.array(scores, dtype=np.float32, copy=True)
    for r, items in enumerate(train_lists):
        if len(items):
            s[r, np.asarray(items, dtype=np.int64)] = -np.inf
    n = s.shape[1]
    ids = np.broadcast_to(np.arange(n), s.shape)
    order = np.lexsort((ids, -s), axis=1)[:, :kmax]
    return order


def user_auc(scores_row, train_items, pos_items, n_items):
    """test_flag='full' (batch_test.py:38-54, metrics.py:95-100): roc_auc_score of the scores of all candidates (items not in
    train_items), label 1 for the user's truth items; 0. when sklearn refuses (a single class)."""
    from sklearn.metrics import roc_auc_score
    cand = sorted(set(range(n_items)) - set(train_items))
    pos = set(pos_items)
    y = [1 if i in pos else 0 for i in cand]
    try:
        return float(roc_auc_score(y_true=y, y_score=[scores_row[i] for i in cand]))
    except Exception:
        return 0.0


def evaluate(U, I, data: OracleData, users, cfg: OracleConfig, is_val=False, faithful=True, full=False):
    """test_torch: blocks of 2*batch_size users, fp32 scores, per-user ranking, metrics averaged
    over n_test_users by sequential float64 accumulation (batch_test.py:112-169)."""
    Ks = list(cfg.Ks)
    kmax = max(Ks)
    res = {k: np.zeros(len(Ks)) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    res["auc"] = 0.0
    truth = data.val_set if is_val else data.test_set
    step = cfg.batch_size * 2
    n = len(users)
    tops = {}
    for s in range(0, n, step):
        blk = users[s:s + step]
        rate = torch.matmul(U[blk], I.t()).detach().cpu().numpy()                # batch_test.py:150-154
        if faithful:
            lists = [rank_user_heapq(rate[j], data.train_items.get(u, []), data.n_items, kmax) for j, u in enumerate(blk)]
        else:
            lists = rank_users_numpy(rate, [data.train_items.get(u, []) for u in blk], kmax).tolist()
        for j, (u, top) in enumerate(zip(blk, lists)):
            pos = truth[u]
            r = [1 if i in pos else 0 for i in top]
            m = user_metrics(r, len(pos), Ks)
            for k in ("precision", "recall", "ndcg", "hit_ratio"):
                res[k] += m[k] / n
            if full:
                res["auc"] += user_auc(rate[j], data.train_items.get(u, []), pos, data.n_items) / n
            tops[u] = top
    return res, tops


# ----------------------------------------------------------------------------------------
# trainer (main.py:37-110,199-283)
# ----------------------------------------------------------------------------------------
def set_seed(seed):
    np.random.seed(seed); _pyrandom.seed(seed); torch.manual_seed(seed)


class OracleTrainer:
    def __init__(self, data: OracleData, cfg: OracleConfig, device="cpu"):
        """device="cuda" runs the SAME torch ops through cuSPARSE / cuBLAS / ATen (the "reference torch.sparse on the same
        B200" baseline of BASELINE.json configs[1]; the reference builds on CPU and calls .cuda(), main.py:91,96)."""
        self.data, self.cfg = data, cfg
        dev = self.device = torch.device(device)
        self.ui, self.iu = (g.to(dev) for g in build_graphs(data.train_mat))
        self.params = {k: v.detach().to(dev).requires_grad_(True) for k, v in init_params(cfg, data).items()}
        self.feats = dict(image=torch.tensor(data.image_feats).float().to(dev), text=torch.tensor(data.text_feats).float().to(dev),
                          user=torch.tensor(data.user_feats).float().to(dev),
                          item={k: torch.tensor(v).float().to(dev) for k, v in data.item_feats.items()})
        # torch default AdamW: betas (0.9, 0.999), eps 1e-8, weight_decay 0.01 (main.py:100-104)
        self.opt = torch.optim.AdamW(list(self.params.values()), lr=cfg.lr)

    def forward(self):
        return forward(self.params, self.feats, self.ui, self.iu, self.cfg)

    def step(self, users, pos, neg):
        out = self.forward()
        total, parts = batch_loss(out, users, pos, neg, self.data.n_items, self.cfg)
        self.opt.zero_grad()
        total.backward()
        self.opt.step()
        return float(total), {k: float(v) for k, v in parts.items()}

    def train_epoch(self):
        n_batch = self.data.n_train // self.cfg.batch_size + 1           # main.py:203
        tot = mf = emb = 0.0
        n_inter = 0
        for _ in range(n_batch):
            u, p, n = sample_batch(self.data, self.cfg)
            l, parts = self.step(u, p, n)
            tot += l; mf += parts["mf"]; emb += parts["emb"]
            n_inter += len(u)
        return dict(loss=tot, mf_loss=mf, emb_loss=emb, n_batch=n_batch, triplets=n_inter)

    def test(self, users=None, is_val=False, faithful=True, full=False):
        with torch.no_grad():
            out = self.forward()
        users = list(self.data.test_set.keys()) if users is None else users
        return evaluate(out["U"], out["I"], self.data, users, self.cfg, is_val, faithful, full)
