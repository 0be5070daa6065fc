"""Generate golden vectors by running the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py

Writes tests/golden/tiny_ref.npz: outputs of /root/reference's own MM_Model / Trainer /
test_torch on the seeded tiny synthetic dataset (llmrec_b200.synth, seed 0, 300 users x 400
items, feature dims 32/64/96, d=64, L=2, batch 128, --seed 2022).  Inputs are NOT stored:
tests regenerate them from the same seeds.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

TINY = dict(dataset="netflix", n_users=300, n_items=400, n_inter=1500, dims=(32, 64, 96), seed=0)
TINY_FLAGS = ["--batch_size", "128", "--epoch", "1", "--debug", "--seed", "2022"]


def main():
    import torch
    from llmrec_b200.synth import make_dataset
    from oracle.ref_shim import import_reference

    root = tempfile.mkdtemp(prefix="llmrec_golden_") + "/"
    make_dataset(root, **TINY)
    torch.set_num_threads(1)
    ref = import_reference(["--data_path", root, "--dataset", "netflix_valid_item"] + TINY_FLAGS)
    ref.set_seed(2022)
    tr = ref.Trainer(data_config={})
    out = {}
    for k, v in tr.model_mm.state_dict().items():
        if k.startswith("batch_norm"):
            continue
        a0 = v.detach().numpy()
        out["init/" + k] = a0[:8].copy()                       # head rows + checksum pin the RNG order
        out["initsum/" + k] = np.float64(a0.astype(np.float64).sum())

    # forward at init
    tr.model_mm.eval()
    with torch.no_grad():
        tup = tr.model_mm(tr.ui_graph, tr.iu_graph, tr.image_ui_graph, tr.image_iu_graph, tr.text_ui_graph, tr.text_iu_graph)
    names = ["U", "I", "img_i", "txt_i", "img_u", "txt_u", "p_usr", None, "prof_u", "prof_i"]
    for n, t in zip(names, tup):
        if n:
            out["fwd/" + n] = t.numpy().copy() if n in ("U", "I") else t.numpy()[:32].copy()
    for k, t in tup[10].items():
        out["fwd/att_u/" + k] = t.numpy()[:32].copy()
    for k, t in tup[11].items():
        out["fwd/att_i/" + k] = t.numpy()[:32].copy()

    # one fixed batch: loss pieces + gradients through the reference's own bpr/prune code
    ref.set_seed(7)
    users, pos, neg = ref.data_generator.sample()
    out["batch/users"], out["batch/pos"], out["batch/neg"] = map(np.asarray, (users, pos, neg))
    tr.model_mm.train()
    tup = tr.model_mm(tr.ui_graph, tr.iu_graph, tr.image_ui_graph, tr.image_iu_graph, tr.text_ui_graph, tr.text_iu_graph)
    U, I, img_i, txt_i, img_u, txt_u, _, _, prof_u, prof_i, att_u, att_i, _, _ = tup
    mf, emb, _ = tr.bpr_loss(U[users], I[pos], I[neg])
    mf_img, _, _ = tr.bpr_loss(img_u[users], img_i[pos], img_i[neg])
    mf_txt, _, _ = tr.bpr_loss(txt_u[users], txt_i[pos], txt_i[neg])
    mf_aug = 0
    for k in att_i:
        t, _, _ = tr.bpr_loss(prof_u[users], att_i[k][pos], att_i[k][neg])
        mf_aug = mf_aug + t
    feat = tr.feat_reg_loss_calculation(img_i, txt_i, img_u, txt_u)
    a = ref.args
    total = mf + emb + 0.0 + feat + a.aug_mf_rate * mf_aug + a.mm_mf_rate * (mf_img + mf_txt)
    tr.optimizer.zero_grad()
    total.backward()
    out["loss/parts"] = np.array([float(total), float(mf), float(emb), float(feat), float(mf_aug), float(mf_img), float(mf_txt)])
    for k, p in tr.model_mm.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy().copy()
    tr.optimizer.zero_grad()

    # one full epoch through the reference's own train() (sampling, aug edges, AdamW, eval)
    logs = []
    tr.logger.logging = lambda s: logs.append(str(s))
    ref.set_seed(2022)
    tr2 = ref.Trainer(data_config={})
    tr2.logger.logging = lambda s: logs.append(str(s))
    tr2.train()
    for k, v in tr2.model_mm.state_dict().items():
        if k.startswith("batch_norm"):
            continue
        out["epoch1/" + k] = v.detach().numpy().copy()
    line = [s for s in logs if s.startswith("Epoch 0 [")][0]
    out["epoch1/log"] = np.array(line)
    res = tr2.test(list(ref.data_generator.test_set.keys()), is_val=False)
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        out["epoch1/metric/" + k] = np.asarray(res[k], dtype=np.float64)

    # hit vectors of the reference's own ranker for every test user at the trained weights
    tr2.model_mm.eval()
    with torch.no_grad():
        ua, ia, *_ = tr2.model_mm(tr2.ui_graph, tr2.iu_graph, tr2.image_ui_graph, tr2.image_iu_graph, tr2.text_ui_graph, tr2.text_iu_graph)
    rate = torch.matmul(ua, ia.t()).numpy()
    dg = ref.data_generator
    hits = []
    for u in sorted(dg.test_set.keys()):
        cand = list(set(range(dg.n_items)) - set(dg.train_items[u]))
        r, _ = ref.ranklist_by_heapq(dg.test_set[u], cand, rate[u], ref.Ks)
        hits.append(r)
    out["epoch1/hits"] = np.asarray(hits, dtype=np.uint8)

    # sampler stream: first 3 batches after set_seed(2022), incl. augmented edges (main.py:216-224)
    import pickle, random
    ref.set_seed(2022)
    aug = pickle.load(open(os.path.join(root, "netflix_valid_item", "augmented_sample_dict"), "rb"))
    for b in range(3):
        u, p, n = dg.sample()
        ua_ = random.sample(u, int(len(u) * a.aug_sample_rate))
        ok = [x for x in ua_ if aug[x][0] < dg.n_items and aug[x][1] < dg.n_items]
        out[f"sampler/{b}"] = np.asarray([u + ok, p + [aug[x][0] for x in ok], n + [aug[x][1] for x in ok]])

    # metric known-answers straight from the reference's metrics module
    m = ref.metrics
    r1 = [0, 1, 0, 0, 1] + [0] * 45
    out["ka/ndcg10"] = np.float64(m.ndcg_at_k(r1, 10))
    out["ka/recall10"] = np.float64(m.recall_at_k(r1, 10, 3))
    r2 = [0] * 15 + [1] + [0] * 34
    out["ka/ndcg20_rank16"] = np.float64(m.ndcg_at_k(r2, 20))

    np.savez_compressed(os.path.join(HERE, "tiny_ref.npz"), **out)
    print("wrote", os.path.join(HERE, "tiny_ref.npz"), {k: getattr(v, "shape", ()) for k, v in list(out.items())[:6]})
    print(line)


if __name__ == "__main__":
    main()
