"""-m gpu: the assembled hot path (MM_Model / Trainer / test_torch) against the golden vectors of the
unmodified reference and against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY_FLAGS = ["--batch_size", "128", "--epoch", "1", "--debug", "--seed", "2022"]


def _trainer(root, extra=()):
    from llmrec_b200 import main as M
    from llmrec_b200.runtime import set_args
    from llmrec_b200.utility import batch_test
    from llmrec_b200.utility.load_data import Data
    from llmrec_b200.utility.parser import parse_args, resolve_dataset_dir
    args = set_args(parse_args(["--data_path", root, "--dataset", "netflix"] + TINY_FLAGS + list(extra)))
    M.set_seed(args.seed)
    gen = Data(path=resolve_dataset_dir(args.data_path, args.dataset), batch_size=args.batch_size, sampler=args.host_sampler)
    batch_test.init(gen, args)
    return M.Trainer(data_config={}, data_generator=gen), gen, M


@pytest.mark.parametrize("mode", ["3xtf32", "fp32"])
def test_forward_matches_reference_golden(tiny_root, golden, mode):
    tr, gen, M = _trainer(tiny_root, ["--proj_mode", mode])
    sd = tr.model_mm.state_dict()
    for k in ("image_trans.weight", "user_id_embedding.weight", "item_trans.bias"):
        np.testing.assert_array_equal(sd[k].cpu().numpy()[:8], golden["init/" + k])          # same RNG order as the reference
    out = tr.model_mm(tr.ui_graph, tr.iu_graph, tr.image_ui_graph, tr.image_iu_graph, tr.text_ui_graph, tr.text_iu_graph)
    assert len(out) == 14
    U, I, img_i, txt_i, img_u, txt_u, p_usr, att_i, prof_u, prof_i, att_u, att_i2, _, _ = out
    tol = dict(rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(U.detach().cpu().numpy(), golden["fwd/U"], **tol)
    np.testing.assert_allclose(I.detach().cpu().numpy(), golden["fwd/I"], **tol)
    for name, t in (("img_i", img_i), ("txt_i", txt_i), ("img_u", img_u), ("txt_u", txt_u), ("p_usr", p_usr), ("prof_u", prof_u), ("prof_i", prof_i)):
        np.testing.assert_allclose(t.detach().cpu().numpy()[:32], golden["fwd/" + name], **tol)
    for k in att_i:
        np.testing.assert_allclose(att_u[k].detach().cpu().numpy()[:32], golden["fwd/att_u/" + k], **tol)
        np.testing.assert_allclose(att_i[k].detach().cpu().numpy()[:32], golden["fwd/att_i/" + k], **tol)


def test_autograd_through_model_matches_reference_grads(tiny_root, golden):
    """loss assembled with the Trainer's torch-level bpr_loss API; gradients via the custom backward."""
    tr, gen, M = _trainer(tiny_root, ["--proj_mode", "fp32"])
    users, pos, neg = (golden["batch/" + k].tolist() for k in ("users", "pos", "neg"))
    out = tr.model_mm(tr.ui_graph, tr.iu_graph)
    U, I, img_i, txt_i, img_u, txt_u, _, att_i, prof_u, prof_i, att_u, _, _, _ = out
    a = tr.args
    mf, emb, _ = tr.bpr_loss(U[users], I[pos], I[neg])
    mf_i, _, _ = tr.bpr_loss(img_u[users], img_i[pos], img_i[neg])
    mf_t, _, _ = tr.bpr_loss(txt_u[users], txt_i[pos], txt_i[neg])
    aug = 0
    for k in att_i:
        t, _, _ = tr.bpr_loss(prof_u[users], att_i[k][pos], att_i[k][neg])
        aug = aug + t
    feat = tr.feat_reg_loss_calculation(img_i, txt_i, img_u, txt_u)
    total = mf + emb + feat + a.aug_mf_rate * aug + a.mm_mf_rate * (mf_i + mf_t)
    total.backward()
    got = np.array([float(total), float(mf), float(emb), float(feat), float(aug), float(mf_i), float(mf_t)])
    np.testing.assert_allclose(got, golden["loss/parts"], rtol=2e-5, atol=1e-9)
    for name, p in tr.model_mm.named_parameters():
        if name.startswith("batch_norm"):
            continue
        ref = golden["grad/" + name]
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref, rtol=2e-3, atol=1e-9 + 2e-5 * np.abs(ref).max(), err_msg=name)


@pytest.mark.parametrize("mode", ["3xtf32", "fp32"])
def test_fused_step_grads_match_reference(tiny_root, golden, mode):
    tr, gen, M = _trainer(tiny_root, ["--proj_mode", mode])
    users, pos, neg = (golden["batch/" + k].tolist() for k in ("users", "pos", "neg"))
    u, p, n = tr.upload_batch(users, pos, neg)
    hp = tr.hot
    hp.forward()
    hp.loss_and_output_grads(u, p, n)
    grads = hp.backward()
    parts = golden["loss/parts"]
    assert abs(float(hp.loss) - parts[0]) < 2e-5 * abs(parts[0])
    ho = hp.head_out.cpu().view(-1, 4)
    assert abs(float(ho[0, 0]) - parts[1]) < 1e-5 and abs(float(ho[0, 1]) - parts[2]) < 1e-9
    for name, gten in grads.items():
        ref = golden["grad/" + name]
        np.testing.assert_allclose(gten.cpu().numpy(), ref, rtol=2e-3, atol=1e-9 + 2e-5 * np.abs(ref).max(), err_msg=name)


@pytest.mark.parametrize("mode,graph,hoist", [("3xtf32", 1, 0), ("fp32", 1, 0), ("3xtf32", 0, 0), ("3xtf32", 1, 1), ("3xtf32", 0, 1)])
def test_epoch_matches_reference_golden(tiny_root, golden, mode, graph, hoist):
    """Same seed, same sampler stream, 8 steps of AdamW, then full-catalog eval: params, epoch loss, metrics, hit vectors.
    graph=1 replays the step from a CUDA graph (the default), graph=0 launches eagerly; hoist=1 is the hoisted side-feature
    engine (hoist.py), held to the SAME golden tolerances as the default engine."""
    tr, gen, M = _trainer(tiny_root, ["--proj_mode", mode, "--cuda_graph", str(graph), "--hoist_side", str(hoist)])
    assert tr.hoisted == bool(hoist)
    M.set_seed(2022)
    logs = []
    tr.logger.logging = lambda s: logs.append(str(s))
    tr.train()
    line = [s for s in logs if s.startswith("Epoch 0 [")][0]
    ref_line = str(golden["epoch1/log"])
    loss, mf = (float(x) for x in line.split("train==[")[1].split("+")[0].split("="))
    rloss, rmf = (float(x) for x in ref_line.split("train==[")[1].split("+")[0].split("="))
    assert abs(loss - rloss) < 5e-4 and abs(mf - rmf) < 5e-4, (line, ref_line)
    sd = tr.model_mm.state_dict()
    for k in sd:
        if k.startswith("batch_norm"):
            continue
        np.testing.assert_allclose(sd[k].cpu().numpy(), golden["epoch1/" + k], rtol=2e-4, atol=2e-6, err_msg=k)
    res = tr.test(list(gen.test_set.keys()), is_val=False)
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        np.testing.assert_allclose(res[k], golden["epoch1/metric/" + k], rtol=0, atol=1e-4)     # north_star tolerance
    from llmrec_b200.utility import batch_test
    ua, ia = tr.hot.forward()
    users = sorted(gen.test_set.keys())
    _, hits = batch_test.rank_block(ua, ia, users, False)
    mism = int((hits.cpu().numpy() != golden["epoch1/hits"]).sum())
    assert mism <= 2, mism


def test_sampler_stream_matches_reference(tiny_root, golden):
    tr, gen, M = _trainer(tiny_root)
    M.set_seed(2022)
    for b in range(3):
        u, p, n = tr.sample_batch()
        np.testing.assert_array_equal(np.asarray([u, p, n]), golden[f"sampler/{b}"])


def test_no_feature_mode_and_layers(tiny_root):
    """ID-only propagation (the large synthetic configuration) vs the oracle's ID chain, L = 1 and 3."""
    from llmrec_b200.engine import HotPath, HotPathConfig
    from llmrec_b200.graph import BipartiteGraph
    from oracle import llmrec_oracle as O
    data = O.load_dataset(os.path.join(tiny_root, "netflix_valid_item"))
    for L in (1, 3):
        torch.manual_seed(L)
        Eu, Ei = torch.randn(data.n_users, 64) * 0.1, torch.randn(data.n_items, 64) * 0.1
        g = BipartiteGraph(data.train_mat, "cuda")
        params = {"user_id_embedding.weight": Eu.cuda(), "item_id_embedding.weight": Ei.cuda()}
        hp = HotPath((g.ui, g.iu, g.uiT, g.iuT), params, None, HotPathConfig(embed_size=64, n_layers=L, batch_size=128))
        U, I = hp.forward()
        ui, iu = O.build_graphs(data.train_mat)
        eu, ei = Eu.clone().requires_grad_(True), Ei.clone().requires_grad_(True)
        us, its, a, b = [eu], [ei], eu, ei
        for l in range(L):
            a = torch.mm(ui, b); a = torch.softmax(a, -1) if l == L - 1 else a
            b = torch.mm(iu, a); b = torch.softmax(b, -1) if l == L - 1 else b
            us.append(a); its.append(b)
        Ur, Ir = torch.mean(torch.stack(us), 0), torch.mean(torch.stack(its), 0)
        torch.testing.assert_close(U.cpu(), Ur.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(I.cpu(), Ir.detach(), rtol=1e-5, atol=1e-6)
        rng = np.random.default_rng(L)
        users = torch.from_numpy(rng.integers(0, data.n_users, 140)); pos = torch.from_numpy(rng.integers(0, data.n_items, 140)); neg = torch.from_numpy(rng.integers(0, data.n_items, 140))
        cfg = O.OracleConfig(batch_size=128)
        mf, emb = O.bpr_head(Ur[users], Ir[pos], Ir[neg], cfg)
        (mf + emb).backward()
        i32 = lambda t: t.to(torch.int32).cuda()
        hp.loss_and_output_grads(i32(users), i32(pos), i32(neg))
        grads = hp.backward()
        assert abs(float(hp.loss) - float(mf + emb)) < 1e-5
        torch.testing.assert_close(grads["user_id_embedding.weight"].cpu(), eu.grad, rtol=1e-3, atol=1e-8)
        torch.testing.assert_close(grads["item_id_embedding.weight"].cpu(), ei.grad, rtol=1e-3, atol=1e-8)


def test_movielens_config_d128_L3_vs_oracle(tmp_path):
    """BASELINE.json configs[2]: movielens keys, d=128, 3 ID layers, image/text/attribute projections + fusion:
    forward, loss and one fused AdamW step against the CPU oracle on the same seeded inputs."""
    from llmrec_b200 import main as M
    from llmrec_b200.runtime import set_args
    from llmrec_b200.synth import make_dataset
    from llmrec_b200.utility import batch_test
    from llmrec_b200.utility.load_data import Data
    from llmrec_b200.utility.parser import parse_args, resolve_dataset_dir
    from oracle import llmrec_oracle as O
    root = str(tmp_path) + "/"
    make_dataset(root, dataset="movielens", n_users=260, n_items=330, n_inter=1300, dims=(64, 96, 128), seed=3)
    args = set_args(parse_args(["--data_path", root, "--dataset", "movielens", "--batch_size", "128", "--debug", "--embed_size", "128",
                                "--weight_size", "[128,128,128]", "--lr", "0.001"]))
    ddir = resolve_dataset_dir(args.data_path, args.dataset)
    M.set_seed(11)
    gen = Data(path=ddir, batch_size=128)
    batch_test.init(gen, args)
    tr = M.Trainer(data_config={}, data_generator=gen)
    torch.set_num_threads(2)
    data = O.load_dataset(ddir)
    O.set_seed(11)
    cfg = O.OracleConfig(embed_size=128, weight_size=(128, 128, 128), batch_size=128, lr=1e-3)
    otr = O.OracleTrainer(data, cfg)
    for k in O.PARAM_NAMES:
        np.testing.assert_array_equal(tr.model_mm.state_dict()[k].cpu().numpy(), otr.params[k].detach().numpy())
    U, I = tr.hot.forward()
    with torch.no_grad():
        out = otr.forward()
    np.testing.assert_allclose(U.cpu().numpy(), out["U"].numpy(), rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(I.cpu().numpy(), out["I"].numpy(), rtol=3e-5, atol=3e-6)
    M.set_seed(5)
    users, pos, neg = tr.sample_batch()
    loss = float(tr.train_batch(users, pos, neg))
    oloss, _ = otr.step(users, pos, neg)
    assert abs(loss - oloss) < 3e-5 * max(1.0, abs(oloss)), (loss, oloss)
    sd = tr.model_mm.state_dict()
    for k in O.PARAM_NAMES:
        np.testing.assert_allclose(sd[k].cpu().numpy(), otr.params[k].detach().numpy(), rtol=5e-3, atol=2e-5, err_msg=k)
    res = tr.test(list(gen.test_set.keys()), is_val=False)
    ores, _ = otr.test()
    for k in ("recall", "ndcg", "precision", "hit_ratio"):
        np.testing.assert_allclose(res[k], ores[k], atol=1e-4)


def test_graphed_steps_with_varying_batch_length_match_eager(tiny_root):
    """ONE captured graph serves every batch length (B' and n_keep come from the device): alternate two lengths over many steps,
    allocate and free memory in between (a per-B' scratch buffer freed under a live graph was a use-after-free) and
    hold the result to the eager path on identical batches."""
    trg, gen, M = _trainer(tiny_root, ["--cuda_graph", "1"])
    tre, _, _ = _trainer(tiny_root, ["--cuda_graph", "0"])
    M.set_seed(11)
    batches = [trg.sample_batch() for _ in range(14)]
    junk = []
    for i, (u, p, n) in enumerate(batches):
        B = len(u) if i % 2 == 0 else len(u) - 9 - (i % 5)
        u, p, n = u[:B], p[:B], n[:B]
        lg = float(trg.train_batch(u, p, n))
        le = float(tre.train_batch(u, p, n))
        assert abs(lg - le) <= 2e-5 * max(1.0, abs(le)), (i, B, lg, le)
        junk.append(torch.full((1 << (16 + i % 4),), float(i), device="cuda"))      # churn the caching allocator between replays
        if i % 3 == 2:
            junk.clear()
    assert trg.hot._graph is not None                                               # a single graph, whatever the batch length
    sg, se = trg.model_mm.state_dict(), tre.model_mm.state_dict()
    for k in sg:
        if not k.startswith("batch_norm"):
            torch.testing.assert_close(sg[k], se[k], rtol=2e-5, atol=2e-7)
    for t in junk:
        assert float(t[0]) == float(t[-1])                                          # nothing scribbled over live memory


def test_hoisting_switches_itself_off_when_dropout_or_mask_is_on(tiny_root):
    """The linearity argument of hoist.py needs Dropout to be the identity: --drop_rate > 0 (or the mask branch) falls back to the default engine."""
    tr, gen, M = _trainer(tiny_root, ["--hoist_side", "1", "--drop_rate", "0.2"])
    assert tr.hoisted is False and type(tr.hot).__name__ == "HotPath"


def test_hoisted_step_matches_default_engine_on_the_same_batches(tiny_root):
    """Default vs hoisted engine, same init, same batches, 6 steps: losses, every parameter, and the eval forward agree to fp32
    reassociation level (tighter than the golden tolerances)."""
    a, gen, M = _trainer(tiny_root, ["--cuda_graph", "0"])
    b, _, _ = _trainer(tiny_root, ["--cuda_graph", "1", "--hoist_side", "1"])
    M.set_seed(5)
    for i in range(6):
        u, p, n = a.sample_batch()
        la, lb = float(a.train_batch(u, p, n)), float(b.train_batch(u, p, n))
        assert abs(la - lb) <= 2e-5 * max(1.0, abs(la)), (i, la, lb)
    sa, sb = a.model_mm.state_dict(), b.model_mm.state_dict()
    for k in sa:
        if not k.startswith("batch_norm"):
            torch.testing.assert_close(sb[k], sa[k], rtol=1e-4, atol=1e-6, msg=k)
    Ua, Ia = a.hot.forward(); Ub, Ib = b.hot.forward()
    torch.testing.assert_close(Ub, Ua, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(Ib, Ia, rtol=1e-4, atol=1e-6)


def test_full_test_flag_auc_matches_sklearn_oracle(tiny_root):
    """--test_flag full (batch_test.py:38-68): same hit-based metrics plus the per-user ROC-AUC over every candidate, against the
    oracle's sklearn.roc_auc_score on the same embeddings."""
    from oracle import llmrec_oracle as O
    tr, gen, M = _trainer(tiny_root, ["--test_flag", "full"])
    M.set_seed(3)
    for _ in range(3):
        tr.train_batch(*tr.sample_batch())
    users = list(gen.test_set.keys())
    res = tr.test(users, is_val=False)
    data = O.load_dataset(gen.path if hasattr(gen, "path") else os.path.join(tiny_root, "netflix_valid_item"))
    U, I = tr.hot.U.cpu(), tr.hot.I.cpu()
    ores, _ = O.evaluate(U, I, data, users, O.OracleConfig(batch_size=128), is_val=False, faithful=True, full=True)
    assert res["auc"] > 0.3 and abs(res["auc"] - ores["auc"]) < 1e-5, (res["auc"], ores["auc"])
    for k in ("recall", "ndcg", "precision", "hit_ratio"):
        np.testing.assert_allclose(res[k], ores[k], atol=1e-4)


def test_mask_dropout_restoration_branch_matches_oracle(tiny_root):
    """--mask 1 --mask_rate 0.1 --drop_rate 0.2 --att_re_rate 0.5 (Models.py:131-150, main.py:258-271): the eager optional branch of
    Trainer against the oracle's autograd on the same state -- identical torch.randperm draws (CPU generator state copied), the
    dropout masks the GPU step drew, the same Decoder weights."""
    from oracle import llmrec_oracle as O
    flags = ["--mask", "1", "--mask_rate", "0.1", "--drop_rate", "0.2", "--att_re_rate", "0.5", "--cuda_graph", "1", "--hoist_side", "1"]
    tr, gen, M = _trainer(tiny_root, flags)
    assert tr.masked_mode and not tr.hoisted
    data = O.load_dataset(os.path.join(tiny_root, "netflix_valid_item"))
    cfg = O.OracleConfig(batch_size=128)
    otr = O.OracleTrainer(data, cfg)
    sd = tr.model_mm.state_dict()
    with torch.no_grad():
        for k in O.PARAM_NAMES:
            otr.params[k].copy_(sd[k].cpu())
    m = tr.model_mm
    otr.feats = dict(image=m.image_feats.cpu().clone(), text=m.text_feats.cpu().clone(), user=m.user_feats.cpu().clone(),
                     item={k: v.cpu().clone() for k, v in m.item_feats.items()})
    dec = dict(u_w=tr.decoder.u_net[0].weight.detach().cpu(), u_b=tr.decoder.u_net[0].bias.detach().cpu(),
               i_w=tr.decoder.i_net[0].weight.detach().cpu(), i_b=tr.decoder.i_net[0].bias.detach().cpu())
    raw_u = torch.tensor(tr.user_init_embedding).float()
    raw_i = {k: torch.tensor(v).float() for k, v in tr.item_attribute_embedding.items()}
    M.set_seed(9)
    for step in range(2):
        users, pos, neg = tr.sample_batch()
        state = torch.get_rng_state()
        got = float(tr.train_batch(users, pos, neg))
        torch.set_rng_state(state)                                    # the oracle draws the same permutations
        i_mask, u_mask = O.mask_features(otr.feats, data.n_users, data.n_items, True, 0.1)
        drop = [mk.cpu() for mk in tr._last_dropout_masks]
        assert 0.7 < float(drop[0].gt(0).float().mean()) < 0.9 and abs(float(drop[0].max()) - 1.25) < 1e-6
        out = O.forward(otr.params, otr.feats, otr.ui, otr.iu, cfg, drop=drop)
        total, _ = O.batch_loss(out, users, pos, neg, data.n_items, cfg)
        total = total + 0.5 * O.restoration_loss(out, dec, raw_u, raw_i, i_mask, u_mask, alpha=2, kind="sce")
        otr.opt.zero_grad(); total.backward(); otr.opt.step()
        assert abs(got - float(total)) <= 5e-5 * max(1.0, abs(float(total))), (step, got, float(total))
        torch.testing.assert_close(m.user_feats.cpu(), otr.feats["user"], rtol=1e-6, atol=1e-6)           # same rows were overwritten
    sd = tr.model_mm.state_dict()
    for k in O.PARAM_NAMES:
        np.testing.assert_allclose(sd[k].cpu().numpy(), otr.params[k].detach().numpy(), rtol=3e-4, atol=3e-6, err_msg=k)
    res = tr.test(list(gen.test_set.keys()), is_val=False)          # eval keeps masking (Models.py:131-142 is unconditional)
    assert np.isfinite(res["recall"]).all()


def test_device_sampler_batches_are_valid_and_reproducible(tiny_root):
    """--device_sampler 1: structural validity of the batches the GPU draws (what Data.sample() + main.py:216-224
    guarantee), reproducibility per seed, rough uniformity, and graph replay == eager launches on the same seed."""
    tr, gen, M = _trainer(tiny_root, ["--device_sampler", "1", "--cuda_graph", "0"])
    assert tr.device_sampler is not None
    hp, ds = tr.hot, tr.device_sampler
    B = gen.batch_size
    aug = tr.augmented_sample_dict
    seen, counts = [], np.zeros(gen.n_users)
    for step in range(40):
        hp.pre_step()
        gi = hp._gidx.cpu().numpy()
        Bp, n_keep = int(gi[3, 0]), int(gi[3, 1])
        assert B <= Bp <= B + int(B * 0.1) and n_keep == int((1 - 0.71) * Bp)
        u, p, n = gi[0, :Bp], gi[1, :Bp], gi[2, :Bp]
        assert len(set(u[:B].tolist())) == B and set(u[:B].tolist()) <= set(gen.exist_users)                 # a subset, no repeats
        for b in range(B):
            items = gen.train_items[int(u[b])]
            assert int(p[b]) in items and int(n[b]) not in items and 0 <= int(n[b]) < gen.n_items
        for b in range(B, Bp):                                                                              # augmented edges: batch users, table values, valid ids
            uu = int(u[b])
            assert uu in set(u[:B].tolist()) and (int(p[b]), int(n[b])) == (aug[uu][0], aug[uu][1]) and max(aug[uu][0], aug[uu][1]) < gen.n_items
        assert len(set(u[B:Bp].tolist())) == Bp - B
        seen.append(u[:B].copy()); counts[u[:B]] += 1
    assert not np.array_equal(np.sort(seen[0]), np.sort(seen[1]))
    expect = 40 * B / len(gen.exist_users)
    assert counts[gen.exist_users].min() >= 1 and abs(counts[gen.exist_users].mean() - expect) < 1e-6 and counts.max() < 3 * expect + 10
    # same seed -> same batches; graph replay == eager
    a, _, _ = _trainer(tiny_root, ["--device_sampler", "1", "--cuda_graph", "0"])
    b, _, _ = _trainer(tiny_root, ["--device_sampler", "1", "--cuda_graph", "1"])
    la = [float(a.train_next_batch()[0]) for _ in range(5)]
    lb = [float(b.train_next_batch()[0]) for _ in range(5)]
    assert all(abs(x - y) <= 2e-5 * max(1.0, abs(x)) for x, y in zip(la, lb)), (la, lb)
    assert np.isfinite(la).all()
    assert int(a._epoch_stats[3]) == int(b._epoch_stats[3]) >= 5 * B
