"""hoist.HoistedHotPath's algebra and orchestration on CPU: with torch stand-ins for the kernels
(tests/ops_emulator.py) the hoisted engine -- propagated tables precomputed once, per-step work on the batch's rows only, feat_reg
through Gram matrices -- must track the CPU oracle (autograd through the full dense model) step for step, exactly like the
default engine does in test_engine_emulated.py.  Both the eager form (varying B') and the capacity form (`meta` = {B', n_keep},
what the CUDA graph replays) are exercised."""
import os
import sys

import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, ddir, out):
    sys.path.insert(0, HERE); sys.path.insert(0, REPO)
    torch.set_num_threads(2)
    import ops_emulator
    ops_emulator.install()
    from llmrec_b200.engine import HotPathConfig, PARAM_ORDER
    from llmrec_b200.graph import BipartiteGraph
    from llmrec_b200.hoist import HoistedHotPath
    from oracle import llmrec_oracle as O
    data = O.load_dataset(ddir)
    ok = True
    for weight_size, d, capacity in (("[64, 64]", 64, False), ("[32,32,32]", 32, True)):
        ocfg = O.OracleConfig(batch_size=128, embed_size=d, weight_size=eval(weight_size), lr=1e-3)
        O.set_seed(2022)
        otr = O.OracleTrainer(data, ocfg)
        params = {k: otr.params[k].detach().clone() for k in PARAM_ORDER}
        feats = dict(image=otr.feats["image"].clone(), text=otr.feats["text"].clone(), user=otr.feats["user"].clone(),
                     item={k: v.clone() for k, v in otr.feats["item"].items()})
        g = BipartiteGraph(data.train_mat, "cpu")
        cfg = HotPathConfig(embed_size=d, n_layers=len(eval(weight_size)), batch_size=128)
        hp = HoistedHotPath((g.ui, g.iu, g.uiT, g.iuT), params, feats, cfg, g.ones_propagated())
        hp.set_optimizer(lr=1e-3)
        O.set_seed(7)
        for step in range(3):
            users, pos, neg = O.sample_batch(data, ocfg)
            B = len(users)
            if capacity:                                  # the graph path: capacity-sized index rows + device-side {B', n_keep}
                gi = hp.index_buffer(B)
                gi.zero_()
                gi[0, :B] = torch.tensor(users, dtype=torch.int32); gi[1, :B] = torch.tensor(pos, dtype=torch.int32); gi[2, :B] = torch.tensor(neg, dtype=torch.int32)
                gi[3, 0], gi[3, 1] = hp.meta_row(B)
                got = float(hp.train_step(gi[0], gi[1], gi[2], gi[3]))
            else:
                t = lambda x: torch.tensor(x, dtype=torch.int32)
                got = float(hp.train_step(t(users), t(pos), t(neg)))
            want, _ = otr.step(users, pos, neg)
            ok &= abs(got - want) < 2e-5 * max(1.0, abs(want))
        for k in PARAM_ORDER:
            ok &= bool(torch.allclose(params[k], otr.params[k].detach(), rtol=2e-4, atol=2e-6))
        U, I = hp.forward()
        with torch.no_grad():
            o = otr.forward()
        ok &= bool(torch.allclose(U, o["U"], rtol=1e-4, atol=1e-6)) and bool(torch.allclose(I, o["I"], rtol=1e-4, atol=1e-6))
        v = hp.side_views()
        ok &= bool(torch.allclose(v["img_i"], o["img_i"], rtol=1e-4, atol=1e-6)) and bool(torch.allclose(v["prof_u"], o["prof_u"], rtol=1e-4, atol=1e-6))
    out[0] = bool(ok)


def test_hoisted_engine_tracks_the_oracle(tiny_root):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(os.path.join(tiny_root, "netflix_valid_item"), out), nprocs=1, join=True)
    assert dict(out) == {0: True}
