"""engine.HotPath's ORCHESTRATION on CPU: the hand-scheduled forward / loss heads / backward chain / AdamW of the full
side-feature model runs with torch stand-ins for the CUDA kernels (tests/ops_emulator.py, installed in a child process) and
must track the CPU oracle (autograd + torch.optim.AdamW, pinned to the reference's golden vectors) step for step.
The kernels themselves are covered by the -m gpu tests; this guards everything between them."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, ddir, out):
    sys.path.insert(0, HERE); sys.path.insert(0, REPO)
    torch.set_num_threads(2)
    import ops_emulator
    ops_emulator.install()
    from llmrec_b200.engine import HotPath, HotPathConfig, PARAM_ORDER
    from llmrec_b200.graph import BipartiteGraph
    from oracle import llmrec_oracle as O
    data = O.load_dataset(ddir)
    ok = True
    for weight_size, d, split in (("[64, 64]", 64, False), ("[32,32,32]", 32, False), ("[64, 64]", 64, True), ("[32,32,32]", 32, True)):
        ocfg = O.OracleConfig(batch_size=128, embed_size=d, weight_size=eval(weight_size), lr=1e-3)
        O.set_seed(2022)
        otr = O.OracleTrainer(data, ocfg)
        params = {k: otr.params[k].detach().clone() for k in PARAM_ORDER}
        feats = dict(image=otr.feats["image"].clone(), text=otr.feats["text"].clone(), user=otr.feats["user"].clone(),
                     item={k: v.clone() for k, v in otr.feats["item"].items()})
        g = BipartiteGraph(data.train_mat, "cpu")
        cfg = HotPathConfig(embed_size=d, n_layers=len(eval(weight_size)), batch_size=128)
        hp = HotPath((g.ui, g.iu, g.uiT, g.iuT), params, feats, cfg)
        hp.set_optimizer(lr=1e-3)
        hp.force_split = split              # the branch schedule of train_step (ID layers | user-profile operand | side features), run in line
        O.set_seed(7)
        for step in range(3):
            users, pos, neg = O.sample_batch(data, ocfg)
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
    out[0] = bool(ok)


def test_engine_orchestration_tracks_the_oracle(tiny_root):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(os.path.join(tiny_root, "netflix_valid_item"), out), nprocs=1, join=True)
    assert dict(out) == {0: True}
