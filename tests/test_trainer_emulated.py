"""The whole Trainer flow on CPU against the REFERENCE's golden epoch: seeded init, the C batch sampler through Trainer.train(),
the engine's hand-scheduled step, eval + metrics -- with torch stand-ins for the CUDA kernels (tests/ops_emulator.py) and
for the three CUDA-only host facilities Trainer touches (pinned staging slots, events, synchronize).  Mirrors
tests/test_path_gpu.py::test_epoch_matches_reference_golden, which runs the same flow on the real kernels."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


class _Slot:
    def __init__(self, cap):
        self.host = torch.zeros((4, cap), dtype=torch.int32)
        self.np = self.host.numpy()
        self.event = SimpleNamespace(synchronize=lambda: None, record=lambda: None)


def _worker(rank, root, out):
    sys.path.insert(0, HERE); sys.path.insert(0, REPO)
    torch.set_num_threads(2)
    import ops_emulator
    ops_emulator.install()
    from llmrec_b200 import Models, main as M
    from llmrec_b200.runtime import set_args
    from llmrec_b200.utility import batch_test
    from llmrec_b200.utility.load_data import Data
    from llmrec_b200.utility.parser import parse_args, resolve_dataset_dir
    Models._on_device = lambda t: True
    M._StagingSlot = _Slot
    torch.cuda.is_available = lambda: True
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.manual_seed_all = lambda s: None
    golden = np.load(os.path.join(HERE, "golden", "tiny_ref.npz"), allow_pickle=False)
    args = set_args(parse_args(["--data_path", root, "--dataset", "netflix", "--batch_size", "128", "--epoch", "1", "--debug", "--seed", "2022",
                                "--cuda_graph", "0", "--proj_mode", "fp32"]))
    M.set_seed(args.seed)
    gen = Data(path=resolve_dataset_dir(args.data_path, args.dataset), batch_size=args.batch_size, sampler=args.host_sampler)
    batch_test.init(gen, args)
    tr = M.Trainer(data_config={}, data_generator=gen, device="cpu")
    assert tr._batch_sampler is not None                      # the C batch sampler is the one under test
    M.set_seed(2022)
    logs = []
    tr.logger.logging = lambda s: logs.append(str(s))
    tr.train()
    line = [s for s in logs if s.startswith("Epoch 0 [")][0]
    ref_line = str(golden["epoch1/log"])
    loss, mf = (float(x) for x in line.split("train==[")[1].split("+")[0].split("="))
    rloss, rmf = (float(x) for x in ref_line.split("train==[")[1].split("+")[0].split("="))
    ok = abs(loss - rloss) < 5e-4 and abs(mf - rmf) < 5e-4
    sd = tr.model_mm.state_dict()
    for k in sd:
        if not k.startswith("batch_norm"):
            ok &= bool(np.allclose(sd[k].numpy(), golden["epoch1/" + k], rtol=2e-4, atol=2e-6))
    res = tr.test(list(gen.test_set.keys()), is_val=False)
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        ok &= bool(np.allclose(res[k], golden["epoch1/metric/" + k], rtol=0, atol=1e-4))
    ua, ia = tr.hot.forward()
    _, hits = batch_test.rank_block(ua, ia, sorted(gen.test_set.keys()), False)
    ok &= int((hits.numpy() != golden["epoch1/hits"]).sum()) <= 2
    out[0] = bool(ok)
    out[1] = line


def test_trainer_epoch_matches_reference_golden_on_stand_ins(tiny_root):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(tiny_root, out), nprocs=1, join=True)
    assert out.get(0) is True, dict(out)
