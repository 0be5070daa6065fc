"""The C sampler reproduces the reference's numpy stream bit-for-bit (golden batches of the unmodified reference
and a long differential run against the Python loops)."""
import os
import random

import numpy as np

from llmrec_b200.utility.load_data import Data


def _gen(tiny_root, sampler):
    return Data(path=os.path.join(tiny_root, "netflix_valid_item"), batch_size=128, sampler=sampler)


def test_native_sampler_matches_reference_golden(tiny_root, golden):
    gen = _gen(tiny_root, "native")
    import pickle
    aug = pickle.load(open(os.path.join(tiny_root, "netflix_valid_item", "augmented_sample_dict"), "rb"))
    np.random.seed(2022); random.seed(2022)
    for b in range(3):
        u, p, n = gen.sample()
        ua = random.sample(u, int(len(u) * 0.1))
        ok = [x for x in ua if aug[x][0] < gen.n_items and aug[x][1] < gen.n_items]
        got = np.asarray([u + ok, p + [aug[x][0] for x in ok], n + [aug[x][1] for x in ok]])
        np.testing.assert_array_equal(got, golden[f"sampler/{b}"])


def test_native_equals_python_over_many_batches(tiny_root):
    a, b = _gen(tiny_root, "python"), _gen(tiny_root, "native")
    for seed in (1, 7):
        np.random.seed(seed); random.seed(seed)
        ref = [a.sample() for _ in range(40)]
        st_ref = np.random.get_state()
        np.random.seed(seed); random.seed(seed)
        got = [b.sample() for _ in range(40)]
        st_got = np.random.get_state()
        for (u1, p1, n1), (u2, p2, n2) in zip(ref, got):
            assert u1 == u2 and [int(x) for x in p1] == p2 and [int(x) for x in n1] == n2
        assert st_ref[2] == st_got[2] and (st_ref[1] == st_got[1]).all()      # generator left in the same state


def _ref_batch(gen, aug, rate, limit):
    """Data.sample() + the augmented-edge step exactly as main.py:213-224 writes it (Python lists, global RNG streams)."""
    users, pos, neg = gen.sample()
    ua = random.sample(users, int(len(users) * rate))
    keep = [u for u in ua if (aug[u][0] < limit and aug[u][1] < limit)]
    return users + keep, pos + [aug[u][0] for u in keep], neg + [aug[u][1] for u in keep]


def test_batch_sampler_matches_reference_golden(tiny_root, golden):
    """host_native.BatchSampler (users + items + augmented edges in one C call) against the batches the unmodified
    reference drew."""
    import pickle
    from llmrec_b200.host_native import BatchSampler
    gen = _gen(tiny_root, "native")
    aug = pickle.load(open(os.path.join(tiny_root, "netflix_valid_item", "augmented_sample_dict"), "rb"))
    rp, col = gen.csr("train")
    bs = BatchSampler(gen.exist_users, rp, col, gen.n_items, 128, *BatchSampler.aug_tables(aug, gen.n_users))
    out = np.zeros((3, 300), dtype=np.int32)
    np.random.seed(2022); random.seed(2022)
    for b in range(3):
        B = bs.draw(out, 0.1)
        np.testing.assert_array_equal(out[:, :B], golden[f"sampler/{b}"])


def test_batch_sampler_equals_python_in_every_branch(tmp_path):
    """CPython's random.sample has two algorithms (pool / set) and Data.sample() a third path (choice) when the batch exceeds
    the population; the augmented filter drops out-of-range ids.  Same lists and same final state of BOTH global streams."""
    import contextlib
    import io
    import pickle
    from llmrec_b200.host_native import BatchSampler, _sample_uses_pool
    from llmrec_b200.synth import make_dataset
    seen = set()
    for i, (nu, ni, ne, batch, rate) in enumerate([(300, 400, 1500, 128, 0.1), (5000, 400, 20000, 128, 0.1), (300, 400, 1500, 512, 0.1),
                                                   (300, 400, 1500, 64, 0.5), (300, 400, 1500, 4, 0.5), (3000, 500, 12000, 1024, 0.0)]):
        with contextlib.redirect_stdout(io.StringIO()):
            p = make_dataset(str(tmp_path / str(i)), n_users=nu, n_items=ni, n_inter=ne, dims=(8, 8, 8), seed=1)
            gen = Data(p, batch, sampler="python")
        aug = pickle.load(open(os.path.join(p, "augmented_sample_dict"), "rb"))
        for u in list(aug)[::3]:
            aug[u][0] = ni + 5                                               # filtered out (main.py:221)
        rp, col = gen.csr("train")
        limit = ni - 7                                                       # aug_limit differs from the negative range on purpose
        bs = BatchSampler(gen.exist_users, rp, col, gen.n_items, batch, *BatchSampler.aug_tables(aug, gen.n_users), aug_limit=limit)
        out = np.zeros((3, 2 * batch + 8), dtype=np.int32)
        random.seed(5); np.random.seed(5)
        want = [_ref_batch(gen, aug, rate, limit) for _ in range(6)]
        tail = (random.random(), float(np.random.rand()))
        random.seed(5); np.random.seed(5)
        for w in want:
            B = bs.draw(out, rate)
            assert (out[0, :B].tolist(), out[1, :B].tolist(), out[2, :B].tolist()) == tuple(w)
        assert tail == (random.random(), float(np.random.rand()))
        n_aug = int(batch * rate)
        seen.add(("choice" if batch > len(gen.exist_users) else "pool" if bs.users_pool else "set",
                  "none" if n_aug == 0 else "pool" if _sample_uses_pool(batch, n_aug) else "set"))
    assert {s[0] for s in seen} == {"choice", "pool", "set"} and {s[1] for s in seen} == {"none", "pool", "set"}


def test_batch_sampler_missing_aug_user_raises(tiny_root):
    from llmrec_b200.host_native import BatchSampler
    import pytest
    gen = _gen(tiny_root, "native")
    rp, col = gen.csr("train")
    bs = BatchSampler(gen.exist_users, rp, col, gen.n_items, 128, *BatchSampler.aug_tables({}, gen.n_users))
    with pytest.raises(KeyError):
        bs.draw(np.zeros((3, 300), dtype=np.int32), 0.1)


def test_negative_augmented_ids_wrap_like_python_indexing():
    """upstream a negative augmented id passes the `< n_items` filter (main.py:219-221) and indexes from the end; the kernels take
    row ids literally, so the tables (and the Python sampling path) apply the wrap once; ids below -n_items are marked missing."""
    from llmrec_b200.host_native import BatchSampler
    aug = {0: {0: 5, 1: -1}, 1: {0: -7, 1: 3}, 2: {0: -100, 1: 2}}
    pos, neg = BatchSampler.aug_tables(aug, 3, n_items=10)
    assert pos.tolist() == [5, 3, BatchSampler.MISSING] and neg.tolist() == [9, 3, 2]
    pos, neg = BatchSampler.aug_tables(aug, 3)                      # without n_items the raw values are kept (reference-stream tests)
    assert pos.tolist() == [5, -7, -100] and neg.tolist() == [-1, 3, 2]
