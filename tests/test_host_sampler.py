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
