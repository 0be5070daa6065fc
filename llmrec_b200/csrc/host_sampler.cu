// Host-side BPR item sampler, bit-identical to the reference's Python loops (utility/load_data.py:166-187):
//   pos = train_items[u][ np.random.randint(0, len, size=1)[0] ]
//   neg = rejection-sample np.random.randint(0, n_items, size=1)[0] until it is not in train_items[u]
// The reference draws from numpy's GLOBAL legacy RandomState: MT19937 + masked rejection of 32-bit outputs
// (numpy/random/src/distributions: random_bounded_uint64_fill with use_masked = true; a range of 1 consumes no
// output).  The caller hands the generator over with np.random.get_state() and puts it back with set_state(), so a
// seeded run draws exactly the reference's batches -- in ~20 us instead of ~13 ms per batch.  Pure host code.
#include <limits.h>
#include <stdint.h>
#include "../../include/llmrec_b200.h"

namespace {
constexpr int N = 624, M = 397;
struct MT { uint32_t* key; int pos; };

inline void mt_regen(MT& s) {
  uint32_t* k = s.key;
  int i;
  for (i = 0; i < N - M; ++i) {
    uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
    k[i] = k[i + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  }
  for (; i < N - 1; ++i) {
    uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
    k[i] = k[i + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  }
  uint32_t y = (k[N - 1] & 0x80000000u) | (k[0] & 0x7fffffffu);
  k[N - 1] = k[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  s.pos = 0;
}
inline uint32_t mt_next32(MT& s) {
  if (s.pos == N) mt_regen(s);
  uint32_t y = s.key[s.pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// np.random.randint(0, high, size=1)[0] of the legacy RandomState for 0 < high <= 2^32
inline uint32_t legacy_randint(MT& s, uint32_t high) {
  const uint32_t rng = high - 1;
  if (rng == 0) return 0;  // no generator output consumed
  uint32_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (mt_next32(s) & mask)) > rng) {}
  return v;
}
// CPython's random.Random._randbelow_with_getrandbits(n): k = n.bit_length(); r = getrandbits(k) until r < n, where
// getrandbits(k <= 32) is one MT19937 output shifted right by 32 - k (Modules/_randommodule.c).
inline uint32_t py_randbelow(MT& s, uint32_t n) {
  int k = 32 - __builtin_clz(n);
  uint32_t r;
  do { r = mt_next32(s) >> (32 - k); } while (r >= n);
  return r;
}
// CPython's random.sample(population, k) over positions 0..n-1 (Lib/random.py): `pool_branch` (n <= setsize, decided by
// the caller with Python's own float arithmetic) swaps drawn entries out of a pool copy; otherwise draws are repeated until
// unseen.  `stamp` marks seen positions with `epoch` (no clearing between calls); `pool` holds >= n ints when pool_branch.
inline void py_sample(MT& s, int32_t n, int32_t k, int pool_branch, int32_t* pool, int32_t* stamp, int32_t epoch, int32_t* out_pos) {
  if (pool_branch) {
    for (int32_t i = 0; i < n; ++i) pool[i] = i;
    for (int32_t i = 0; i < k; ++i) {
      const uint32_t j = py_randbelow(s, (uint32_t)(n - i));
      out_pos[i] = pool[j];
      pool[j] = pool[n - i - 1];
    }
  } else {
    for (int32_t i = 0; i < k; ++i) {
      uint32_t j = py_randbelow(s, (uint32_t)n);
      while (stamp[j] == epoch) j = py_randbelow(s, (uint32_t)n);
      stamp[j] = epoch;
      out_pos[i] = (int32_t)j;
    }
  }
}
inline int sample_items_for(MT& s, const int32_t* users, int32_t nb, const int32_t* train_rowptr, const int32_t* train_col, int32_t n_items,
                            int32_t* pos_out, int32_t* neg_out) {
  for (int32_t b = 0; b < nb; ++b) {
    const int32_t u = users[b];
    const int32_t beg = train_rowptr[u], end = train_rowptr[u + 1];
    if (end <= beg) return 2;  // the reference would raise on an empty train list as well
    pos_out[b] = train_col[beg + (int32_t)legacy_randint(s, (uint32_t)(end - beg))];
    if (end - beg >= n_items) return 3;  // no negative exists: the reference loops forever
    for (;;) {
      const int32_t c = (int32_t)legacy_randint(s, (uint32_t)n_items);
      bool seen = false;
      for (int32_t e = beg; e < end; ++e) seen = seen || (train_col[e] == c);
      if (!seen) { neg_out[b] = c; break; }
    }
  }
  return 0;
}
}  // namespace

extern "C" int llmrec_host_sample_batch(uint32_t* py_key, int32_t* py_pos, uint32_t* np_key, int32_t* np_pos,
                                        const int32_t* exist_users, int32_t n_exist, int32_t batch, int32_t users_pool_branch,
                                        const int32_t* train_rowptr, const int32_t* train_col, int32_t n_items,
                                        int32_t n_aug, int32_t aug_pool_branch, const int32_t* aug_pos, const int32_t* aug_neg,
                                        int32_t n_aug_table, int32_t aug_limit, int32_t* stamp, int32_t epoch, int32_t* pool,
                                        int32_t* out, int64_t ld, int32_t* n_out) {
  MT py{py_key, *py_pos}, np{np_key, *np_pos};
  if (py.pos < 0 || py.pos > N || np.pos < 0 || np.pos > N || n_items <= 0 || n_exist <= 0 || batch <= 0 || n_aug < 0 || n_aug > batch ||
      ld < (int64_t)batch + n_aug)
    return 1;
  int32_t* users = out; int32_t* pos = out + ld; int32_t* neg = out + 2 * ld;
  // users: random.sample(exist_users, batch), or `batch` random.choice draws when the batch exceeds the population (load_data.py:158-161)
  if (batch <= n_exist) {
    py_sample(py, n_exist, batch, users_pool_branch, pool, stamp, epoch, users);
    for (int32_t i = 0; i < batch; ++i) users[i] = exist_users[users[i]];
  } else {
    for (int32_t i = 0; i < batch; ++i) users[i] = exist_users[py_randbelow(py, (uint32_t)n_exist)];
  }
  // one positive + one rejection-sampled negative per user from numpy's global stream (load_data.py:166-187)
  int rc = sample_items_for(np, users, batch, train_rowptr, train_col, n_items, pos, neg);
  if (rc) return rc;
  // augmented edges (main.py:216-224): random.sample(users, n_aug) over the batch list, kept when both ids < aug_limit (= n_items of train_mat)
  int32_t B = batch;
  if (n_aug > 0) {
    int32_t* pick = pool + batch;                                // positions into the batch list; pool[0, batch) is work space
    if (aug_pool_branch) py_sample(py, batch, n_aug, 1, pool, nullptr, 0, pick);
    else {
      for (int32_t i = 0; i < batch; ++i) pool[i] = 0;           // per-call stamp over the POSITIONS of the batch list
      py_sample(py, batch, n_aug, 0, nullptr, pool, 1, pick);
    }
    for (int32_t i = 0; i < n_aug; ++i) {
      const int32_t u = users[pick[i]];
      if (u < 0 || u >= n_aug_table || aug_pos[u] == INT32_MIN || aug_neg[u] == INT32_MIN) return 4;   // KeyError upstream
      if (aug_pos[u] < aug_limit && aug_neg[u] < aug_limit) { users[B] = u; pos[B] = aug_pos[u]; neg[B] = aug_neg[u]; ++B; }
    }
  }
  *n_out = B;
  *py_pos = py.pos; *np_pos = np.pos;
  return 0;
}

extern "C" int llmrec_host_sample_items(uint32_t* mt_key /* [624] in/out */, int32_t* mt_pos /* in/out */,
                                        const int32_t* users, int32_t n_users_in_batch,
                                        const int32_t* train_rowptr, const int32_t* train_col, int32_t n_items,
                                        int32_t* pos_out, int32_t* neg_out) {
  MT s{mt_key, *mt_pos};
  if (s.pos < 0 || s.pos > N || n_items <= 0) return 1;
  int rc = sample_items_for(s, users, n_users_in_batch, train_rowptr, train_col, n_items, pos_out, neg_out);
  if (rc) return rc;
  *mt_pos = s.pos;
  return 0;
}
