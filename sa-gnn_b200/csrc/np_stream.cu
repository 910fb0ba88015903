// Stream-exact samplers (parity mode): sampleSslBatch / sampleTrainBatch / negSamp of LIU-YUXI/SA-GNN
// (model.py:252-339, DataHandler.py:28-41) with the SAME random draws as the reference for the same seeds.
//
// The reference draws from two Mersenne-Twister streams -- numpy's global RandomState (np.random.choice: model.py:271,
// 322,325; DataHandler.py:36; np.random.permutation: model.py:342) and CPython's `random` module (randint:
// model.py:277) -- both seeded in main.py:21-22.  The streams are sequential and their consumption is data dependent
// (rejection sampling), so this mode is HOST code: MT19937 + the two libraries' bounded-integer algorithms, restated
// here, over the caller's CSR arrays (the matrices the reference densifies with .toarray() every step).  The device
// samplers of sampler.cu are the throughput mode (own counter-based stream, same output contract).
//
// Published algorithms restated (numpy 1.16 .. 2.x legacy RandomState; CPython 3.x Lib/random.py, _randommodule.c):
//   np.random.seed(int)        init_genrand(seed)                                   (mt19937_seed)
//   random.seed(int)           init_by_array(32-bit limbs of |seed|)                (random_seed)
//   np.random.randint(lo, hi)  rng = hi-1-lo; 0: no draw; <= 2^32-1: mask = 2^k-1 >= rng, draw 32 bits & mask until
//                              <= rng; wider: 64 bits (high word first) the same way  (random_bounded_uint64_fill, masked)
//   np.random.choice(n)        randint(0, n);  choice(arr, m): arr[randint(0, len, m)]
//   np.random.permutation(n)   Fisher-Yates from the top: j = interval(i) (32-bit masked rejection), swap(i, j)
//   random.randint(a, b)       a + randbelow(b-a+1): k = bit_length(n), r = getrandbits(k) = next32 >> (32-k) until r < n
// Pinned by tests/test_np_stream.py against numpy / random themselves and against fixtures produced by calling the
// reference's own Recommender.sampleSslBatch / sampleTrainBatch (tests/golden/make_golden_sampler.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "sagnn_b200.h"

namespace {

constexpr int N = 624, M = 397;

inline void mt_gen(sagnn_mt19937* s) {
  uint32_t* mt = s->key;
  auto tw = [](uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  };
  int k = 0;
  for (; k < N - M; ++k) mt[k] = mt[k + M] ^ tw(mt[k], mt[k + 1]);
  for (; k < N - 1; ++k) mt[k] = mt[k + (M - N)] ^ tw(mt[k], mt[k + 1]);
  mt[N - 1] = mt[M - 1] ^ tw(mt[N - 1], mt[0]);
  s->pos = 0;
}

inline uint32_t next32(sagnn_mt19937* s) {
  if (s->pos >= N) mt_gen(s);
  uint32_t y = s->key[s->pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

inline void init_genrand(sagnn_mt19937* s, uint32_t seed) {
  for (int i = 0; i < N; ++i) {
    s->key[i] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
  }
  s->pos = N;
}

// numpy: rng = hi - 1 - lo, drawn with the legacy masked rejection
inline uint64_t np_bounded(sagnn_mt19937* s, uint64_t rng) {
  if (rng == 0) return 0;
  uint64_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
  if (rng <= 0xffffffffull) {
    if (rng == 0xffffffffull) return next32(s);
    uint32_t v;
    while ((v = next32(s) & (uint32_t)mask) > rng) {}
    return v;
  }
  uint64_t v;
  do {
    const uint64_t hi = next32(s), lo = next32(s);      // mt19937_next64: (next32 << 32) | next32
    v = ((hi << 32) | lo) & mask;
  } while (v > rng);
  return v;
}

inline int64_t np_choice(sagnn_mt19937* s, int64_t n) { return (int64_t)np_bounded(s, (uint64_t)(n - 1)); }

// CPython: random.randint(a, b)
inline int64_t py_randbelow(sagnn_mt19937* s, uint64_t n) {
  int k = 0;
  for (uint64_t t = n; t; t >>= 1) ++k;
  uint64_t r;
  do {
    if (k <= 32) {
      r = next32(s) >> (32 - k);
    } else {                                    // getrandbits(k > 32): little-endian 32-bit words, the last one shifted
      const uint64_t lo = next32(s);
      r = lo | ((uint64_t)(next32(s) >> (64 - k)) << 32);
    }
  } while (r >= n);
  return (int64_t)r;
}

// column c present (and non-zero) in the CSR row [b, e)?
inline bool row_has(const int32_t* idx, const uint8_t* nz, int64_t b, int64_t e, int32_t c) {
  const int32_t* p = std::lower_bound(idx + b, idx + e, c);
  return p != idx + e && *p == c && (!nz || nz[p - idx]);
}

}  // namespace

extern "C" void sagnn_mt19937_seed_numpy(sagnn_mt19937* s, uint32_t seed) { init_genrand(s, seed); }

extern "C" void sagnn_mt19937_seed_python(sagnn_mt19937* s, uint64_t seed) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int klen = key[1] ? 2 : 1;
  init_genrand(s, 19650218u);
  uint32_t* mt = s->key;
  int i = 1, j = 0;
  for (int k = N > klen ? N : klen; k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    if (++i >= N) { mt[0] = mt[N - 1]; i = 1; }
    if (++j >= klen) j = 0;
  }
  for (int k = N - 1; k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    if (++i >= N) { mt[0] = mt[N - 1]; i = 1; }
  }
  mt[0] = 0x80000000u;
  s->pos = N;
}

extern "C" uint32_t sagnn_mt19937_next32(sagnn_mt19937* s) { return next32(s); }

extern "C" int sagnn_np_randint(sagnn_mt19937* s, int64_t low, int64_t high, int64_t n, int64_t* out) {
  SAGNN_REQUIRE(s && (out || n == 0) && n >= 0, SAGNN_INVALID_ARG, "np_randint: NULL state / output");
  SAGNN_REQUIRE(low < high, SAGNN_INVALID_ARG, "np_randint: low >= high (numpy raises ValueError)");
  const uint64_t rng = (uint64_t)(high - 1) - (uint64_t)low;
  for (int64_t i = 0; i < n; ++i) out[i] = low + (int64_t)np_bounded(s, rng);
  return SAGNN_OK;
}

extern "C" int sagnn_np_permutation(sagnn_mt19937* s, int64_t n, int64_t* out) {
  SAGNN_REQUIRE(s && (out || n == 0) && n >= 0, SAGNN_INVALID_ARG, "np_permutation: NULL state / output");
  for (int64_t i = 0; i < n; ++i) out[i] = i;
  for (int64_t i = n - 1; i >= 1; --i) {
    const int64_t j = (int64_t)np_bounded(s, (uint64_t)i);
    std::swap(out[i], out[j]);
  }
  return SAGNN_OK;
}

extern "C" int sagnn_py_randint(sagnn_mt19937* s, int64_t a, int64_t b, int64_t* out) {
  SAGNN_REQUIRE(s && out, SAGNN_INVALID_ARG, "py_randint: NULL state / output");
  SAGNN_REQUIRE(a <= b, SAGNN_INVALID_ARG, "py_randint: empty range (random.randint raises ValueError)");
  *out = a + py_randbelow(s, (uint64_t)(b - a) + 1);
  return SAGNN_OK;
}

// model.py:304-339
extern "C" int sagnn_np_sample_ssl_batch(sagnn_mt19937* np_state, int T, const int32_t* const* indptr,
                                         const int32_t* const* indices, const uint8_t* const* nonzero,
                                         const int32_t* bat_ids, int batch, int ssl_num, int n_user, int n_item,
                                         int32_t* u_locs, int32_t* i_locs, int32_t* u_locs_seq, int64_t* n_out) {
  SAGNN_REQUIRE(np_state && indptr && indices && (bat_ids || batch == 0) && u_locs && i_locs && u_locs_seq && n_out,
                SAGNN_INVALID_ARG, "np_sample_ssl_batch: NULL argument");
  SAGNN_REQUIRE(T >= 1 && batch >= 0 && ssl_num >= 0 && n_item >= 1, SAGNN_INVALID_ARG,
                "np_sample_ssl_batch: T=%d batch=%d ssl_num=%d n_item=%d", T, batch, ssl_num, n_item);
  for (int b = 0; b < batch; ++b)
    SAGNN_REQUIRE(bat_ids[b] >= 0 && bat_ids[b] < n_user, SAGNN_OUT_OF_RANGE, "np_sample_ssl_batch: user %d outside [0,%d)",
                  bat_ids[b], n_user);
  const int64_t cap = (int64_t)batch * 2 * ssl_num;
  std::vector<int32_t> posset;
  std::vector<int64_t> all;
  for (int k = 0; k < T; ++k) {                                     // model.py:313 `for k in range(args.graphNum)`
    SAGNN_REQUIRE(indptr[k] && indices[k], SAGNN_INVALID_ARG, "np_sample_ssl_batch: NULL CSR of interval %d", k);
    const uint8_t* nz = nonzero ? nonzero[k] : nullptr;
    int32_t *ul = u_locs + k * cap, *il = i_locs + k * cap, *us = u_locs_seq + k * cap;
    int64_t cur = 0;
    for (int b = 0; b < batch; ++b) {
      const int32_t u = bat_ids[b];
      posset.clear();                                              // np.argwhere(temLabel[k][i] != 0): ascending columns
      for (int64_t e = indptr[k][u]; e < indptr[k][u + 1]; ++e)
        if (!nz || nz[e]) posset.push_back(indices[k][e]);
      const int64_t s = std::min<int64_t>(ssl_num, (int64_t)posset.size() / 2);
      if (s == 0) {
        (void)np_choice(np_state, n_item);                         // model.py:322: drawn, never used
        continue;
      }
      all.resize(2 * s);                                           // np.random.choice(posset, sslNum*2)
      for (int64_t j = 0; j < 2 * s; ++j) all[j] = posset[np_choice(np_state, (int64_t)posset.size())];
      for (int64_t j = 0; j < s; ++j) {                            // model.py:328-335: (pos, neg) interleaved
        ul[cur] = ul[cur + 1] = u;
        us[cur] = us[cur + 1] = b;
        il[cur] = (int32_t)all[j];
        il[cur + 1] = (int32_t)all[s + j];
        cur += 2;
      }
    }
    n_out[k] = cur;
  }
  return SAGNN_OK;
}

// model.py:252-302 + DataHandler.py:28-41
extern "C" int sagnn_np_sample_train_batch(sagnn_mt19937* np_state, sagnn_mt19937* py_state, const int64_t* seq_ptr,
                                           const int32_t* seq_items, const int32_t* tst_int, const int32_t* label_indptr,
                                           const int32_t* label_indices, const uint8_t* label_nonzero,
                                           const int32_t* bat_ids, int batch, int batch_pad, int train_sample_num,
                                           int pred_num, int pos_length, int n_user, int n_item, int32_t* u_locs,
                                           int32_t* i_locs, int32_t* u_locs_seq, int64_t* sequence, double* mask,
                                           int32_t* choose_out, int64_t* n_out) {
  SAGNN_REQUIRE(np_state && py_state && seq_ptr && seq_items && label_indptr && label_indices && (bat_ids || batch == 0) &&
                    u_locs && i_locs && u_locs_seq && sequence && mask && n_out,
                SAGNN_INVALID_ARG, "np_sample_train_batch: NULL argument");
  SAGNN_REQUIRE(batch >= 0 && batch_pad >= batch && train_sample_num >= 0 && pos_length >= 1 && n_item >= 1,
                SAGNN_INVALID_ARG, "np_sample_train_batch: batch=%d batch_pad=%d train_sample_num=%d pos_length=%d n_item=%d",
                batch, batch_pad, train_sample_num, pos_length, n_item);
  for (int b = 0; b < batch; ++b) {
    SAGNN_REQUIRE(bat_ids[b] >= 0 && bat_ids[b] < n_user, SAGNN_OUT_OF_RANGE,
                  "np_sample_train_batch: user %d outside [0,%d)", bat_ids[b], n_user);
    // a sequence shorter than 3 leaves posset[:-choose] empty and the reference's `sequence[i][-0:] = []` raises
    SAGNN_REQUIRE(seq_ptr[bat_ids[b] + 1] - seq_ptr[bat_ids[b]] >= 3, SAGNN_INVALID_ARG,
                  "np_sample_train_batch: user %d has fewer than 3 interactions (the reference raises ValueError at "
                  "model.py:293)", bat_ids[b]);
  }
  const int64_t half = (int64_t)batch * train_sample_num;          // temlen // 2
  std::fill(sequence, sequence + (int64_t)batch_pad * pos_length, (int64_t)0);
  std::fill(mask, mask + (int64_t)batch_pad * pos_length, 0.0);
  int64_t cur = 0;
  for (int b = 0; b < batch; ++b) {
    const int32_t u = bat_ids[b];
    const int32_t* seq = seq_items + seq_ptr[u];
    const int64_t len = seq_ptr[u + 1] - seq_ptr[u];
    const int64_t np_ = len - 1;                                   // posset = sequence[u][:-1]
    const int64_t samp = std::min<int64_t>(train_sample_num, np_);
    int64_t choose = 1;
    if (samp == 0) {
      (void)np_choice(np_state, n_item);                           // model.py:271 (train_sample_num == 0): drawn, never used
    } else {
      const int64_t top = std::max<int64_t>(std::min<int64_t>((int64_t)pred_num + 1, np_ - 3), 1);
      choose = 1 + py_randbelow(py_state, (uint64_t)top);          // randint(1, top)
      const int32_t pos = seq[np_ - choose];                       // posset[-choose]
      const int32_t last = seq[len - 1], tst = tst_int ? tst_int[u] : -1;
      const int64_t lb = label_indptr[u], le = label_indptr[u + 1];
      for (int64_t j = 0; j < samp;) {                             // negSamp
        const int32_t r = (int32_t)np_choice(np_state, n_item);
        if (!row_has(label_indices, label_nonzero, lb, le, r) && r != last && r != tst) {
          u_locs[cur] = u_locs[half + cur] = u;
          u_locs_seq[cur] = u_locs_seq[half + cur] = b;
          i_locs[cur] = pos;
          i_locs[half + cur] = r;
          ++cur; ++j;
        }
      }
    }
    if (choose_out) choose_out[b] = (int32_t)choose;
    const int64_t keep = np_ - choose;                             // posset[:-choose]
    int64_t* srow = sequence + (int64_t)b * pos_length;
    double* mrow = mask + (int64_t)b * pos_length;
    if (keep <= pos_length) {
      for (int64_t j = 0; j < keep; ++j) { srow[pos_length - keep + j] = seq[j]; mrow[pos_length - keep + j] = 1.0; }
    } else {
      for (int64_t j = 0; j < pos_length; ++j) { srow[j] = seq[keep - pos_length + j]; mrow[j] = 1.0; }
    }
  }
  // uLocs[:cur] + uLocs[temlen//2 : temlen//2 + cur]: close the gap between the positives and the negatives
  if (cur < half) {
    std::memmove(u_locs + cur, u_locs + half, sizeof(int32_t) * cur);
    std::memmove(i_locs + cur, i_locs + half, sizeof(int32_t) * cur);
    std::memmove(u_locs_seq + cur, u_locs_seq + half, sizeof(int32_t) * cur);
  }
  *n_out = 2 * cur;
  return SAGNN_OK;
}
