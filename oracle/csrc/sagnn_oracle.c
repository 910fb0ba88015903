/*
 * C restatement of the SelfGNN interval-graph propagation -- TEST INFRASTRUCTURE ONLY.
 *
 * Checker / CPU baseline for the CUDA path; never linked into or called by the
 * product library.  parity status: PARITY UNPINNED for the propagation (the
 * reference ships no vectors and TF 1.14 is not importable); validated against
 * oracle/propagate_oracle.py (numpy, op by op) in tests/test_oracle.py.
 *
 * What it follows (paths in LIU-YUXI/SA-GNN):
 *   model.py:80-92     messagePropagate: gather by edge col, segment-sum by edge
 *                      row (edge order inside a row = CSR order), LeakyReLU
 *   Utils/NNLayers.py:135-136  leakyRelu = max(leaky*x, x)
 *   model.py:118-129   per interval: E0^{l+1} = E0^l + s(A E1^l),
 *                      E1^{l+1} = E1^l + s(A^T E0^l); outputs = sum over layers
 *   model.py:250       backward = TF autodiff; recurrence restated from the op
 *                      list (MaximumGrad tie rule: z == 0 takes the leaky branch)
 *
 * Unlike the numpy oracle this one is fused (no [E,d] temporary) so that the
 * full-size configs check in seconds; the per-row summation order is still the
 * sequential edge order of TF's CPU SegmentSum.  Built by oracle/Makefile.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int sagnn_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void sagnn_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

#define DEFINE_ORACLE(REAL, SUFFIX)                                                              \
  /* Z[r,:] = sum_{e in row r} w_e * src[idx[e],:]   (model.py:86-87)                         */  \
  static void spmm_##SUFFIX(int R, int d, const int64_t* ptr, const int32_t* idx, const REAL* w, \
                            const REAL* src, REAL* z) {                                          \
    _Pragma("omp parallel for schedule(dynamic, 64)") for (int r = 0; r < R; ++r) {              \
      REAL* zr = z + (size_t)r * d;                                                              \
      for (int j = 0; j < d; ++j) zr[j] = (REAL)0;                                               \
      for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {                                            \
        const REAL* s = src + (size_t)idx[e] * d;                                                \
        if (w) {                                                                                 \
          REAL we = w[e];                                                                        \
          for (int j = 0; j < d; ++j) zr[j] += we * s[j];                                        \
        } else {                                                                                 \
          for (int j = 0; j < d; ++j) zr[j] += s[j];                                             \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
  }                                                                                              \
                                                                                                 \
  /* S[r,:] = sum_e |w_e * src[idx[e],:]|  -- the rounding scale of Z, for near-tie detection  */  \
  static void spmm_abs_##SUFFIX(int R, int d, const int64_t* ptr, const int32_t* idx,            \
                                const REAL* w, const REAL* src, REAL* z) {                       \
    _Pragma("omp parallel for schedule(dynamic, 64)") for (int r = 0; r < R; ++r) {              \
      REAL* zr = z + (size_t)r * d;                                                              \
      for (int j = 0; j < d; ++j) zr[j] = (REAL)0;                                               \
      for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {                                            \
        const REAL* s = src + (size_t)idx[e] * d;                                                \
        REAL we = w ? w[e] : (REAL)1;                                                            \
        for (int j = 0; j < d; ++j) { REAL t = we * s[j]; zr[j] += (t < 0 ? -t : t); }           \
      }                                                                                          \
    }                                                                                            \
  }                                                                                              \
                                                                                                 \
  /* one interval, forward (+ backward when gU != NULL).  Returns 0, or -1 on OOM.              \
   * masks are bytes, layout [L][U*d | I*d], 1 = MaximumGrad passes the gradient unscaled:      \
   *   mask_out (nullable)  receives the oracle's own decisions;                                \
   *   mask_in  (nullable)  overrides them in the backward (parity GIVEN the masks);            \
   *   mask_cmp (nullable)  is compared with the oracle's: stats[0] += mismatches,              \
   *                        stats[1] += mismatches with |z| > tie_tol * sum|terms| (not ties) */ \
  int sagnn_oracle_interval_##SUFFIX(                                                            \
      int U, int I, int d, int L, double leaky_d, const int64_t* uptr, const int32_t* ucol,      \
      const int64_t* iptr, const int32_t* irow, const REAL* uw, const REAL* iw, const REAL* uE,  \
      const REAL* iE, const REAL* gU, const REAL* gI, REAL* uOut, REAL* iOut, REAL* dU,          \
      REAL* dI, uint8_t* mask_out, const uint8_t* mask_in, const uint8_t* mask_cmp,              \
      double tie_tol, int64_t* stats) {                                                          \
    const REAL leaky = (REAL)leaky_d;                                                            \
    const size_t nu = (size_t)U * d, ni = (size_t)I * d;                                         \
    REAL* e0 = (REAL*)malloc(nu * sizeof(REAL));                                                 \
    REAL* e1 = (REAL*)malloc(ni * sizeof(REAL));                                                 \
    REAL* z0 = (REAL*)malloc(nu * sizeof(REAL) * (size_t)L);                                     \
    REAL* z1 = (REAL*)malloc(ni * sizeof(REAL) * (size_t)L);                                     \
    if (!e0 || !e1 || !z0 || !z1) {                                                              \
      free(e0); free(e1); free(z0); free(z1);                                                    \
      return -1;                                                                                 \
    }                                                                                            \
    memcpy(e0, uE, nu * sizeof(REAL));                       /* model.py:119 */                  \
    memcpy(e1, iE, ni * sizeof(REAL));                       /* model.py:120 */                  \
    memcpy(uOut, uE, nu * sizeof(REAL));                                                         \
    memcpy(iOut, iE, ni * sizeof(REAL));                                                         \
    for (int l = 0; l < L; ++l) {                            /* model.py:121-125 */              \
      REAL* zl0 = z0 + (size_t)l * nu;                                                           \
      REAL* zl1 = z1 + (size_t)l * ni;                                                           \
      spmm_##SUFFIX(U, d, uptr, ucol, uw, e1, zl0);          /* both use layer-l inputs */       \
      spmm_##SUFFIX(I, d, iptr, irow, iw, e0, zl1);                                              \
      if (mask_out) {                                                                            \
        uint8_t* mo = mask_out + (size_t)l * (nu + ni);                                          \
        for (size_t t = 0; t < nu; ++t) mo[t] = !(leaky * zl0[t] >= zl0[t]);                     \
        for (size_t t = 0; t < ni; ++t) mo[nu + t] = !(leaky * zl1[t] >= zl1[t]);                \
      }                                                                                          \
      if (mask_cmp && stats) {                                                                   \
        const uint8_t* mc = mask_cmp + (size_t)l * (nu + ni);                                    \
        REAL* a0 = (REAL*)malloc(nu * sizeof(REAL));                                             \
        REAL* a1 = (REAL*)malloc(ni * sizeof(REAL));                                             \
        if (!a0 || !a1) { free(a0); free(a1); free(e0); free(e1); free(z0); free(z1); return -1; } \
        spmm_abs_##SUFFIX(U, d, uptr, ucol, uw, e1, a0);                                         \
        spmm_abs_##SUFFIX(I, d, iptr, irow, iw, e0, a1);                                         \
        for (size_t t = 0; t < nu; ++t) {                                                        \
          if (mc[t] != (uint8_t)(!(leaky * zl0[t] >= zl0[t]))) {                                 \
            REAL az = zl0[t] < 0 ? -zl0[t] : zl0[t];                                             \
            stats[0]++; if ((double)az > tie_tol * (double)a0[t]) stats[1]++;                    \
          }                                                                                      \
        }                                                                                        \
        for (size_t t = 0; t < ni; ++t) {                                                        \
          if (mc[nu + t] != (uint8_t)(!(leaky * zl1[t] >= zl1[t]))) {                            \
            REAL az = zl1[t] < 0 ? -zl1[t] : zl1[t];                                             \
            stats[0]++; if ((double)az > tie_tol * (double)a1[t]) stats[1]++;                    \
          }                                                                                      \
        }                                                                                        \
        free(a0); free(a1);                                                                      \
      }                                                                                          \
      _Pragma("omp parallel for") for (size_t t = 0; t < nu; ++t) {                              \
        REAL z = zl0[t], lz = leaky * z;                                                         \
        e0[t] += (lz > z ? lz : z);                                                              \
        uOut[t] += e0[t];                                    /* tf.add_n, model.py:126 */        \
      }                                                                                          \
      _Pragma("omp parallel for") for (size_t t = 0; t < ni; ++t) {                              \
        REAL z = zl1[t], lz = leaky * z;                                                         \
        e1[t] += (lz > z ? lz : z);                                                              \
        iOut[t] += e1[t];                                                                        \
      }                                                                                          \
    }                                                                                            \
    if (gU && gI && dU && dI) {                                                                  \
      REAL* g0 = (REAL*)malloc(nu * sizeof(REAL));                                               \
      REAL* g1 = (REAL*)malloc(ni * sizeof(REAL));                                               \
      REAL* s0 = (REAL*)malloc(nu * sizeof(REAL));                                               \
      REAL* s1 = (REAL*)malloc(ni * sizeof(REAL));                                               \
      if (!g0 || !g1 || !s0 || !s1) {                                                            \
        free(g0); free(g1); free(s0); free(s1); free(e0); free(e1); free(z0); free(z1);          \
        return -1;                                                                               \
      }                                                                                          \
      memcpy(g0, gU, nu * sizeof(REAL));                                                         \
      memcpy(g1, gI, ni * sizeof(REAL));                                                         \
      for (int l = L - 1; l >= 0; --l) {                                                         \
        const REAL* zl0 = z0 + (size_t)l * nu;                                                   \
        const REAL* zl1 = z1 + (size_t)l * ni;                                                   \
        /* MaximumGrad: gradient goes to leaky*z where leaky*z >= z */                           \
        const uint8_t* mi = mask_in ? mask_in + (size_t)l * (nu + ni) : NULL;                    \
        _Pragma("omp parallel for") for (size_t t = 0; t < nu; ++t) {                            \
          REAL z = zl0[t];                                                                       \
          int pass = mi ? mi[t] : !(leaky * z >= z);                                             \
          s0[t] = pass ? g0[t] : leaky * g0[t];                                                  \
        }                                                                                        \
        _Pragma("omp parallel for") for (size_t t = 0; t < ni; ++t) {                            \
          REAL z = zl1[t];                                                                       \
          int pass = mi ? mi[nu + t] : !(leaky * z >= z);                                        \
          s1[t] = pass ? g1[t] : leaky * g1[t];                                                  \
        }                                                                                        \
        /* d/dE0^l of s(A^T E0^l) = A (m1 . g1): gather over the A-CSR; likewise for E1 */       \
        spmm_##SUFFIX(U, d, uptr, ucol, uw, s1, e0);         /* e0/e1 reused as scratch */       \
        spmm_##SUFFIX(I, d, iptr, irow, iw, s0, e1);                                             \
        _Pragma("omp parallel for") for (size_t t = 0; t < nu; ++t) g0[t] = gU[t] + g0[t] + e0[t]; \
        _Pragma("omp parallel for") for (size_t t = 0; t < ni; ++t) g1[t] = gI[t] + g1[t] + e1[t]; \
      }                                                                                          \
      memcpy(dU, g0, nu * sizeof(REAL));                                                         \
      memcpy(dI, g1, ni * sizeof(REAL));                                                         \
      free(g0); free(g1); free(s0); free(s1);                                                    \
    }                                                                                            \
    free(e0); free(e1); free(z0); free(z1);                                                      \
    return 0;                                                                                    \
  }

DEFINE_ORACLE(double, f64)
DEFINE_ORACLE(float, f32)
