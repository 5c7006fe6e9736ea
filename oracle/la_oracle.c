/*
 * la_oracle.c -- CPU restatement of rust-la's dense hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity ORACLE. It is never part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (rust-la_b200/csrc) has no CPU fallback and never links or calls this file.
 *
 * The reference (xasmx/rust-la, crate `la` 0.2.0) is Rust and cannot be compiled in this image
 * (no rustc/cargo; unpinned git deps `simd`, `opencl`).  Its hot loops are in-repo scalar loops, so the
 * oracle restates them in C, preserving the per-element floating-point operation order:
 *   - products and sums are rounded separately (Rust never contracts to FMA) -> build with -ffp-contract=off
 *   - no reassociation                                                       -> never -ffast-math
 * Parity is PINNED by the golden vectors of the reference's own tests (tests/golden/ref_tests.json),
 * which tests/test_oracle_golden.py checks bit-for-bit.
 *
 * Two forms per function:
 *   *_canon : the literal loop nest of the reference (same memory access pattern; this is what the CPU baseline times)
 *   *_fast  : order-preserving rearrangement (rows in parallel, contiguous inner loops); proven bit-identical to
 *             *_canon by tests/test_oracle_forms.py.  Used to make parity tests and fixtures affordable.
 *
 * Reference citations are relative to /root/reference.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic inputs: counter-based splitmix64 finaliser.  Element `idx` of the matrix
 * with seed `s` is a pure function of (s, idx), so host oracle and device generate identical data.
 * Distribution matches Matrix::random (src/matrix/mod.rs:842-851): uniform [0,1).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t la_hash64(uint64_t seed, uint64_t idx) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
ORACLE_API void oracle_fill_f64(double* dst, size_t count, uint64_t seed, uint64_t first_idx) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < count; ++i)
    dst[i] = (double)(la_hash64(seed, first_idx + i) >> 11) * 0x1.0p-53;
}
ORACLE_API void oracle_fill_f32(float* dst, size_t count, uint64_t seed, uint64_t first_idx) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < count; ++i)
    dst[i] = (float)(la_hash64(seed, first_idx + i) >> 40) * 0x1.0p-24f;
}

/* ------------------------------------------------------------------------------------------------
 * GEMM  --  impl Mul<&Matrix<T>> for &Matrix<T>, src/matrix/mod.rs:957-980 (loop nest :965-973,
 * accumulate expression :969) and Matrix::mmul, src/matrix/mmatrix.rs:82-98 (same loops).
 * C[m x n] = A[m x k] * B[k x n], all row-major with tight leading dimensions.
 * Per element: res = zero; for idx in 0..k: res = res + a[row][idx] * b[idx][col].
 * ---------------------------------------------------------------------------------------------- */
#define DEFINE_GEMM(T, SUF)                                                                              \
  /* literal i-j-k nest, strided walk down B's column exactly like m.get(idx, col) */                    \
  ORACLE_API void oracle_gemm_canon_##SUF(const T* a, const T* b, T* c, size_t m, size_t k, size_t n) {  \
    for (size_t row = 0; row < m; ++row)                                                                 \
      for (size_t col = 0; col < n; ++col) {                                                             \
        T res = (T)0;                                                                                    \
        for (size_t idx = 0; idx < k; ++idx) res = res + a[row * k + idx] * b[idx * n + col];            \
        c[row * n + col] = res;                                                                          \
      }                                                                                                  \
  }                                                                                                      \
  /* rows [row0,row1) only: used by the sampled-row parity check and the bounded CPU baseline */         \
  ORACLE_API void oracle_gemm_canon_rows_##SUF(const T* a, const T* b, T* c, size_t m, size_t k,         \
                                               size_t n, size_t row0, size_t row1, int threads) {        \
    (void)m;                                                                                             \
    _Pragma("omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)")              \
    for (size_t row = row0; row < row1; ++row)                                                           \
      for (size_t col = 0; col < n; ++col) {                                                             \
        T res = (T)0;                                                                                    \
        for (size_t idx = 0; idx < k; ++idx) res = res + a[row * k + idx] * b[idx * n + col];            \
        c[(row - row0) * n + col] = res;                                                                 \
      }                                                                                                  \
  }                                                                                                      \
  /* order-preserving fast form: i-k-j.  For every (row,col) the additions still happen for idx        \
     ascending starting from zero, product rounded before the add, so results are bit-identical. */      \
  ORACLE_API void oracle_gemm_fast_rows_##SUF(const T* a, const T* b, T* c, size_t m, size_t k,          \
                                              size_t n, size_t row0, size_t row1) {                      \
    (void)m;                                                                                             \
    _Pragma("omp parallel for schedule(static)")                                                         \
    for (size_t row = row0; row < row1; ++row) {                                                         \
      T* crow = c + (row - row0) * n;                                                                    \
      for (size_t col = 0; col < n; ++col) crow[col] = (T)0;                                             \
      for (size_t idx = 0; idx < k; ++idx) {                                                             \
        const T av = a[row * k + idx];                                                                   \
        const T* brow = b + idx * n;                                                                     \
        for (size_t col = 0; col < n; ++col) crow[col] = crow[col] + av * brow[col];                     \
      }                                                                                                  \
    }                                                                                                    \
  }                                                                                                      \
  ORACLE_API void oracle_gemm_fast_##SUF(const T* a, const T* b, T* c, size_t m, size_t k, size_t n) {   \
    oracle_gemm_fast_rows_##SUF(a, b, c, m, k, n, 0, m);                                                 \
  }

DEFINE_GEMM(double, f64)
DEFINE_GEMM(float, f32)

/* Integer instance: the reference tests Mul on integer matrices (src/matrix/mod.rs:1479-1484). */
ORACLE_API void oracle_gemm_canon_i64(const int64_t* a, const int64_t* b, int64_t* c, size_t m, size_t k, size_t n) {
  for (size_t row = 0; row < m; ++row)
    for (size_t col = 0; col < n; ++col) {
      int64_t res = 0;
      for (size_t idx = 0; idx < k; ++idx) res = res + a[row * k + idx] * b[idx * n + col];
      c[row * n + col] = res;
    }
}

/* ------------------------------------------------------------------------------------------------
 * LU  --  LUDecomposition::new, src/decomp/lu.rs:104-168.
 * In place on `lu` (caller passes a copy of A, mirroring ludata = a.get_data().clone(), :105).
 * piv[i] = original row now at row i (:108-111, :147-149); *pospivsign flips per swap (:151).
 * ---------------------------------------------------------------------------------------------- */
#define DEFINE_LU(T, SUF, ABS)                                                                           \
  ORACLE_API void oracle_lu_canon_##SUF(T* lu, size_t m, size_t n, uint64_t* piv, int* pospivsign) {     \
    for (size_t i = 0; i < m; ++i) piv[i] = i;                                                           \
    int pos = 1;                                                                                         \
    for (size_t j = 0; j < n; ++j) {                /* all n columns, also when m < n (:116) */           \
      for (size_t i = 0; i < m; ++i) {              /* :122-129 */                                       \
        T s = (T)0;                                                                                      \
        size_t kmax = i < j ? i : j;                                                                     \
        for (size_t k = 0; k < kmax; ++k) s = s + lu[i * n + k] * lu[k * n + j];                         \
        lu[i * n + j] = lu[i * n + j] - s;                                                               \
      }                                                                                                  \
      size_t p = j;                                 /* :132-137, strict '>' : lowest index wins ties */  \
      for (size_t i = j + 1; i < m; ++i)                                                                 \
        if (ABS(lu[i * n + j]) > ABS(lu[p * n + j])) p = i;                                              \
      if (p != j) {                                 /* :140-152 (only reachable when j < m) */           \
        for (size_t k = 0; k < n; ++k) {                                                                 \
          T t = lu[p * n + k]; lu[p * n + k] = lu[j * n + k]; lu[j * n + k] = t;                         \
        }                                                                                                \
        uint64_t t = piv[p]; piv[p] = piv[j]; piv[j] = t;                                                \
        pos = !pos;                                                                                      \
      }                                                                                                  \
      if (j < m && lu[j * n + j] != (T)0)           /* :156-160 true division */                         \
        for (size_t i = j + 1; i < m; ++i) lu[i * n + j] = lu[i * n + j] / lu[j * n + j];                \
    }                                                                                                    \
    *pospivsign = pos;                                                                                   \
  }                                                                                                      \
  /* Order-preserving fast form.  Column j is first gathered into a contiguous buffer (as JAMA does);    \
     each s is still the sequential k-ascending sum of separately rounded products, subtracted once.     \
     Rows i > j are independent of each other within a column, so they run in parallel and eight at a    \
     time (eight independent dependency chains, no reassociation inside a chain). */                     \
  ORACLE_API void oracle_lu_fast_##SUF(T* lu, size_t m, size_t n, uint64_t* piv, int* pospivsign) {      \
    for (size_t i = 0; i < m; ++i) piv[i] = i;                                                           \
    int pos = 1;                                                                                         \
    T* colj = (T*)malloc(sizeof(T) * (m ? m : 1));                                                       \
    for (size_t j = 0; j < n; ++j) {                                                                     \
      for (size_t i = 0; i < m; ++i) colj[i] = lu[i * n + j];                                            \
      size_t top = j < m ? j : m;                                                                        \
      for (size_t i = 0; i < top; ++i) {            /* U part: depends on colj[k], k < i, in order */     \
        const T* row = lu + i * n;                                                                       \
        T s = (T)0;                                                                                      \
        for (size_t k = 0; k < i; ++k) s = s + row[k] * colj[k];                                         \
        colj[i] = colj[i] - s;                                                                           \
      }                                                                                                  \
      if (j < m) {                                  /* L part: rows i >= j, kmax = j for all */          \
        const size_t cnt = m - j;                                                                        \
        const size_t blocks = (cnt + 7) / 8;                                                             \
        _Pragma("omp parallel for schedule(static) if (cnt * j > 65536)")                                \
        for (size_t blk = 0; blk < blocks; ++blk) {                                                      \
          size_t i0 = j + blk * 8;                                                                       \
          size_t nr = (i0 + 8 <= m) ? 8 : (m - i0);                                                      \
          if (nr == 8) {                                                                                 \
            const T *r0 = lu + (i0 + 0) * n, *r1 = lu + (i0 + 1) * n, *r2 = lu + (i0 + 2) * n,           \
                    *r3 = lu + (i0 + 3) * n, *r4 = lu + (i0 + 4) * n, *r5 = lu + (i0 + 5) * n,           \
                    *r6 = lu + (i0 + 6) * n, *r7 = lu + (i0 + 7) * n;                                    \
            T s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;                            \
            for (size_t k = 0; k < j; ++k) {                                                             \
              const T u = colj[k];                                                                       \
              s0 = s0 + r0[k] * u; s1 = s1 + r1[k] * u; s2 = s2 + r2[k] * u; s3 = s3 + r3[k] * u;        \
              s4 = s4 + r4[k] * u; s5 = s5 + r5[k] * u; s6 = s6 + r6[k] * u; s7 = s7 + r7[k] * u;        \
            }                                                                                            \
            colj[i0 + 0] = colj[i0 + 0] - s0; colj[i0 + 1] = colj[i0 + 1] - s1;                          \
            colj[i0 + 2] = colj[i0 + 2] - s2; colj[i0 + 3] = colj[i0 + 3] - s3;                          \
            colj[i0 + 4] = colj[i0 + 4] - s4; colj[i0 + 5] = colj[i0 + 5] - s5;                          \
            colj[i0 + 6] = colj[i0 + 6] - s6; colj[i0 + 7] = colj[i0 + 7] - s7;                          \
          } else {                                                                                       \
            for (size_t r = 0; r < nr; ++r) {                                                            \
              const T* row = lu + (i0 + r) * n;                                                          \
              T s = (T)0;                                                                                \
              for (size_t k = 0; k < j; ++k) s = s + row[k] * colj[k];                                   \
              colj[i0 + r] = colj[i0 + r] - s;                                                           \
            }                                                                                            \
          }                                                                                              \
        }                                                                                                \
      }                                                                                                  \
      for (size_t i = 0; i < m; ++i) lu[i * n + j] = colj[i];                                            \
      size_t p = j;                                                                                      \
      for (size_t i = j + 1; i < m; ++i)                                                                 \
        if (ABS(colj[i]) > ABS(colj[p])) p = i;                                                          \
      if (p != j) {                                                                                      \
        T* rp = lu + p * n; T* rj = lu + j * n;                                                          \
        for (size_t k = 0; k < n; ++k) { T t = rp[k]; rp[k] = rj[k]; rj[k] = t; }                        \
        uint64_t t = piv[p]; piv[p] = piv[j]; piv[j] = t;                                                \
        pos = !pos;                                                                                      \
      }                                                                                                  \
      if (j < m && lu[j * n + j] != (T)0) {                                                              \
        const T d = lu[j * n + j];                                                                       \
        for (size_t i = j + 1; i < m; ++i) lu[i * n + j] = lu[i * n + j] / d;                            \
      }                                                                                                  \
    }                                                                                                    \
    free(colj);                                                                                          \
    *pospivsign = pos;                                                                                   \
  }                                                                                                      \
  /* is_non_singular, src/decomp/lu.rs:174-182: exact == 0 test on lu[j*n+j], j < n (caller ensures      \
     the index is in range, i.e. m >= n, as the reference would panic otherwise). */                     \
  ORACLE_API int oracle_lu_is_non_singular_##SUF(const T* lu, size_t n) {                                \
    for (size_t j = 0; j < n; ++j)                                                                       \
      if (lu[j * n + j] == (T)0) return 0;                                                               \
    return 1;                                                                                            \
  }                                                                                                      \
  /* det, src/decomp/lu.rs:224-232: sequential product in index order, sign from pospivsign. */          \
  ORACLE_API T oracle_lu_det_##SUF(const T* lu, size_t n, int pospivsign) {                              \
    T d = pospivsign ? (T)1 : -(T)1;                                                                     \
    for (size_t j = 0; j < n; ++j) d = d * lu[j * n + j];                                                \
    return d;                                                                                            \
  }                                                                                                      \
  /* solve, src/decomp/lu.rs:237-278.  Returns 0 and leaves x untouched when singular (None, :241-243).  \
     x is m x nx.  Gather :246-254, forward :257-263, backward :266-275. */                              \
  ORACLE_API int oracle_lu_solve_##SUF(const T* lu, size_t m, size_t n, const uint64_t* piv,             \
                                       const T* b, size_t nx, T* x) {                                    \
    if (!oracle_lu_is_non_singular_##SUF(lu, n)) return 0;                                               \
    for (size_t i = 0; i < m; ++i)                                                                       \
      for (size_t j = 0; j < nx; ++j) x[i * nx + j] = b[piv[i] * nx + j];                                \
    for (size_t k = 0; k < n; ++k)                                                                       \
      for (size_t i = k + 1; i < n; ++i)                                                                 \
        for (size_t j = 0; j < nx; ++j) x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * lu[i * n + k];   \
    for (size_t k = n; k-- > 0;) {                                                                       \
      for (size_t j = 0; j < nx; ++j) x[k * nx + j] = x[k * nx + j] / lu[k * n + k];                     \
      for (size_t i = 0; i < k; ++i)                                                                     \
        for (size_t j = 0; j < nx; ++j) x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * lu[i * n + k];   \
    }                                                                                                    \
    return 1;                                                                                            \
  }                                                                                                      \
  /* Order-preserving parallel solve: every x[i][j] receives its updates for k ascending (forward) /     \
     descending (backward) exactly as above; rows i are independent inside a k step. */                  \
  ORACLE_API int oracle_lu_solve_fast_##SUF(const T* lu, size_t m, size_t n, const uint64_t* piv,        \
                                            const T* b, size_t nx, T* x) {                               \
    if (!oracle_lu_is_non_singular_##SUF(lu, n)) return 0;                                               \
    for (size_t i = 0; i < m; ++i)                                                                       \
      for (size_t j = 0; j < nx; ++j) x[i * nx + j] = b[piv[i] * nx + j];                                \
    _Pragma("omp parallel")                                                                              \
    {                                                                                                    \
      for (size_t k = 0; k < n; ++k) {                                                                   \
        _Pragma("omp for schedule(static)")                                                              \
        for (size_t i = k + 1; i < n; ++i)                                                               \
          for (size_t j = 0; j < nx; ++j) x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * lu[i * n + k]; \
      }                                                                                                  \
      for (size_t k = n; k-- > 0;) {                                                                     \
        _Pragma("omp single")                                                                            \
        for (size_t j = 0; j < nx; ++j) x[k * nx + j] = x[k * nx + j] / lu[k * n + k];                   \
        _Pragma("omp for schedule(static)")                                                              \
        for (size_t i = 0; i < k; ++i)                                                                   \
          for (size_t j = 0; j < nx; ++j) x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * lu[i * n + k]; \
      }                                                                                                  \
    }                                                                                                    \
    return 1;                                                                                            \
  }                                                                                                      \
  /* get_l :184-202, get_u :204-215 (unpack), used for residual checks */                               \
  ORACLE_API void oracle_lu_get_l_##SUF(const T* lu, size_t m, size_t n, T* l) {                         \
    size_t nn = m >= n ? n : m;                                                                          \
    for (size_t i = 0; i < m; ++i)                                                                       \
      for (size_t j = 0; j < nn; ++j) l[i * nn + j] = i > j ? lu[i * n + j] : (i == j ? (T)1 : (T)0);    \
  }                                                                                                      \
  ORACLE_API void oracle_lu_get_u_##SUF(const T* lu, size_t m, size_t n, T* u) {                         \
    size_t mm = m >= n ? n : m;                                                                          \
    for (size_t i = 0; i < mm; ++i)                                                                      \
      for (size_t j = 0; j < n; ++j) u[i * n + j] = i <= j ? lu[i * n + j] : (T)0;                       \
  }

DEFINE_LU(double, f64, fabs)
DEFINE_LU(float, f32, fabsf)

/* Matrix::id, src/matrix/mod.rs:416-426 (RHS of inverse, :1034-1037) */
ORACLE_API void oracle_identity_f64(double* d, size_t n) {
  memset(d, 0, sizeof(double) * n * n);
  for (size_t i = 0; i < n; ++i) d[i * n + i] = 1.0;
}
ORACLE_API void oracle_identity_f32(float* d, size_t n) {
  memset(d, 0, sizeof(float) * n * n);
  for (size_t i = 0; i < n; ++i) d[i * n + i] = 1.0f;
}

/* ||P*A - L*U||_F / ||A||_F with P*A = A(piv,:), computed in long double for the harness.
   A is regenerated from (seed) if a == NULL is not supported here: caller passes A. */
ORACLE_API double oracle_lu_backward_error_f64(const double* a, const double* lu, size_t m, size_t n,
                                               const uint64_t* piv) {
  long double num = 0, den = 0;
  size_t r = m < n ? m : n;
#pragma omp parallel for schedule(static) reduction(+ : num, den)
  for (size_t i = 0; i < m; ++i) {
    long double* acc = (long double*)calloc(n, sizeof(long double));
    size_t kmax = i < r ? i : r; /* L[i][k], k < min(i, r); plus unit diagonal if i < r */
    for (size_t k = 0; k < kmax; ++k) {
      long double l = lu[i * n + k];
      const double* urow = lu + k * n;
      for (size_t j = k; j < n; ++j) acc[j] += l * (long double)urow[j];
    }
    if (i < r)
      for (size_t j = i; j < n; ++j) acc[j] += (long double)lu[i * n + j];
    const double* arow = a + piv[i] * n;
    for (size_t j = 0; j < n; ++j) {
      long double d = (long double)arow[j] - acc[j];
      num += d * d;
      den += (long double)arow[j] * (long double)arow[j];
    }
    free(acc);
  }
  return (double)sqrtl(num / den);
}
ORACLE_API double oracle_lu_backward_error_f32(const float* a, const float* lu, size_t m, size_t n,
                                               const uint64_t* piv) {
  double num = 0, den = 0;
  size_t r = m < n ? m : n;
#pragma omp parallel for schedule(static) reduction(+ : num, den)
  for (size_t i = 0; i < m; ++i) {
    double* acc = (double*)calloc(n, sizeof(double));
    size_t kmax = i < r ? i : r;
    for (size_t k = 0; k < kmax; ++k) {
      double l = lu[i * n + k];
      const float* urow = lu + k * n;
      for (size_t j = k; j < n; ++j) acc[j] += l * (double)urow[j];
    }
    if (i < r)
      for (size_t j = i; j < n; ++j) acc[j] += (double)lu[i * n + j];
    const float* arow = a + piv[i] * n;
    for (size_t j = 0; j < n; ++j) {
      double d = (double)arow[j] - acc[j];
      num += d * d;
      den += (double)arow[j] * (double)arow[j];
    }
    free(acc);
  }
  return sqrt(num / den);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Cholesky: CholeskyDecomposition::new  src/decomp/cholesky.rs:56-110,
 * solve :116-144.  Returns 1 for Some(..), 0 for None (not square is the caller's check, :57-59; not symmetric :91-93,
 * exact `!=` so a NaN pair is "not symmetric"; not positive definite :99-102, `d <= 0`, so a NaN d is NOT rejected).
 *   canon : the reference's row-by-row loop nest, including the point at which it gives up.
 *   fast  : column by column -- every element's sum keeps its i-ascending order and separate roundings, rows of a
 *           column are independent (OpenMP); the symmetry test is hoisted (it reads only the input).  Same Some/None
 *           and, for Some, bit-identical L (tests/test_oracle_cholesky.py).
 * --------------------------------------------------------------------------------------------------------------- */
#define DEFINE_CHOL(T, SUF, SQRT)                                                                           \
  ORACLE_API int oracle_chol_canon_##SUF(const T* a, size_t n, T* l) {                                      \
    for (size_t j = 0; j < n; ++j) {                                                                        \
      T d = (T)0;                                                                                           \
      for (size_t k = 0; k < j; ++k) {                                                                      \
        T s = (T)0;                                                                                         \
        for (size_t i = 0; i < k; ++i) s = s + l[k * n + i] * l[j * n + i]; /* cholesky.rs:80-82 */          \
        s = (a[j * n + k] - s) / l[k * n + k];                              /* :85 */                        \
        l[j * n + k] = s;                                                                                   \
        d = d + s * s;                                                      /* :89 */                        \
        if (a[k * n + j] != a[j * n + k]) return 0;                         /* :92-94 */                     \
      }                                                                                                     \
      d = a[j * n + j] - d;                                                 /* :99 */                        \
      if (d <= (T)0) return 0;                                              /* :100-103 */                   \
      l[j * n + j] = SQRT(d);                                                                               \
      for (size_t k = j + 1; k < n; ++k) l[j * n + k] = (T)0;               /* :107-109 */                   \
    }                                                                                                       \
    return 1;                                                                                               \
  }                                                                                                         \
  ORACLE_API int oracle_chol_fast_##SUF(const T* a, size_t n, T* l) {                                       \
    int sym = 1;                                                                                            \
    _Pragma("omp parallel for schedule(dynamic, 16) reduction(&& : sym)")                                   \
    for (size_t j = 0; j < n; ++j)                                                                          \
      for (size_t k = 0; k < j; ++k)                                                                        \
        if (a[k * n + j] != a[j * n + k]) sym = 0;                                                          \
    /* The reference meets `d <= 0` of row j before it looks at the symmetry of rows > j, but both end in None. */ \
    for (size_t k = 0; k < n; ++k) {                                                                        \
      T d = (T)0;                                                                                           \
      for (size_t i = 0; i < k; ++i) d = d + l[k * n + i] * l[k * n + i];                                   \
      d = a[k * n + k] - d;                                                                                 \
      if (d <= (T)0) return 0;                                                                              \
      l[k * n + k] = SQRT(d);                                                                               \
      for (size_t c = k + 1; c < n; ++c) l[k * n + c] = (T)0;                                               \
      const T lkk = l[k * n + k];                                                                           \
      _Pragma("omp parallel for schedule(static)")                                                          \
      for (size_t j = k + 1; j < n; ++j) {                                                                  \
        T s = (T)0;                                                                                         \
        for (size_t i = 0; i < k; ++i) s = s + l[k * n + i] * l[j * n + i];                                 \
        l[j * n + k] = (a[j * n + k] - s) / lkk;                                                            \
      }                                                                                                     \
    }                                                                                                       \
    return sym;                                                                                             \
  }                                                                                                         \
  /* solve :116-144: forward with L (divide by the diagonal), backward with L' */                           \
  ORACLE_API void oracle_chol_solve_##SUF(const T* l, size_t n, const T* b, size_t nx, T* x) {              \
    memcpy(x, b, n * nx * sizeof(T));                                                                       \
    for (size_t k = 0; k < n; ++k)                                                                          \
      for (size_t j = 0; j < nx; ++j) {                                                                     \
        for (size_t i = 0; i < k; ++i) x[k * nx + j] = x[k * nx + j] - x[i * nx + j] * l[k * n + i];        \
        x[k * nx + j] = x[k * nx + j] / l[k * n + k];                                                       \
      }                                                                                                     \
    for (size_t k = n; k-- > 0;)                                                                            \
      for (size_t j = 0; j < nx; ++j) {                                                                     \
        for (size_t i = k + 1; i < n; ++i) x[k * nx + j] = x[k * nx + j] - x[i * nx + j] * l[i * n + k];    \
        x[k * nx + j] = x[k * nx + j] / l[k * n + k];                                                       \
      }                                                                                                     \
  }
DEFINE_CHOL(double, f64, sqrt)
DEFINE_CHOL(float, f32, sqrtf)

/* ---------------------------------------------------------------------------------------------------------------
 * QR: QRDecomposition::new  src/decomp/qr.rs:26-106 (Householder reflections, the
 * k-th vector u = x - a e_k left UNNORMALISED in column k from the diagonal down, a = -+|x| in rdiag[k]; a zero column
 * is skipped :62), get_q :151-194, solve :199-238.
 *   canon : the literal loop nests (column walks with stride n).
 *   fast  : per reflection, rows outermost and columns innermost, columns split over threads -- every dot product
 *           still receives its terms row-ascending with separately rounded multiply/add, so the packed qr and rdiag
 *           are bit-identical to canon (tests/test_oracle_qr.py).
 * Quirks kept by the callers (oracle.py / the mirrors), not here: is_full_rank indexes rdiag[0..cols) (:110-117,
 * out of bounds for m < n) and solve builds Matrix::new(cols, nx, <m*nx values>) (:237, panics unless m == n).
 * --------------------------------------------------------------------------------------------------------------- */
#define DEFINE_QR(T, SUF, SQRT)                                                                             \
  ORACLE_API void oracle_qr_canon_##SUF(const T* a, size_t m, size_t n, T* qr, T* rdiag) {                  \
    memcpy(qr, a, m * n * sizeof(T));                                                                       \
    size_t dc = m < n ? m : n;                                                                              \
    for (size_t minor = 0; minor < dc; ++minor) {                                                           \
      T x_norm_sqr = (T)0;                                                                                  \
      for (size_t i = minor; i < m; ++i) {                                  /* qr.rs:50-53 */                \
        T c = qr[i * n + minor];                                                                            \
        x_norm_sqr = x_norm_sqr + c * c;                                                                    \
      }                                                                                                     \
      T aa = qr[minor * n + minor] > (T)0 ? -SQRT(x_norm_sqr) : SQRT(x_norm_sqr); /* :58 */                 \
      rdiag[minor] = aa;                                                                                    \
      if (aa != (T)0) {                                                     /* :62 */                        \
        qr[minor * n + minor] = qr[minor * n + minor] - aa;                 /* :77 */                        \
        for (size_t column = minor + 1; column < n; ++column) {             /* :94-105 */                    \
          T x_dot_u = (T)0;                                                                                 \
          for (size_t row = minor; row < m; ++row) x_dot_u = x_dot_u + qr[row * n + minor] * qr[row * n + column]; \
          T factor = x_dot_u / (aa * qr[minor * n + minor]);                                                \
          for (size_t row = minor; row < m; ++row)                                                          \
            qr[row * n + column] = qr[row * n + column] + factor * qr[row * n + minor];                     \
        }                                                                                                   \
      }                                                                                                     \
    }                                                                                                       \
  }                                                                                                         \
  ORACLE_API void oracle_qr_fast_##SUF(const T* a, size_t m, size_t n, T* qr, T* rdiag) {                   \
    memcpy(qr, a, m * n * sizeof(T));                                                                       \
    size_t dc = m < n ? m : n;                                                                              \
    T* dots = (T*)malloc(n * sizeof(T));                                                                    \
    T* u = (T*)malloc(m * sizeof(T));                                                                       \
    for (size_t minor = 0; minor < dc; ++minor) {                                                           \
      T x_norm_sqr = (T)0;                                                                                  \
      for (size_t i = minor; i < m; ++i) {                                                                  \
        T c = qr[i * n + minor];                                                                            \
        x_norm_sqr = x_norm_sqr + c * c;                                                                    \
      }                                                                                                     \
      T aa = qr[minor * n + minor] > (T)0 ? -SQRT(x_norm_sqr) : SQRT(x_norm_sqr);                           \
      rdiag[minor] = aa;                                                                                    \
      if (aa == (T)0) continue;                                                                             \
      qr[minor * n + minor] = qr[minor * n + minor] - aa;                                                   \
      for (size_t row = minor; row < m; ++row) u[row] = qr[row * n + minor];                                \
      const T den = aa * qr[minor * n + minor];                                                             \
      const size_t c0 = minor + 1;                                                                          \
      _Pragma("omp parallel")                                                                               \
      {                                                                                                     \
        int nt = 1, tid = 0;                                                                                \
        OMP_IDS(nt, tid)                                                                                    \
        size_t span = n - c0, per = (span + (size_t)nt - 1) / (size_t)nt;                                   \
        size_t lo = c0 + per * (size_t)tid, hi = lo + per;                                                  \
        if (hi > n) hi = n;                                                                                 \
        if (lo < hi) {                                                                                      \
          for (size_t c = lo; c < hi; ++c) dots[c] = (T)0;                                                  \
          for (size_t row = minor; row < m; ++row) {                                                        \
            const T ur = u[row];                                                                            \
            const T* r = qr + row * n;                                                                      \
            for (size_t c = lo; c < hi; ++c) dots[c] = dots[c] + ur * r[c];                                 \
          }                                                                                                 \
          for (size_t c = lo; c < hi; ++c) dots[c] = dots[c] / den;                                         \
          for (size_t row = minor; row < m; ++row) {                                                        \
            const T ur = u[row];                                                                            \
            T* r = qr + row * n;                                                                            \
            for (size_t c = lo; c < hi; ++c) r[c] = r[c] + dots[c] * ur;                                    \
          }                                                                                                 \
        }                                                                                                   \
      }                                                                                                     \
    }                                                                                                       \
    free(dots);                                                                                             \
    free(u);                                                                                                \
  }                                                                                                         \
  /* get_q :151-194: the reflections applied in reverse order to the m x m identity (columns minor.. only) */ \
  ORACLE_API void oracle_qr_get_q_##SUF(const T* qr, const T* rdiag, size_t m, size_t n, T* q) {            \
    size_t dc = m < n ? m : n;                                                                              \
    for (size_t i = 0; i < m * m; ++i) q[i] = (T)0;                                                         \
    for (size_t k = 0; k < dc; ++k) q[k * m + k] = (T)1;                                                    \
    for (size_t minor = dc; minor-- > 0;) {                                                                 \
      if (qr[minor * n + minor] == (T)0) continue;                          /* :166 */                       \
      const T den = rdiag[minor] * qr[minor * n + minor];                                                   \
      _Pragma("omp parallel for schedule(static)")                                                          \
      for (size_t column = minor; column < m; ++column) {                                                   \
        T x_dot_u = (T)0;                                                                                   \
        for (size_t row = minor; row < m; ++row) x_dot_u = x_dot_u + qr[row * n + minor] * q[row * m + column]; \
        T factor = x_dot_u / den;                                                                           \
        for (size_t row = minor; row < m; ++row) q[row * m + column] = q[row * m + column] + factor * qr[row * n + minor]; \
      }                                                                                                     \
    }                                                                                                       \
  }                                                                                                         \
  /* solve :199-238 on the full m x nx work array (the caller keeps the reference's shape quirk); x = m*nx values */ \
  ORACLE_API void oracle_qr_solve_##SUF(const T* qr, const T* rdiag, size_t m, size_t n, const T* b, size_t nx, T* x) { \
    memcpy(x, b, m * nx * sizeof(T));                                                                       \
    for (size_t k = 0; k < n; ++k)                                          /* Y = Q' B, :214-224 */         \
      for (size_t j = 0; j < nx; ++j) {                                                                     \
        T s = (T)0;                                                                                         \
        for (size_t i = k; i < m; ++i) s = s + qr[i * n + k] * x[i * nx + j];                               \
        s = -s / qr[k * n + k];                                                                             \
        for (size_t i = k; i < m; ++i) x[i * nx + j] = x[i * nx + j] + s * qr[i * n + k];                   \
      }                                                                                                     \
    for (size_t k = n; k-- > 0;) {                                          /* R X = Y, :227-236 */          \
      for (size_t j = 0; j < nx; ++j) x[k * nx + j] = x[k * nx + j] / rdiag[k];                             \
      for (size_t i = 0; i < k; ++i)                                                                        \
        for (size_t j = 0; j < nx; ++j) x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * qr[i * n + k];      \
    }                                                                                                       \
  }
#ifdef _OPENMP
#define OMP_IDS(nt, tid) nt = omp_get_num_threads(); tid = omp_get_thread_num();
#else
#define OMP_IDS(nt, tid)
#endif
DEFINE_QR(double, f64, sqrt)
DEFINE_QR(float, f32, sqrtf)
