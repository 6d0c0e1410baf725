/*
 * ozaki.c - exact integer reference of the INT8-slice ("Ozaki") form of the two contractions (TEST INFRASTRUCTURE, see
 * oracle.h).  Groundwork for the tcgen05 kind::i8 kernels planned in DESIGN.md section 3: it fixes the slicing convention
 * so that a device kernel's INT32 accumulators can be compared BIT-EXACTLY with this file.
 *
 * Convention: every operand row x[0..K) (one basis function over the points of a block, or one column of P_s) carries one
 * exponent e = ceil(log2 max|x|) (|x| 2^-e <= 1; e = 0 for an all-zero row) and k signed slices
 *     r_0 = x 2^-e,   s_i = clamp(rint(128 r_i), -127, 127),   r_{i+1} = 128 r_i - s_i,
 * i.e. x = 2^e sum_i s_i 128^-(i+1) + O(2^e 128^-k).  Slice products are exact integers; pairs with equal d = i + j share one
 * INT32 accumulator (|acc_d| <= (d + 1) K 127^2 < 2^31 for K <= 26000), and
 *     C[m][n] = 2^(eA_m + eB_n) sum_{d < k} 128^-(d+2) acc_d[m][n].
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

void orc_ozaki_slice_rows(const double* X, int rows, int cols, int k, signed char* S, int* e) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < rows; ++r) {
    const double* x = X + (size_t)r * cols;
    double amax = 0.0;
    for (int c = 0; c < cols; ++c) amax = fmax(amax, fabs(x[c]));
    int ex = 0;
    if (amax > 0.0) {
      frexp(amax, &ex); /* amax = f 2^ex, f in [0.5, 1) -> amax <= 2^ex; exact powers of two use ex - 1 */
      if (ldexp(1.0, ex - 1) == amax) ex -= 1;
    }
    e[r] = ex;
    for (int c = 0; c < cols; ++c) {
      double rem = ldexp(x[c], -ex); /* exact */
      for (int i = 0; i < k; ++i) {
        rem *= 128.0; /* exact: a power of two */
        double s = rint(rem);
        if (s > 127.0) s = 127.0;
        if (s < -127.0) s = -127.0;
        S[((size_t)i * rows + r) * cols + c] = (signed char)s;
        rem -= s; /* exact: both are multiples of the same ulp and |rem| <= 128 */
      }
    }
  }
}

void orc_ozaki_gemm_i32(const signed char* A, const signed char* B, int k, int m, int n, int K, int* acc) {
  memset(acc, 0, sizeof(int) * (size_t)k * m * n);
#pragma omp parallel for schedule(static)
  for (int a = 0; a < m; ++a)
    for (int i = 0; i < k; ++i)
      for (int j = 0; i + j < k; ++j) {
        const signed char* ar = A + ((size_t)i * m + a) * K;
        int* out = acc + ((size_t)(i + j) * m + a) * n;
        for (int b = 0; b < n; ++b) {
          const signed char* br = B + ((size_t)j * n + b) * K;
          int s = 0;
          for (int c = 0; c < K; ++c) s += (int)ar[c] * (int)br[c];
          out[b] += s;
        }
      }
}

void orc_ozaki_combine(const int* acc, int k, int m, int n, const int* eA, const int* eB, double* C) {
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < n; ++b) {
      double s = 0.0;
      for (int d = k - 1; d >= 0; --d) /* small terms first */
        s += ldexp((double)acc[((size_t)d * m + a) * n + b], -7 * (d + 2));
      C[(size_t)a * n + b] = ldexp(s, eA[a] + eB[b]);
    }
}
