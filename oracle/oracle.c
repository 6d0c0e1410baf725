/*
 * oracle.c - CPU restatement of the grid path (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Rows 8a-1 (basis functions on grid), 8a-2 (density on grid), 8a-5 (grid -> matrix), 8a-6/7 (the two
 * Potential classes) of SURVEY.md section 8.  Same blocking, thresholds and loop structure as the reference:
 * one `omp parallel for schedule(dynamic)` over blocks per phase, per-thread workspaces, basis functions
 * evaluated once per phase (i.e. twice per Fock build), per-thread nb x nb accumulators + serial reduction.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_max_threads(void) { return 1; }
static int omp_get_thread_num(void) { return 0; }
static double omp_get_wtime(void) { return 0.0; }
static void omp_set_num_threads(int n) { (void)n; }
#endif

#include "harmonics_table.h"
#include "harmonics_gen.h"

#define ORC_AM_MAX 6 /* parameters/Constants.h:31 */

int orc_max_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }
/* thread 0's share of the work, split by phase (bench.py reports the GEMM rate per core from it) */
static double g_probe[3]; /* seconds in basis evaluation, seconds in the dense products, flops of those products */
void orc_probe_reset(void) { g_probe[0] = g_probe[1] = g_probe[2] = 0.0; }
void orc_probe_get(double* out3) { memcpy(out3, g_probe, sizeof(g_probe)); }

/* nBlocks = ceil(nPoints / maxBlockSize), BasisFunctionOnGridController.cpp:83 */
int orc_nblocks(const orc_grid* g) { return (int)((g->npts + g->blocksize - 1) / g->blocksize); }

static int block_size(const orc_grid* g, int block) {
  /* BasisFunctionOnGridController.cpp:170-179 */
  int nb = orc_nblocks(g);
  if (block == nb - 1) {
    int r = (int)(g->npts % g->blocksize);
    return r == 0 ? g->blocksize : r;
  }
  return g->blocksize;
}

static int nfunc_of_shell(int l, int pure) { return pure ? 2 * l + 1 : (l + 1) * (l + 2) / 2; }

/* ------------------------------------------------------------------------------------------------
 * 8a-1  basis functions on one block
 * ------------------------------------------------------------------------------------------------ */

/* monomial x^a y^b z^c differentiated (da,db,dc) times: returns coefficient factor and lowers exponents */
static inline double mono_deriv(const double* x, const double* y, const double* z, int a, int b, int c, int da,
                                int db, int dc) {
  double f = 1.0;
  for (int i = 0; i < da; ++i) {
    if (a == 0) return 0.0;
    f *= a;
    --a;
  }
  for (int i = 0; i < db; ++i) {
    if (b == 0) return 0.0;
    f *= b;
    --b;
  }
  for (int i = 0; i < dc; ++i) {
    if (c == 0) return 0.0;
    f *= c;
    --c;
  }
  return f * x[a] * y[b] * z[c];
}

static inline double harm_eval(const orc_harm_t* h, const double* x, const double* y, const double* z, int da,
                               int db, int dc) {
  double s = 0.0;
  for (int t = 0; t < h->nterms; ++t) s += h->coef[t] * mono_deriv(x, y, z, h->ex[t], h->ey[t], h->ez[t], da, db, dc);
  return s;
}

/* Block prescreening, BasisFunctionOnGridController.cpp:211-255.  shell_neg[nshell]. */
static void block_prescreen(const orc_basis* b, const orc_grid* g, double radial_thr, long first, int n,
                            int* shell_neg, double* centre) {
  double c[3] = {0.0, 0.0, 0.0};
  for (int p = 0; p < n; ++p)
    for (int k = 0; k < 3; ++k) c[k] += g->xyz[3 * (first + p) + k];
  for (int k = 0; k < 3; ++k) c[k] /= (double)n;
  double spread = 0.0;
  for (int p = 0; p < n; ++p) {
    double d0 = g->xyz[3 * (first + p)] - c[0], d1 = g->xyz[3 * (first + p) + 1] - c[1],
           d2 = g->xyz[3 * (first + p) + 2] - c[2];
    double d = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (d > spread) spread = d;
  }
  if (centre) memcpy(centre, c, sizeof(c));
  for (int s = 0; s < b->nshell; ++s) {
    shell_neg[s] = 0;
    const double px = c[0] - b->centre[3 * s], py = c[1] - b->centre[3 * s + 1], pz = c[2] - b->centre[3 * s + 2];
    double dist = sqrt(px * px + py * py + pz * pz) - spread;
    if (dist < 1.0) continue; /* :235 */
    dist = dist * dist;
    double radial = 0.0;
    const double* al = b->alpha + b->prim_off[s];
    const double* co = b->coeff + b->prim_off[s];
    for (int i = 0; i < b->nprim[s]; ++i) radial += co[i] * exp(-(al[i] * dist));
    if (fabs(radial) < radial_thr) shell_neg[s] = 1; /* :248 */
  }
}

/* Evaluate all non-negligible shells on the n points of a block.  Arrays are n x nbf column-major with
 * leading dimension n.  BasisFunctionOnGridController.cpp:259-1098. */
static void block_evaluate(const orc_basis* b, const orc_grid* g, double radial_thr, int deriv, long first, int n,
                           const int* shell_neg, double* val, double* d1[3], double* d2[6]) {
  const double exp_thr = -log(radial_thr); /* _exponentThreshold, :86 */
  for (int p = 0; p < n; ++p) {
    const double pX = g->xyz[3 * (first + p)], pY = g->xyz[3 * (first + p) + 1], pZ = g->xyz[3 * (first + p) + 2];
    for (int s = 0; s < b->nshell; ++s) {
      if (shell_neg[s]) continue;
      const int l = b->l[s];
      const int nf = nfunc_of_shell(l, b->pure[s]);
      const long i0 = (long)b->first_bf[s] * n + p;
      const double dxp = pX - b->centre[3 * s], dyp = pY - b->centre[3 * s + 1], dzp = pZ - b->centre[3 * s + 2];
      const double r2 = dxp * dxp + dyp * dyp + dzp * dzp;
      double radial = 0.0, dradial = 0.0, ddradial = 0.0;
      const double* al = b->alpha + b->prim_off[s];
      const double* co = b->coeff + b->prim_off[s];
      for (int i = 0; i < b->nprim[s]; ++i) {
        const double tmp = al[i] * r2;
        if (tmp < exp_thr) { /* :300 */
          const double e = exp(-tmp);
          radial += co[i] * e;
          dradial -= 2.0 * al[i] * co[i] * e;
          ddradial += 4.0 * al[i] * al[i] * co[i] * e;
        }
      }
      if (fabs(radial) < radial_thr) { /* :312-329 */
        for (int m = 0; m < nf; ++m) {
          val[i0 + (long)m * n] = 0.0;
          if (deriv >= 1)
            for (int k = 0; k < 3; ++k) d1[k][i0 + (long)m * n] = 0.0;
          if (deriv >= 2)
            for (int k = 0; k < 6; ++k) d2[k][i0 + (long)m * n] = 0.0;
        }
        continue;
      }
      double x[ORC_AM_MAX + 3], y[ORC_AM_MAX + 3], z[ORC_AM_MAX + 3];
      x[0] = y[0] = z[0] = 1.0;
      for (int e = 1; e <= l + 2; ++e) {
        x[e] = x[e - 1] * dxp;
        y[e] = y[e - 1] * dyp;
        z[e] = z[e - 1] * dzp;
      }
      if (!b->pure[s]) {
        /* Cartesian, :359-440; order a = l..0, b = l-a..0 */
        int m = 0;
        for (int a = l; a >= 0; --a) {
          for (int bb = l - a; bb >= 0; --bb, ++m) {
            const int c = l - a - bb;
            const double nrm = b->normfac[b->first_bf[s] + m];
            const long idx = i0 + (long)m * n;
            val[idx] = x[a] * y[bb] * z[c] * radial * nrm;
            if (deriv >= 1) {
              double vx = dradial * x[a + 1] * y[bb] * z[c] * nrm;
              if (a > 0) vx += a * x[a - 1] * y[bb] * z[c] * radial * nrm;
              double vy = dradial * x[a] * y[bb + 1] * z[c] * nrm;
              if (bb > 0) vy += bb * x[a] * y[bb - 1] * z[c] * radial * nrm;
              double vz = dradial * x[a] * y[bb] * z[c + 1] * nrm;
              if (c > 0) vz += c * x[a] * y[bb] * z[c - 1] * radial * nrm;
              d1[0][idx] = vx;
              d1[1][idx] = vy;
              d1[2][idx] = vz;
            }
            if (deriv >= 2) {
              /* d2/dq2 [q^k R] = dd*q^(k+2) + d*(2k+1) q^k + k(k-1) q^(k-2) R, :384-437 */
              double hxx = (ddradial * x[a + 2] + dradial * x[a] * (2 * a + 1)) * y[bb] * z[c] * nrm;
              if (a > 1) hxx += a * (a - 1) * x[a - 2] * y[bb] * z[c] * radial * nrm;
              double hyy = (ddradial * y[bb + 2] + dradial * y[bb] * (2 * bb + 1)) * x[a] * z[c] * nrm;
              if (bb > 1) hyy += bb * (bb - 1) * y[bb - 2] * x[a] * z[c] * radial * nrm;
              double hzz = (ddradial * z[c + 2] + dradial * z[c] * (2 * c + 1)) * x[a] * y[bb] * nrm;
              if (c > 1) hzz += c * (c - 1) * z[c - 2] * x[a] * y[bb] * radial * nrm;
              double hxy = ddradial * x[a + 1] * y[bb + 1] * z[c] * nrm;
              if (a > 0) hxy += dradial * a * x[a - 1] * y[bb + 1] * z[c] * nrm;
              if (bb > 0) {
                hxy += dradial * bb * x[a + 1] * y[bb - 1] * z[c] * nrm;
                if (a > 0) hxy += a * bb * x[a - 1] * y[bb - 1] * z[c] * nrm * radial;
              }
              double hxz = ddradial * x[a + 1] * y[bb] * z[c + 1] * nrm;
              if (a > 0) hxz += dradial * a * x[a - 1] * y[bb] * z[c + 1] * nrm;
              if (c > 0) {
                hxz += dradial * c * x[a + 1] * y[bb] * z[c - 1] * nrm;
                if (a > 0) hxz += a * c * x[a - 1] * y[bb] * z[c - 1] * nrm * radial;
              }
              double hyz = ddradial * x[a] * y[bb + 1] * z[c + 1] * nrm;
              if (bb > 0) hyz += dradial * bb * x[a] * y[bb - 1] * z[c + 1] * nrm;
              if (c > 0) {
                hyz += dradial * c * x[a] * y[bb + 1] * z[c - 1] * nrm;
                if (bb > 0) hyz += bb * c * x[a] * y[bb - 1] * z[c - 1] * nrm * radial;
              }
              d2[0][idx] = hxx;
              d2[1][idx] = hxy;
              d2[2][idx] = hxz;
              d2[3][idx] = hyy;
              d2[4][idx] = hyz;
              d2[5][idx] = hzz;
            }
          }
        }
      } else {
        /* spherical, :441-1095: phi = R*Y, d_x phi = R dY/dx + R' x Y, ... (finalisation :1068-1095) */
        if (deriv <= 1 && l <= ORC_GEN_LMAX) { /* straight-line formulas per l, like the reference's switch(l) */
          double Y[2 * ORC_GEN_LMAX + 1], Yx[2 * ORC_GEN_LMAX + 1], Yy[2 * ORC_GEN_LMAX + 1], Yz[2 * ORC_GEN_LMAX + 1];
          switch (l) {
            case 0: orc_harm_l0(x, y, z, Y, Yx, Yy, Yz); break;
            case 1: orc_harm_l1(x, y, z, Y, Yx, Yy, Yz); break;
            case 2: orc_harm_l2(x, y, z, Y, Yx, Yy, Yz); break;
            case 3: orc_harm_l3(x, y, z, Y, Yx, Yy, Yz); break;
            default: orc_harm_l4(x, y, z, Y, Yx, Yy, Yz); break;
          }
          for (int m = 0; m < nf; ++m) {
            const long idx = i0 + (long)m * n;
            val[idx] = radial * Y[m];
            if (deriv >= 1) {
              d1[0][idx] = radial * Yx[m] + dradial * x[1] * Y[m];
              d1[1][idx] = radial * Yy[m] + dradial * y[1] * Y[m];
              d1[2][idx] = radial * Yz[m] + dradial * z[1] * Y[m];
            }
          }
          continue;
        }
        for (int m = 0; m < nf; ++m) {
          const orc_harm_t* h = &ORC_HARM[l][m];
          const long idx = i0 + (long)m * n;
          const double Y = harm_eval(h, x, y, z, 0, 0, 0);
          val[idx] = radial * Y;
          if (deriv >= 1) {
            const double Yx = harm_eval(h, x, y, z, 1, 0, 0), Yy = harm_eval(h, x, y, z, 0, 1, 0),
                         Yz = harm_eval(h, x, y, z, 0, 0, 1);
            d1[0][idx] = radial * Yx + dradial * x[1] * Y;
            d1[1][idx] = radial * Yy + dradial * y[1] * Y;
            d1[2][idx] = radial * Yz + dradial * z[1] * Y;
            if (deriv >= 2) {
              const double Yxx = harm_eval(h, x, y, z, 2, 0, 0), Yxy = harm_eval(h, x, y, z, 1, 1, 0),
                           Yxz = harm_eval(h, x, y, z, 1, 0, 1), Yyy = harm_eval(h, x, y, z, 0, 2, 0),
                           Yyz = harm_eval(h, x, y, z, 0, 1, 1), Yzz = harm_eval(h, x, y, z, 0, 0, 2);
              d2[0][idx] = radial * Yxx + 2.0 * dradial * x[1] * Yx + ddradial * x[2] * Y + dradial * Y;
              d2[1][idx] = radial * Yxy + dradial * x[1] * Yy + dradial * y[1] * Yx + ddradial * x[1] * y[1] * Y;
              d2[2][idx] = radial * Yxz + dradial * x[1] * Yz + dradial * z[1] * Yx + ddradial * x[1] * z[1] * Y;
              d2[3][idx] = radial * Yyy + 2.0 * dradial * y[1] * Yy + ddradial * y[2] * Y + dradial * Y;
              d2[4][idx] = radial * Yyz + dradial * y[1] * Yz + dradial * z[1] * Yy + ddradial * y[1] * z[1] * Y;
              d2[5][idx] = radial * Yzz + 2.0 * dradial * z[1] * Yz + ddradial * z[2] * Y + dradial * Y;
            }
          }
        }
      }
    }
  }
}

int orc_basis_block(const orc_basis* b, const orc_grid* g, double radial_thr, int deriv, int block, double* val,
                    double* dx, double* dy, double* dz, double* hxx, double* hxy, double* hxz, double* hyy,
                    double* hyz, double* hzz, int* negligible, double* centre) {
  const int n = block_size(g, block);
  const long first = (long)block * g->blocksize;
  int* shell_neg = (int*)malloc(sizeof(int) * (size_t)b->nshell);
  block_prescreen(b, g, radial_thr, first, n, shell_neg, centre);
  const size_t sz = (size_t)n * (size_t)b->nbf;
  memset(val, 0, sz * sizeof(double));
  double* d1[3] = {dx, dy, dz};
  double* d2[6] = {hxx, hxy, hxz, hyy, hyz, hzz};
  if (deriv >= 1)
    for (int k = 0; k < 3; ++k) memset(d1[k], 0, sz * sizeof(double));
  if (deriv >= 2)
    for (int k = 0; k < 6; ++k) memset(d2[k], 0, sz * sizeof(double));
  block_evaluate(b, g, radial_thr, deriv, first, n, shell_neg, val, d1, d2);
  for (int s = 0; s < b->nshell; ++s) {
    const int nf = nfunc_of_shell(b->l[s], b->pure[s]);
    for (int m = 0; m < nf; ++m) negligible[b->first_bf[s] + m] = shell_neg[s];
  }
  free(shell_neg);
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * per-thread workspace (reference: BasisFunctionOnGridController.cpp:90-99) and the projection
 * (misc/HelperFunctions.h:117-137: selector of the non-negligible columns)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int* shell_neg;
  int* sig;       /* indices of significant basis functions */
  double* val;    /* bs x nbf */
  double* d1[3];
  double* d2[6];
  double* cval;   /* compacted: bs x s */
  double* cd1[3];
  double* Ps;     /* s x s */
  double* BP;     /* bs x s */
  double* G;      /* bs x s */
  double* T;      /* s x s */
} orc_ws;

static void ws_alloc(orc_ws* w, const orc_basis* b, int bs, int deriv) {
  const size_t nb = (size_t)b->nbf, sz = (size_t)bs * nb;
  w->shell_neg = (int*)malloc(sizeof(int) * (size_t)b->nshell);
  w->sig = (int*)malloc(sizeof(int) * nb);
  w->val = (double*)malloc(sizeof(double) * sz);
  for (int k = 0; k < 3; ++k) w->d1[k] = deriv >= 1 ? (double*)malloc(sizeof(double) * sz) : NULL;
  for (int k = 0; k < 6; ++k) w->d2[k] = deriv >= 2 ? (double*)malloc(sizeof(double) * sz) : NULL;
  w->cval = (double*)malloc(sizeof(double) * sz);
  for (int k = 0; k < 3; ++k) w->cd1[k] = deriv >= 1 ? (double*)malloc(sizeof(double) * sz) : NULL;
  w->Ps = (double*)malloc(sizeof(double) * nb * nb);
  w->BP = (double*)malloc(sizeof(double) * sz);
  w->G = (double*)malloc(sizeof(double) * sz);
  w->T = (double*)malloc(sizeof(double) * nb * nb);
}

static void ws_free(orc_ws* w) {
  free(w->shell_neg);
  free(w->sig);
  free(w->val);
  for (int k = 0; k < 3; ++k) free(w->d1[k]);
  for (int k = 0; k < 6; ++k) free(w->d2[k]);
  free(w->cval);
  for (int k = 0; k < 3; ++k) free(w->cd1[k]);
  free(w->Ps);
  free(w->BP);
  free(w->G);
  free(w->T);
}

/* evaluate block into the workspace, build the significant list; returns s */
static int ws_block(orc_ws* w, const orc_basis* b, const orc_grid* g, double radial_thr, int deriv, int block,
                    int* n_out, double* t_basis) {
  const int n = block_size(g, block);
  const long first = (long)block * g->blocksize;
  const int probe = omp_get_thread_num() == 0;
  const double t0 = probe ? omp_get_wtime() : 0.0;
  block_prescreen(b, g, radial_thr, first, n, w->shell_neg, NULL);
  block_evaluate(b, g, radial_thr, deriv, first, n, w->shell_neg, w->val, w->d1, w->d2);
  if (probe) {
    const double dt = omp_get_wtime() - t0;
    g_probe[0] += dt;
    if (t_basis) *t_basis += dt;
  }
  int s = 0;
  for (int sh = 0; sh < b->nshell; ++sh) {
    if (w->shell_neg[sh]) continue;
    const int nf = nfunc_of_shell(b->l[sh], b->pure[sh]);
    for (int m = 0; m < nf; ++m) w->sig[s++] = b->first_bf[sh] + m;
  }
  *n_out = n;
  return s;
}

/* ------------------------------------------------------------------------------------------------
 * small dense kernels (stand-in for Eigen's GEMM; column-major)
 * ------------------------------------------------------------------------------------------------ */
/* C (m x n) = A (m x k) * B (k x n) */
/* Optional BLAS back end (bench.py: OpenBLAS cblas_dgemm from the scipy wheel, single-threaded per call inside the
 * OpenMP loop over blocks): the reference's products are Eigen GEMMs (MatrixOperatorToGridTransformer.cpp:157,
 * ScalarOperatorToMatrixAdder.cpp:282); a tuned library GEMM is the honest stand-in for the timed CPU baseline.  Without
 * it the compiler-vectorised loops below run (golden-vector tests, machines without the wheel). */
typedef void (*orc_dgemm_fn)(int order, int transa, int transb, int m, int n, int k, double alpha, const double* A, int lda,
                             const double* B, int ldb, double beta, double* C, int ldc);
static orc_dgemm_fn g_dgemm = NULL;
void orc_set_dgemm(void* fn) { g_dgemm = (orc_dgemm_fn)fn; }
int orc_has_dgemm(void) { return g_dgemm != NULL; }
enum { ORC_COLMAJOR = 102, ORC_NOTRANS = 111, ORC_TRANS = 112 };

static void gemm_nn_impl(int m, int n, int k, const double* restrict A, int lda, const double* restrict B, int ldb,
                    double* restrict C, int ldc) {
  if (g_dgemm && m > 0 && n > 0 && k > 0) {
    g_dgemm(ORC_COLMAJOR, ORC_NOTRANS, ORC_NOTRANS, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, ldc);
    return;
  }
  for (int j = 0; j < n; ++j) memset(C + (size_t)j * ldc, 0, sizeof(double) * (size_t)m);
  int j = 0;
  for (; j + 4 <= n; j += 4) {
    double* restrict c0 = C + (size_t)j * ldc;
    double* restrict c1 = c0 + ldc;
    double* restrict c2 = c1 + ldc;
    double* restrict c3 = c2 + ldc;
    int kk = 0;
    for (; kk + 4 <= k; kk += 4) {
      const double* restrict a0 = A + (size_t)kk * lda;
      const double* restrict a1 = a0 + lda;
      const double* restrict a2 = a1 + lda;
      const double* restrict a3 = a2 + lda;
      const double b00 = B[kk + (size_t)j * ldb], b10 = B[kk + 1 + (size_t)j * ldb],
                   b20 = B[kk + 2 + (size_t)j * ldb], b30 = B[kk + 3 + (size_t)j * ldb];
      const double b01 = B[kk + (size_t)(j + 1) * ldb], b11 = B[kk + 1 + (size_t)(j + 1) * ldb],
                   b21 = B[kk + 2 + (size_t)(j + 1) * ldb], b31 = B[kk + 3 + (size_t)(j + 1) * ldb];
      const double b02 = B[kk + (size_t)(j + 2) * ldb], b12 = B[kk + 1 + (size_t)(j + 2) * ldb],
                   b22 = B[kk + 2 + (size_t)(j + 2) * ldb], b32 = B[kk + 3 + (size_t)(j + 2) * ldb];
      const double b03 = B[kk + (size_t)(j + 3) * ldb], b13 = B[kk + 1 + (size_t)(j + 3) * ldb],
                   b23 = B[kk + 2 + (size_t)(j + 3) * ldb], b33 = B[kk + 3 + (size_t)(j + 3) * ldb];
#pragma omp simd
      for (int i = 0; i < m; ++i) {
        const double x0 = a0[i], x1 = a1[i], x2 = a2[i], x3 = a3[i];
        c0[i] += x0 * b00 + x1 * b10 + x2 * b20 + x3 * b30;
        c1[i] += x0 * b01 + x1 * b11 + x2 * b21 + x3 * b31;
        c2[i] += x0 * b02 + x1 * b12 + x2 * b22 + x3 * b32;
        c3[i] += x0 * b03 + x1 * b13 + x2 * b23 + x3 * b33;
      }
    }
    for (; kk < k; ++kk) {
      const double* restrict a0 = A + (size_t)kk * lda;
      const double b0 = B[kk + (size_t)j * ldb], b1 = B[kk + (size_t)(j + 1) * ldb],
                   b2 = B[kk + (size_t)(j + 2) * ldb], b3 = B[kk + (size_t)(j + 3) * ldb];
#pragma omp simd
      for (int i = 0; i < m; ++i) {
        c0[i] += a0[i] * b0;
        c1[i] += a0[i] * b1;
        c2[i] += a0[i] * b2;
        c3[i] += a0[i] * b3;
      }
    }
  }
  for (; j < n; ++j) {
    double* restrict c0 = C + (size_t)j * ldc;
    for (int kk = 0; kk < k; ++kk) {
      const double* restrict a0 = A + (size_t)kk * lda;
      const double b0 = B[kk + (size_t)j * ldb];
#pragma omp simd
      for (int i = 0; i < m; ++i) c0[i] += a0[i] * b0;
    }
  }
}

/* C (m x n) = A^T (A is k x m) * B (k x n) */
static void gemm_tn_impl(int m, int n, int k, const double* restrict A, int lda, const double* restrict B, int ldb,
                    double* restrict C, int ldc) {
  if (g_dgemm && m > 0 && n > 0 && k > 0) {
    g_dgemm(ORC_COLMAJOR, ORC_TRANS, ORC_NOTRANS, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, ldc);
    return;
  }
  int j = 0;
  for (; j + 2 <= n; j += 2) {
    const double* restrict b0 = B + (size_t)j * ldb;
    const double* restrict b1 = b0 + ldb;
    int i = 0;
    for (; i + 4 <= m; i += 4) {
      const double* restrict a0 = A + (size_t)i * lda;
      const double* restrict a1 = a0 + lda;
      const double* restrict a2 = a1 + lda;
      const double* restrict a3 = a2 + lda;
      double s00 = 0, s10 = 0, s20 = 0, s30 = 0, s01 = 0, s11 = 0, s21 = 0, s31 = 0;
#pragma omp simd reduction(+ : s00, s10, s20, s30, s01, s11, s21, s31)
      for (int p = 0; p < k; ++p) {
        const double y0 = b0[p], y1 = b1[p];
        s00 += a0[p] * y0;
        s10 += a1[p] * y0;
        s20 += a2[p] * y0;
        s30 += a3[p] * y0;
        s01 += a0[p] * y1;
        s11 += a1[p] * y1;
        s21 += a2[p] * y1;
        s31 += a3[p] * y1;
      }
      C[i + (size_t)j * ldc] = s00;
      C[i + 1 + (size_t)j * ldc] = s10;
      C[i + 2 + (size_t)j * ldc] = s20;
      C[i + 3 + (size_t)j * ldc] = s30;
      C[i + (size_t)(j + 1) * ldc] = s01;
      C[i + 1 + (size_t)(j + 1) * ldc] = s11;
      C[i + 2 + (size_t)(j + 1) * ldc] = s21;
      C[i + 3 + (size_t)(j + 1) * ldc] = s31;
    }
    for (; i < m; ++i) {
      const double* restrict a0 = A + (size_t)i * lda;
      double s0 = 0, s1 = 0;
#pragma omp simd reduction(+ : s0, s1)
      for (int p = 0; p < k; ++p) {
        s0 += a0[p] * b0[p];
        s1 += a0[p] * b1[p];
      }
      C[i + (size_t)j * ldc] = s0;
      C[i + (size_t)(j + 1) * ldc] = s1;
    }
  }
  for (; j < n; ++j) {
    const double* restrict b0 = B + (size_t)j * ldb;
    for (int i = 0; i < m; ++i) {
      const double* restrict a0 = A + (size_t)i * lda;
      double s0 = 0;
#pragma omp simd reduction(+ : s0)
      for (int p = 0; p < k; ++p) s0 += a0[p] * b0[p];
      C[i + (size_t)j * ldc] = s0;
    }
  }
}

static void gemm_nn(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  const int probe = omp_get_thread_num() == 0;
  const double t0 = probe ? omp_get_wtime() : 0.0;
  gemm_nn_impl(m, n, k, A, lda, B, ldb, C, ldc);
  if (probe) {
    g_probe[1] += omp_get_wtime() - t0;
    g_probe[2] += 2.0 * m * (double)n * k;
  }
}
static void gemm_tn(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  const int probe = omp_get_thread_num() == 0;
  const double t0 = probe ? omp_get_wtime() : 0.0;
  gemm_tn_impl(m, n, k, A, lda, B, ldb, C, ldc);
  if (probe) {
    g_probe[1] += omp_get_wtime() - t0;
    g_probe[2] += 2.0 * m * (double)n * k;
  }
}

/* ------------------------------------------------------------------------------------------------
 * 8a-2  density (and derivatives) on the grid
 * ------------------------------------------------------------------------------------------------ */
void orc_density_on_grid(const orc_basis* b, const orc_grid* g, double radial_thr, const double* P, double* rho,
                         double* gx, double* gy, double* gz, double* hess6, int* nonneg) {
  const int nblocks = orc_nblocks(g);
  const int nbf = b->nbf;
  const int deriv = hess6 ? 2 : (gx ? 1 : 0);
  const long N = g->npts;
  /* clear outputs, MatrixOperatorToGridTransformer.cpp:70-95 */
  memset(rho, 0, sizeof(double) * (size_t)N);
  if (gx) {
    memset(gx, 0, sizeof(double) * (size_t)N);
    memset(gy, 0, sizeof(double) * (size_t)N);
    memset(gz, 0, sizeof(double) * (size_t)N);
  }
  if (hess6) memset(hess6, 0, sizeof(double) * 6 * (size_t)N);
#pragma omp parallel
  {
    orc_ws w;
    ws_alloc(&w, b, g->blocksize, deriv);
    double* cd2[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
    double* GM[3] = {NULL, NULL, NULL};
    if (deriv >= 2) {
      for (int k = 0; k < 6; ++k) cd2[k] = (double*)malloc(sizeof(double) * (size_t)g->blocksize * nbf);
      for (int k = 0; k < 3; ++k) GM[k] = (double*)malloc(sizeof(double) * (size_t)g->blocksize * nbf);
    }
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; ++blk) { /* MatrixOperatorToGridTransformer.cpp:103 */
      int n;
      const int s = ws_block(&w, b, g, radial_thr, deriv, blk, &n, NULL);
      const long first = (long)blk * g->blocksize;
      if (nonneg) nonneg[blk] = s > 0;
      if (s == 0) continue; /* :117-126 */
      /* phi_s = phi * Proj ; P_s = Proj^T P Proj (:137-157) */
      for (int j = 0; j < s; ++j) {
        memcpy(w.cval + (size_t)j * n, w.val + (size_t)w.sig[j] * n, sizeof(double) * (size_t)n);
        for (int k = 0; k < 3 && deriv >= 1; ++k)
          memcpy(w.cd1[k] + (size_t)j * n, w.d1[k] + (size_t)w.sig[j] * n, sizeof(double) * (size_t)n);
        for (int k = 0; k < 6 && deriv >= 2; ++k)
          memcpy(cd2[k] + (size_t)j * n, w.d2[k] + (size_t)w.sig[j] * n, sizeof(double) * (size_t)n);
        for (int i = 0; i < s; ++i) w.Ps[i + (size_t)j * s] = P[w.sig[i] + (size_t)w.sig[j] * nbf];
      }
      gemm_nn(n, s, s, w.cval, n, w.Ps, s, w.BP, n); /* basis_P = phi_s * P_s (:157) */
      for (int j = 0; j < s; ++j) {
        const double* bp = w.BP + (size_t)j * n;
        const double* f = w.cval + (size_t)j * n;
        for (int p = 0; p < n; ++p) rho[first + p] += bp[p] * f[p]; /* :158 */
        if (deriv >= 1) {
          const double *fx = w.cd1[0] + (size_t)j * n, *fy = w.cd1[1] + (size_t)j * n, *fz = w.cd1[2] + (size_t)j * n;
          for (int p = 0; p < n; ++p) {
            gx[first + p] += 2.0 * bp[p] * fx[p]; /* :161-163 */
            gy[first + p] += 2.0 * bp[p] * fy[p];
            gz[first + p] += 2.0 * bp[p] * fz[p];
          }
        }
      }
      if (deriv >= 2) { /* :166-188 */
        for (int k = 0; k < 3; ++k) gemm_nn(n, s, s, w.cd1[k], n, w.Ps, s, GM[k], n);
        static const int ci[6] = {0, 0, 0, 1, 1, 2}, cj[6] = {0, 1, 2, 1, 2, 2};
        for (int c = 0; c < 6; ++c) {
          double* h = hess6 + (size_t)c * N;
          for (int j = 0; j < s; ++j)
            for (int p = 0; p < n; ++p)
              h[first + p] += 2.0 * (w.BP[p + (size_t)j * n] * cd2[c][p + (size_t)j * n] +
                                     GM[ci[c]][p + (size_t)j * n] * w.cd1[cj[c]][p + (size_t)j * n]);
        }
      }
    }
    for (int k = 0; k < 6; ++k) free(cd2[k]);
    for (int k = 0; k < 3; ++k) free(GM[k]);
    ws_free(&w);
  }
}

/* ------------------------------------------------------------------------------------------------
 * 8a-5  scalar (+ gradient) operator on the grid -> matrix
 * ------------------------------------------------------------------------------------------------ */
void orc_scalar_to_matrix(const orc_basis* b, const orc_grid* g, double radial_thr, double block_ave_thr,
                          const double* v, const double* gx, const double* gy, const double* gz, double* V) {
  const int nblocks = orc_nblocks(g);
  const int nbf = b->nbf;
  const int gga = gx != NULL;
  const int nthreads = omp_get_max_threads();
  /* per-thread full nb x nb accumulators, ScalarOperatorToMatrixAdder.cpp:64 / :99 */
  double** acc = (double**)calloc((size_t)nthreads, sizeof(double*));
#pragma omp parallel
  {
    const int tid = omp_get_thread_num();
    acc[tid] = (double*)calloc((size_t)nbf * nbf, sizeof(double));
    orc_ws w;
    ws_alloc(&w, b, g->blocksize, gga ? 1 : 0);
    double* a = (double*)malloc(sizeof(double) * 4 * (size_t)g->blocksize);
    double *bx = a + g->blocksize, *by = bx + g->blocksize, *bz = by + g->blocksize;
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; ++blk) { /* :66 / :101 */
      int n;
      /* the reference asks for the block data before the average test (:69-70) */
      const int s = ws_block(&w, b, g, radial_thr, gga ? 1 : 0, blk, &n, NULL);
      const long first = (long)blk * g->blocksize;
      double ave = 0.0;
      for (int p = 0; p < n; ++p) {
        a[p] = g->w[first + p] * v[first + p];
        ave += fabs(a[p]);
      }
      if (gga) { /* :253-268 */
        double sx = 0, sy = 0, sz = 0;
        for (int p = 0; p < n; ++p) {
          bx[p] = g->w[first + p] * gx[first + p];
          by[p] = g->w[first + p] * gy[first + p];
          bz[p] = g->w[first + p] * gz[first + p];
          sx += fabs(bx[p]);
          sy += fabs(by[p]);
          sz += fabs(bz[p]);
        }
        ave += sx;
        ave += sy;
        ave += sz;
      }
      if (ave / n < block_ave_thr) continue; /* :203-206 / :267-268 */
      if (s == 0) continue;
      for (int j = 0; j < s; ++j) {
        const double* f = w.val + (size_t)w.sig[j] * n;
        double* cf = w.cval + (size_t)j * n;
        double* G = w.G + (size_t)j * n;
        memcpy(cf, f, sizeof(double) * (size_t)n);
        if (gga) { /* G = diag(bx) dx phi + diag(by) dy phi + diag(bz) dz phi + 0.5 diag(a) phi, :276-281 */
          const double *fx = w.d1[0] + (size_t)w.sig[j] * n, *fy = w.d1[1] + (size_t)w.sig[j] * n,
                       *fz = w.d1[2] + (size_t)w.sig[j] * n;
          for (int p = 0; p < n; ++p) G[p] = bx[p] * fx[p] + by[p] * fy[p] + bz[p] * fz[p] + 0.5 * a[p] * f[p];
        } else { /* scalA^T = diag(a) phi, :210 */
          for (int p = 0; p < n; ++p) G[p] = a[p] * f[p];
        }
      }
      gemm_tn(s, s, n, w.cval, n, w.G, n, w.T, s); /* T = phi_s^T G (:212 / :282) */
      double* A = acc[tid];
      if (gga) { /* V_s = T + T^T (:283-284), V += Proj V_s Proj^T (:301) */
        for (int j = 0; j < s; ++j)
          for (int i = 0; i < s; ++i)
            A[w.sig[i] + (size_t)w.sig[j] * nbf] += w.T[i + (size_t)j * s] + w.T[j + (size_t)i * s];
      } else {
        for (int j = 0; j < s; ++j)
          for (int i = 0; i < s; ++i) A[w.sig[i] + (size_t)w.sig[j] * nbf] += w.T[i + (size_t)j * s];
      }
    }
    free(a);
    ws_free(&w);
  }
  /* serial reduction over threads, :73-75 / :108-110 */
  for (int t = 0; t < nthreads; ++t) {
    if (!acc[t]) continue;
    for (size_t i = 0; i < (size_t)nbf * nbf; ++i) V[i] += acc[t][i];
    free(acc[t]);
  }
  free(acc);
}

/* ------------------------------------------------------------------------------------------------
 * 8a-6  FuncPotential::getMatrix
 * ------------------------------------------------------------------------------------------------ */
int orc_build_xc(const orc_basis* b, const orc_grid* g, const orc_functional* f, double radial_thr,
                 double block_ave_thr, const double* P, double* V, double* E, double* nelec, orc_timings* t) {
  const long N = g->npts;
  const int gga = orc_functional_is_gga(f);
  double* buf = (double*)malloc(sizeof(double) * 9 * (size_t)N);
  if (!buf) return -1;
  double *rho = buf, *gx = rho + N, *gy = gx + N, *gz = gy + N, *ep = gz + N, *vr = ep + N, *vx = vr + N,
         *vy = vx + N, *vz = vy + N;
  const double t0 = omp_get_wtime();
  /* NOTE: FuncPotential always builds a derivative-level-1 controller (BasisFunctionOnGridControllerFactory
   * default, :67) and XCFun::calcData only asks for the gradient when the functional is a GGA (XCFun.cpp:56-61);
   * the DensityMatrixDensityOnGridController of FuncPotential.cpp:61-62 is created with highestDerivative 1. */
  orc_density_on_grid(b, g, radial_thr, P, rho, gx, gy, gz, NULL, NULL);
  const double t1 = omp_get_wtime();
  const double e = orc_functional_on_grid(f, N, g->w, rho, gga ? gx : NULL, gga ? gy : NULL, gga ? gz : NULL, ep, vr,
                                          gga ? vx : NULL, gga ? vy : NULL, gga ? vz : NULL);
  const double t2 = omp_get_wtime();
  memset(V, 0, sizeof(double) * (size_t)b->nbf * b->nbf); /* FuncPotential.cpp:88-92 */
  if (f->ncomp > 0) orc_scalar_to_matrix(b, g, radial_thr, block_ave_thr, vr, gga ? vx : NULL, vy, vz, V);
  const double t3 = omp_get_wtime();
  double ne = 0.0;
  for (long p = 0; p < N; ++p) ne += rho[p] * g->w[p];
  *E = e;
  if (nelec) *nelec = ne;
  if (t) {
    t->basis_on_grid = 0.0;
    t->density_on_grid = t1 - t0;
    t->functional = t2 - t1;
    t->grid_to_matrix = t3 - t2;
    t->total = t3 - t0;
  }
  free(buf);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * f-4  two-basis scatter: ScalarOperatorToMatrixAdder::addBlock with basis A != basis B
 *      (ScalarOperatorToMatrixAdder.cpp:179-223 LDA, :225-303 GGA, branch !useSym :216-220 / :286-300), the operator of
 *      ABFuncPotential / ABNAddFuncPotential (potentials/ABFockMatrixConstruction/ABFuncPotential.cpp:54-160):
 *        m_AB += pA [ phi_A^T diag(a) phi_B + phi_A^T grad_B + grad_A^T phi_B ] pB^T,  grad_X = sum_c diag(b_c) d_c phi_X
 *      V is nbf_A x nbf_B column-major, added into.
 * ------------------------------------------------------------------------------------------------ */
void orc_scalar_to_matrix_ab(const orc_basis* bA, const orc_basis* bB, const orc_grid* g, double radial_thr,
                             double block_ave_thr, const double* v, const double* gx, const double* gy, const double* gz,
                             double* V) {
  const int nblocks = orc_nblocks(g);
  const int nA = bA->nbf, nB = bB->nbf;
  const int gga = gx != NULL;
  const int nthreads = omp_get_max_threads();
  double** acc = (double**)calloc((size_t)nthreads, sizeof(double*));
#pragma omp parallel
  {
    const int tid = omp_get_thread_num();
    acc[tid] = (double*)calloc((size_t)nA * nB, sizeof(double));
    orc_ws wA, wB;
    ws_alloc(&wA, bA, g->blocksize, gga ? 1 : 0);
    ws_alloc(&wB, bB, g->blocksize, gga ? 1 : 0);
    double* a = (double*)malloc(sizeof(double) * 4 * (size_t)g->blocksize);
    double *bx = a + g->blocksize, *by = bx + g->blocksize, *bz = by + g->blocksize;
    double* T = (double*)malloc(sizeof(double) * (size_t)nA * nB);
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; ++blk) {
      int n, n2;
      const int sA = ws_block(&wA, bA, g, radial_thr, gga ? 1 : 0, blk, &n, NULL);
      const int sB = ws_block(&wB, bB, g, radial_thr, gga ? 1 : 0, blk, &n2, NULL);
      const long first = (long)blk * g->blocksize;
      double ave = 0.0;
      for (int p = 0; p < n; ++p) {
        a[p] = g->w[first + p] * v[first + p];
        ave += fabs(a[p]);
      }
      if (gga) { /* :253-268 */
        double sx = 0, sy = 0, sz = 0;
        for (int p = 0; p < n; ++p) {
          bx[p] = g->w[first + p] * gx[first + p];
          by[p] = g->w[first + p] * gy[first + p];
          bz[p] = g->w[first + p] * gz[first + p];
          sx += fabs(bx[p]);
          sy += fabs(by[p]);
          sz += fabs(bz[p]);
        }
        ave += sx;
        ave += sy;
        ave += sz;
      }
      if (ave / n < block_ave_thr) continue;
      if (sA == 0 || sB == 0) continue;
      /* A side: phi_A and grad_A;  B side: phi_B and  a phi_B + grad_B  (so that one product pair gives all three terms) */
      for (int j = 0; j < sA; ++j) {
        const double* f = wA.val + (size_t)wA.sig[j] * n;
        memcpy(wA.cval + (size_t)j * n, f, sizeof(double) * (size_t)n);
        double* G = wA.G + (size_t)j * n;
        if (gga) {
          const double *fx = wA.d1[0] + (size_t)wA.sig[j] * n, *fy = wA.d1[1] + (size_t)wA.sig[j] * n,
                       *fz = wA.d1[2] + (size_t)wA.sig[j] * n;
          for (int p = 0; p < n; ++p) G[p] = bx[p] * fx[p] + by[p] * fy[p] + bz[p] * fz[p];
        }
      }
      for (int j = 0; j < sB; ++j) {
        const double* f = wB.val + (size_t)wB.sig[j] * n;
        memcpy(wB.cval + (size_t)j * n, f, sizeof(double) * (size_t)n);
        double* G = wB.G + (size_t)j * n;
        if (gga) {
          const double *fx = wB.d1[0] + (size_t)wB.sig[j] * n, *fy = wB.d1[1] + (size_t)wB.sig[j] * n,
                       *fz = wB.d1[2] + (size_t)wB.sig[j] * n;
          for (int p = 0; p < n; ++p) G[p] = a[p] * f[p] + bx[p] * fx[p] + by[p] * fy[p] + bz[p] * fz[p];
        } else {
          for (int p = 0; p < n; ++p) G[p] = a[p] * f[p];
        }
      }
      double* A = acc[tid];
      gemm_tn(sA, sB, n, wA.cval, n, wB.G, n, T, sA); /* phi_A^T (a phi_B + grad_B) */
      for (int j = 0; j < sB; ++j)
        for (int i = 0; i < sA; ++i) A[wA.sig[i] + (size_t)wB.sig[j] * nA] += T[i + (size_t)j * sA];
      if (gga) {
        gemm_tn(sA, sB, n, wA.G, n, wB.cval, n, T, sA); /* grad_A^T phi_B */
        for (int j = 0; j < sB; ++j)
          for (int i = 0; i < sA; ++i) A[wA.sig[i] + (size_t)wB.sig[j] * nA] += T[i + (size_t)j * sA];
      }
    }
    free(T);
    free(a);
    ws_free(&wA);
    ws_free(&wB);
  }
  for (int t = 0; t < nthreads; ++t) {
    if (!acc[t]) continue;
    for (size_t i = 0; i < (size_t)nA * nB; ++i) V[i] += acc[t][i];
    free(acc[t]);
  }
  free(acc);
}

/* ------------------------------------------------------------------------------------------------
 * 8a-7  NAddFuncPotential::getMatrix / getEnergy, 8a-3 SupersystemDensityOnGridController::updateData
 * ------------------------------------------------------------------------------------------------ */
int orc_build_nadd(const orc_basis* bA, const double* PA, int nenv, const orc_basis* const* bE,
                   const double* const* PE, const orc_grid* g, const orc_functional* f, double radial_thr,
                   double block_ave_thr, double* VA, double* E_nadd, double* E_parts) {
  const long N = g->npts;
  const int nblocks = orc_nblocks(g);
  const int gga = orc_functional_is_gga(f);
  const int nsub = 1 + nenv;
  double* dens = (double*)malloc(sizeof(double) * 4 * (size_t)N * (size_t)nsub);
  double* tot = (double*)calloc(4 * (size_t)N, sizeof(double));
  double* fd = (double*)malloc(sizeof(double) * 10 * (size_t)N);
  int* nonneg = (int*)malloc(sizeof(int) * (size_t)nblocks * (size_t)nsub);
  if (!dens || !tot || !fd || !nonneg) return -1;
  for (int i = 0; i < nsub; ++i) {
    double* d = dens + 4 * (size_t)N * i;
    orc_density_on_grid(i == 0 ? bA : bE[i - 1], g, radial_thr, i == 0 ? PA : PE[i - 1], d, d + N, d + 2 * N,
                        d + 3 * N, NULL, nonneg + (size_t)nblocks * i);
  }
  /* supersystem sum only over blocks flagged non-negligible per subsystem,
   * SupersystemDensityOnGridController.cpp:112-133, :156-191 */
#pragma omp parallel for schedule(dynamic)
  for (int blk = 0; blk < nblocks; ++blk) {
    const long first = (long)blk * g->blocksize;
    const int n = block_size(g, blk);
    for (int i = 0; i < nsub; ++i) {
      if (!nonneg[(size_t)nblocks * i + blk]) continue;
      const double* d = dens + 4 * (size_t)N * i;
      for (int c = 0; c < 4; ++c)
        for (int p = 0; p < n; ++p) tot[(size_t)c * N + first + p] += d[(size_t)c * N + first + p];
    }
  }
  double *ep = fd, *vr = ep + N, *vx = vr + N, *vy = vx + N, *vz = vy + N;
  double *ep2 = vz + N, *vr2 = ep2 + N, *vx2 = vr2 + N, *vy2 = vx2 + N, *vz2 = vy2 + N;
  /* superFuncDat, activeFuncDat: NAddFuncPotential.cpp:197-198 */
  const double e_tot = orc_functional_on_grid(f, N, g->w, tot, gga ? tot + N : NULL, tot + 2 * N, tot + 3 * N, ep, vr,
                                              gga ? vx : NULL, vy, vz);
  const double e_act = orc_functional_on_grid(f, N, g->w, dens, gga ? dens + N : NULL, dens + 2 * N, dens + 3 * N, ep2,
                                              vr2, gga ? vx2 : NULL, vy2, vz2);
  for (long p = 0; p < N; ++p) vr[p] -= vr2[p]; /* :217 / :221 */
  if (gga)
    for (long p = 0; p < N; ++p) {
      vx[p] -= vx2[p]; /* :222-224 */
      vy[p] -= vy2[p];
      vz[p] -= vz2[p];
    }
  memset(VA, 0, sizeof(double) * (size_t)bA->nbf * bA->nbf);
  if (f->ncomp > 0) orc_scalar_to_matrix(bA, g, radial_thr, block_ave_thr, vr, gga ? vx : NULL, vy, vz, VA);
  double e = e_tot - e_act; /* :249 */
  if (E_parts) {
    E_parts[0] = e_tot;
    E_parts[1] = e_act;
  }
  for (int i = 0; i < nenv; ++i) { /* NAddEnergyHelper::getEnergy, :502-516 (order 0, GRADIENT_INVARIANTS) */
    const double* d = dens + 4 * (size_t)N * (i + 1);
    const double ee = orc_functional_on_grid(f, N, g->w, d, gga ? d + N : NULL, d + 2 * N, d + 3 * N, ep2, vr2,
                                             gga ? vx2 : NULL, vy2, vz2);
    e -= ee;
    if (E_parts) E_parts[2 + i] = ee;
  }
  *E_nadd = e;
  free(dens);
  free(tot);
  free(fd);
  free(nonneg);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * UNRESTRICTED variants: the reference instantiates the same templates with SCFMode = UNRESTRICTED; every
 * per-spin quantity is a {alpha, beta} pair looped with for_spin (data/SpinPolarizedData.h:1880).
 * ------------------------------------------------------------------------------------------------ */
int orc_build_xc_u(const orc_basis* b, const orc_grid* g, const orc_functional* f, double radial_thr,
                   double block_ave_thr, const double* Pa, const double* Pb, double* Va, double* Vb, double* E,
                   double* nelec) {
  const long N = g->npts;
  const int gga = orc_functional_is_gga(f);
  double* buf = (double*)malloc(sizeof(double) * 17 * (size_t)N);
  if (!buf) return -1;
  double *rho = buf, *grad = rho + 2 * N, *ep = grad + 6 * N, *vr = ep + N, *vg = vr + 2 * N;
  /* MatrixOperatorToGridTransformer.h:128-146: the same block data is contracted with P_alpha and P_beta */
  orc_density_on_grid(b, g, radial_thr, Pa, rho, grad, grad + N, grad + 2 * N, NULL, NULL);
  orc_density_on_grid(b, g, radial_thr, Pb, rho + N, grad + 3 * N, grad + 4 * N, grad + 5 * N, NULL, NULL);
  *E = orc_functional_on_grid_u(f, N, g->w, rho, gga ? grad : NULL, ep, vr, gga ? vg : NULL);
  memset(Va, 0, sizeof(double) * (size_t)b->nbf * b->nbf);
  memset(Vb, 0, sizeof(double) * (size_t)b->nbf * b->nbf);
  if (f->ncomp > 0) {
    orc_scalar_to_matrix(b, g, radial_thr, block_ave_thr, vr, gga ? vg : NULL, vg + N, vg + 2 * N, Va);
    orc_scalar_to_matrix(b, g, radial_thr, block_ave_thr, vr + N, gga ? vg + 3 * N : NULL, vg + 4 * N, vg + 5 * N, Vb);
  }
  if (nelec) {
    double ne = 0.0;
    for (long p = 0; p < N; ++p) ne += (rho[p] + rho[N + p]) * g->w[p];
    *nelec = ne;
  }
  free(buf);
  return 0;
}

int orc_build_nadd_u(const orc_basis* bA, const double* PAa, const double* PAb, int nenv, const orc_basis* const* bE,
                     const double* const* PEa, const double* const* PEb, const orc_grid* g, const orc_functional* f,
                     double radial_thr, double block_ave_thr, double* VAa, double* VAb, double* E_nadd, double* E_parts) {
  const long N = g->npts;
  const int nblocks = orc_nblocks(g);
  const int gga = orc_functional_is_gga(f);
  const int nsub = 1 + nenv;
  /* per subsystem: rho[2][N] then grad[2][3][N] */
  double* dens = (double*)malloc(sizeof(double) * 8 * (size_t)N * (size_t)nsub);
  double* tot = (double*)calloc(8 * (size_t)N, sizeof(double));
  double* fd = (double*)malloc(sizeof(double) * 18 * (size_t)N);
  int* nonneg = (int*)malloc(sizeof(int) * (size_t)nblocks * (size_t)nsub);
  if (!dens || !tot || !fd || !nonneg) return -1;
  for (int i = 0; i < nsub; ++i) {
    double* d = dens + 8 * (size_t)N * i;
    const orc_basis* bb = i == 0 ? bA : bE[i - 1];
    orc_density_on_grid(bb, g, radial_thr, i == 0 ? PAa : PEa[i - 1], d, d + 2 * N, d + 3 * N, d + 4 * N, NULL,
                        nonneg + (size_t)nblocks * i);
    orc_density_on_grid(bb, g, radial_thr, i == 0 ? PAb : PEb[i - 1], d + N, d + 5 * N, d + 6 * N, d + 7 * N, NULL,
                        nonneg + (size_t)nblocks * i); /* same flags: they depend on the basis only */
  }
#pragma omp parallel for schedule(dynamic)
  for (int blk = 0; blk < nblocks; ++blk) { /* SupersystemDensityOnGridController.cpp:196-307 */
    const long first = (long)blk * g->blocksize;
    const int n = block_size(g, blk);
    for (int i = 0; i < nsub; ++i) {
      if (!nonneg[(size_t)nblocks * i + blk]) continue;
      const double* d = dens + 8 * (size_t)N * i;
      for (int c = 0; c < 8; ++c)
        for (int p = 0; p < n; ++p) tot[(size_t)c * N + first + p] += d[(size_t)c * N + first + p];
    }
  }
  double *ep = fd, *vr = ep + N, *vg = vr + 2 * N, *ep2 = vg + 6 * N, *vr2 = ep2 + N, *vg2 = vr2 + 2 * N;
  const double e_tot = orc_functional_on_grid_u(f, N, g->w, tot, gga ? tot + 2 * N : NULL, ep, vr, gga ? vg : NULL);
  const double e_act = orc_functional_on_grid_u(f, N, g->w, dens, gga ? dens + 2 * N : NULL, ep2, vr2, gga ? vg2 : NULL);
  for (long p = 0; p < 2 * N; ++p) vr[p] -= vr2[p];
  if (gga)
    for (long p = 0; p < 6 * N; ++p) vg[p] -= vg2[p];
  memset(VAa, 0, sizeof(double) * (size_t)bA->nbf * bA->nbf);
  memset(VAb, 0, sizeof(double) * (size_t)bA->nbf * bA->nbf);
  if (f->ncomp > 0) {
    orc_scalar_to_matrix(bA, g, radial_thr, block_ave_thr, vr, gga ? vg : NULL, vg + N, vg + 2 * N, VAa);
    orc_scalar_to_matrix(bA, g, radial_thr, block_ave_thr, vr + N, gga ? vg + 3 * N : NULL, vg + 4 * N, vg + 5 * N, VAb);
  }
  double e = e_tot - e_act;
  if (E_parts) {
    E_parts[0] = e_tot;
    E_parts[1] = e_act;
  }
  for (int i = 0; i < nenv; ++i) {
    const double* d = dens + 8 * (size_t)N * (i + 1);
    const double ee = orc_functional_on_grid_u(f, N, g->w, d, gga ? d + 2 * N : NULL, ep2, vr2, gga ? vg2 : NULL);
    e -= ee;
    if (E_parts) E_parts[2 + i] = ee;
  }
  *E_nadd = e;
  free(dens);
  free(tot);
  free(fd);
  free(nonneg);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * row f-3  FuncPotential::getGeomGradients, potentials/FuncPotential.cpp:114-239
 * ------------------------------------------------------------------------------------------------ */
/* the O(s^2 n) double loop of FuncPotential.cpp:145-233 / NAddFuncPotential.cpp:383-490 for a given potential on the grid:
 * vr [nspin][N], vg [nspin][3][N] */
static void xc_gradient_contract(const orc_basis* b, const orc_grid* g, double radial_thr, int nspin, const double* const* Ps,
                                 const double* vr, const double* vg, int gga, int natoms, const int* atom_of_bf,
                                 double* grad) {
  const long N = g->npts;
  const int nbf = b->nbf;
  const int nblocks = orc_nblocks(g);
  memset(grad, 0, sizeof(double) * 3 * (size_t)natoms);
#pragma omp parallel
  {
    orc_ws w;
    ws_alloc(&w, b, g->blocksize, gga ? 2 : 1);
    double* priv = (double*)calloc(3 * (size_t)natoms, sizeof(double)); /* gradientContrPriv, :149 */
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; ++blk) {
      const long first = (long)blk * g->blocksize;
      int n = 0;
      const int s = ws_block(&w, b, g, radial_thr, gga ? 2 : 1, blk, &n, NULL);
      for (int sp = 0; sp < nspin; ++sp) {
        const double* P = Ps[sp];
        const double* v = vr + (size_t)sp * N + first;
        const double* vx = vg + (size_t)(3 * sp) * N + first;
        const double* vy = vg + (size_t)(3 * sp + 1) * N + first;
        const double* vz = vg + (size_t)(3 * sp + 2) * N + first;
        const double* wt = g->w + first;
        for (int im = 0; im < s; ++im) { /* negligible functions are skipped, :170 */
          const int mu = w.sig[im];
          const double* fm = w.val + (size_t)mu * n;
          const double *mx = w.d1[0] + (size_t)mu * n, *my = w.d1[1] + (size_t)mu * n, *mz = w.d1[2] + (size_t)mu * n;
          for (int in = 0; in < s; ++in) { /* :178 */
            const int nu = w.sig[in];
            const int A = atom_of_bf[nu];
            const double pre = 2.0 * P[mu + (size_t)nu * nbf]; /* :183 */
            const double *nx = w.d1[0] + (size_t)nu * n, *ny = w.d1[1] + (size_t)nu * n, *nz = w.d1[2] + (size_t)nu * n;
            double gx = 0.0, gy = 0.0, gz = 0.0;
            for (int p = 0; p < n; ++p) { /* LDA-type part, :184-186 */
              const double pm = wt[p] * v[p] * fm[p];
              gx += pm * nx[p];
              gy += pm * ny[p];
              gz += pm * nz[p];
            }
            if (gga) { /* :188-227 */
              const double *hxx = w.d2[0] + (size_t)nu * n, *hxy = w.d2[1] + (size_t)nu * n, *hxz = w.d2[2] + (size_t)nu * n,
                           *hyy = w.d2[3] + (size_t)nu * n, *hyz = w.d2[4] + (size_t)nu * n, *hzz = w.d2[5] + (size_t)nu * n;
              for (int p = 0; p < n; ++p) {
                const double px = wt[p] * vx[p], py = wt[p] * vy[p], pz = wt[p] * vz[p];
                gx += px * (fm[p] * hxx[p] + mx[p] * nx[p]) + py * (fm[p] * hxy[p] + my[p] * nx[p]) +
                      pz * (fm[p] * hxz[p] + mz[p] * nx[p]);
                gy += px * (fm[p] * hxy[p] + mx[p] * ny[p]) + py * (fm[p] * hyy[p] + my[p] * ny[p]) +
                      pz * (fm[p] * hyz[p] + mz[p] * ny[p]);
                gz += px * (fm[p] * hxz[p] + mx[p] * nz[p]) + py * (fm[p] * hyz[p] + my[p] * nz[p]) +
                      pz * (fm[p] * hzz[p] + mz[p] * nz[p]);
              }
            }
            priv[A] -= pre * gx;
            priv[A + natoms] -= pre * gy;
            priv[A + 2 * natoms] -= pre * gz;
          }
        }
      }
    }
#pragma omp critical
    for (int i = 0; i < 3 * natoms; ++i) grad[i] += priv[i]; /* :231-232 */
    free(priv);
    ws_free(&w);
  }
}

int orc_xc_gradient(const orc_basis* b, const orc_grid* g, const orc_functional* f, double radial_thr, int nspin,
                    const double* Pa, const double* Pb, int natoms, const int* atom_of_bf, double* grad) {
  const long N = g->npts;
  const int gga = orc_functional_is_gga(f);
  /* rho[2][N], grad[2][3][N], epuv[N], vr[2][N], vg[2][3][N] (restricted uses the first halves) */
  double* buf = (double*)calloc(17 * (size_t)N, sizeof(double));
  if (!buf) return -1;
  double *rho = buf, *gr = rho + 2 * N, *ep = gr + 6 * N, *vr = ep + N, *vg = vr + 2 * N;
  const double* Ps[2] = {Pa, Pb};
  if (nspin == 1) {
    orc_density_on_grid(b, g, radial_thr, Pa, rho, gr, gr + N, gr + 2 * N, NULL, NULL);
    orc_functional_on_grid(f, N, g->w, rho, gga ? gr : NULL, gr + N, gr + 2 * N, ep, vr, gga ? vg : NULL, vg + N, vg + 2 * N);
  } else {
    orc_density_on_grid(b, g, radial_thr, Pa, rho, gr, gr + N, gr + 2 * N, NULL, NULL);
    orc_density_on_grid(b, g, radial_thr, Pb, rho + N, gr + 3 * N, gr + 4 * N, gr + 5 * N, NULL, NULL);
    orc_functional_on_grid_u(f, N, g->w, rho, gga ? gr : NULL, ep, vr, gga ? vg : NULL);
  }
  xc_gradient_contract(b, g, radial_thr, nspin, Ps, vr, vg, gga, natoms, atom_of_bf, grad);
  free(buf);
  return 0;
}

/* NAddFuncPotential::getGeomGradients (NAddFuncPotential.cpp:329-493), RESTRICTED: v = v[rho_act + sum rho_env] - v[rho_act]
 * (:331-341) contracted with the ACTIVE density matrix over the active system's basis functions and atoms. */
int orc_nadd_gradient(const orc_basis* bA, const double* PA, int nenv, const orc_basis* const* bE, const double* const* PE,
                      const orc_grid* g, const orc_functional* f, double radial_thr, int natoms, const int* atom_of_bf,
                      double* grad) {
  const long N = g->npts;
  const int gga = orc_functional_is_gga(f);
  /* act[4][N], tot[4][N], tmp[4][N], ep[N], vt[4][N], va[4][N] */
  double* buf = (double*)calloc(21 * (size_t)N, sizeof(double));
  if (!buf) return -1;
  double *act = buf, *tot = act + 4 * N, *tmp = tot + 4 * N, *ep = tmp + 4 * N, *vt = ep + N, *va = vt + 4 * N;
  orc_density_on_grid(bA, g, radial_thr, PA, act, act + N, act + 2 * N, act + 3 * N, NULL, NULL);
  memcpy(tot, act, sizeof(double) * 4 * (size_t)N);
  for (int i = 0; i < nenv; ++i) {
    orc_density_on_grid(bE[i], g, radial_thr, PE[i], tmp, tmp + N, tmp + 2 * N, tmp + 3 * N, NULL, NULL);
    for (size_t k = 0; k < 4 * (size_t)N; ++k) tot[k] += tmp[k];
  }
  orc_functional_on_grid(f, N, g->w, tot, gga ? tot + N : NULL, tot + 2 * N, tot + 3 * N, ep, vt, gga ? vt + N : NULL,
                         vt + 2 * N, vt + 3 * N);
  orc_functional_on_grid(f, N, g->w, act, gga ? act + N : NULL, act + 2 * N, act + 3 * N, ep, va, gga ? va + N : NULL,
                         va + 2 * N, va + 3 * N);
  for (size_t k = 0; k < 4 * (size_t)N; ++k) vt[k] -= va[k];
  const double* Ps[2] = {PA, NULL};
  xc_gradient_contract(bA, g, radial_thr, 1, Ps, vt, vt + N, gga, natoms, atom_of_bf, grad);
  free(buf);
  return 0;
}

