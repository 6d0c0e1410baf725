/*
 * gridweights.c - host helper of the synthetic-input producer: molecular partition weights.
 *
 * Input producer, not the hot path.  Restates the weight step of
 * src/grid/construction/GridFactory.cpp:139-266 (Becke: J. Chem. Phys. 88 (1988) 2547;
 * SSF: Stratmann/Scuseria/Frisch, Chem. Phys. Lett. 257 (1996) 213) for one parent atom's points.
 *
 *   flavour 0 = BECKE (with Bragg-Slater size adjustment a_ij passed in), 1 = SSF, 2 = VORONOI (:194-203)
 *
 * pts:  [3*n] points of parent atom k (already shifted to the molecule frame)
 * w:    [n] in: atomic quadrature weights, out: molecular weights (0 where screened out)
 */
#include <math.h>
#include <stdlib.h>

static double becke_smooth(double nu, int smoothing) {
  /* k-fold iterated p(x) = (3 - x^2) x / 2, s = 0.5 (1 - p_k); k = max(1, smoothing)  (GridFactory.cpp:324-347) */
  const int k = smoothing > 1 ? smoothing : 1;
  for (int i = 0; i < k; ++i) nu = (3.0 - nu * nu) * nu / 2.0;
  return 0.5 * (1.0 - nu);
}

void sxc_partition_weights(int flavour, int smoothing, int natoms, const double* coords /*[3*natoms]*/,
                           const double* adist /*[natoms*natoms]*/, const double* aij /*[natoms*natoms] or NULL*/,
                           int k, long n, const double* pts, double* w) {
  double min_dist = 999999999.9;
  for (int l = 0; l < natoms; ++l)
    if (l != k && adist[l + natoms * k] < min_dist) min_dist = adist[l + natoms * k];
#pragma omp parallel
  {
    double* rd = (double*)malloc(sizeof(double) * (size_t)natoms);
#pragma omp for schedule(dynamic, 64)
    for (long p = 0; p < n; ++p) {
      const double x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
      for (int a = 0; a < natoms; ++a) {
        const double dx = coords[3 * a] - x, dy = coords[3 * a + 1] - y, dz = coords[3 * a + 2] - z;
        rd[a] = sqrt(dx * dx + dy * dy + dz * dz);
      }
      double weight = w[p];
      double sum = 0.0;
      if (natoms == 1) continue;
      if (flavour == 0) {
        for (int l = 0; l < natoms; ++l) {
          if (adist[l + natoms * k] >= 40.0) continue; /* significantAtoms, GridFactory.cpp:152-157 */
          double cell = 1.0;
          for (int j = 0; j < natoms; ++j) {
            if (l == j || adist[j + natoms * k] >= 40.0) continue;
            const double mu = (rd[l] - rd[j]) / adist[l + natoms * j];
            const double nu = mu + (aij ? aij[j + natoms * l] : 0.0) * (1.0 - mu * mu);
            cell *= becke_smooth(nu, smoothing);
          }
          if (l == k) weight *= cell;
          sum += cell;
        }
        weight /= sum;
      } else if (flavour == 2) {
        for (int l = 0; l < natoms; ++l) {
          if (adist[l + natoms * k] >= 40.0) continue;
          const double mu = (rd[l] - rd[k]) / adist[l + natoms * k];
          const double nu = mu + (aij ? aij[k + natoms * l] : 0.0) * (1.0 - mu * mu);
          if (nu < 0.0) {
            weight = 0.0;
            break;
          }
        }
      } else {
        /* SSF, GridFactory.cpp:211-263 */
        if (rd[k] >= 0.5 * (1.0 - 0.64) * min_dist) {
          int done = 0;
          for (int j = 0; j < natoms; ++j) {
            if (j == k) continue;
            const double nu = (rd[k] - rd[j]) / adist[k + natoms * j];
            if (nu >= 0.64) {
              done = 1;
              break;
            }
          }
          if (done) {
            w[p] = 0.0;
            continue;
          }
          for (int i = 0; i < natoms; ++i) {
            double cell = 1.0;
            for (int j = 0; j < natoms; ++j) {
              if (i == j) continue;
              double nu = (rd[i] - rd[j]) / adist[i + natoms * j];
              if (nu <= -0.64) continue;
              if (nu >= 0.64) {
                cell = 0.0;
                break;
              }
              nu /= 0.64;
              const double n3 = nu * nu * nu;
              const double poly = (-5.0 * n3 * n3 * nu + 21.0 * n3 * nu * nu - 35.0 * n3 + 35.0 * nu) / 16.0;
              cell *= 0.5 * (1.0 - poly);
            }
            if (i == k) weight *= cell;
            sum += cell;
          }
          weight /= sum;
        }
      }
      w[p] = weight;
    }
    free(rd);
  }
}
