// grid_kernels.cuh - molecular partition weights of the atom-centred integration grid.
//
// Reference: the weight step of GridFactory::produce (src/grid/construction/GridFactory.cpp:139-266): for every point r of
// parent atom k,  w(r) = w_atomic(r) P_k(r) / sum_l P_l(r)  with the cell functions P_l = prod_{j != l} s(nu_lj),
//   BECKE (J. Chem. Phys. 88 (1988) 2547): nu = mu + a_lj (1 - mu^2), s = max(1, smoothing)-fold iterated (3 - x^2) x / 2 (:324-347),
//          over the "significant" atoms (closer than 40 bohr to the parent, :152-157);
//   VORONOI: w = 0 as soon as one nu_lk < 0 (:194-203);
//   SSF   (Chem. Phys. Lett. 257 (1996) 213): Eq. (15) sphere screening (:213), w = 0 if some nu_kj >= 0.64 (:216-225),
//          s(nu) = 1 / polynomial / 0 for nu <= -0.64 / inside / >= 0.64 with the early exits of :232-250.
// The reference is an O(N n_atoms^2) scalar loop nest on the host and dominates set-up for configs 3 and 5.
// B200 design: one warp per grid point.  The lanes compute the n_atoms point-atom distances into a per-warp shared
// array, run the nu_kj screen with a ballot, then take the atoms l round-robin: every lane walks j in the reference's
// order with the reference's early exits (so each P_l is the same product in the same order) and the warp sums the
// cells with shuffles.  SSF cells that a single test against the atom nearest to the point proves to be exactly zero
// skip the j loop (most atoms of a large molecule).  Atom-atom distances (n_atoms^2 doubles) are read through L1/L2.
#pragma once

#include "sxc_common.cuh"

namespace sxc {

constexpr int PW_WARPS = 8;

__device__ __forceinline__ double becke_smooth(double nu, int k) {
  for (int i = 0; i < k; ++i) nu = (3.0 - nu * nu) * nu / 2.0;
  return 0.5 * (1.0 - nu);
}

__global__ void __launch_bounds__(PW_WARPS * 32)
k_partition_weights(int flavour, int smooth_k, int natoms, const double* __restrict__ coords, const double* __restrict__ adist,
                    const double* __restrict__ aij, const double* __restrict__ min_dist, long npts,
                    const double* __restrict__ xyz, const int* __restrict__ parent, double* __restrict__ w) {
  extern __shared__ double rd_all[];  // [PW_WARPS][natoms]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* rd = rd_all + (size_t)warp * natoms;
  const long stride = (long)gridDim.x * PW_WARPS;
  for (long p = (long)blockIdx.x * PW_WARPS + warp; p < npts; p += stride) {
    const double x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
    const int k = parent[p];
    __syncwarp();
    for (int a = lane; a < natoms; a += 32) {
      const double dx = coords[3 * a] - x, dy = coords[3 * a + 1] - y, dz = coords[3 * a + 2] - z;
      rd[a] = sqrt(dx * dx + dy * dy + dz * dz);
    }
    __syncwarp();
    if (natoms == 1) continue;
    double weight = w[p];
    double sum = 0.0, cell_k = 0.0;
    if (flavour == 0) {  // BECKE
      for (int l = lane; l < natoms; l += 32) {
        if (adist[l + natoms * k] >= 40.0) continue;
        double cell = 1.0;
        for (int j = 0; j < natoms; ++j) {
          if (l == j || adist[j + natoms * k] >= 40.0) continue;
          const double mu = (rd[l] - rd[j]) / adist[l + natoms * j];
          const double nu = mu + (aij ? aij[j + natoms * l] : 0.0) * (1.0 - mu * mu);
          cell *= becke_smooth(nu, smooth_k);
        }
        if (l == k) cell_k = cell;
        sum += cell;
      }
    } else if (flavour == 2) {  // VORONOI: the point survives with its atomic weight iff no nu_lk is negative
      bool cut = false;
      for (int l = lane; l < natoms; l += 32) {
        if (adist[l + natoms * k] >= 40.0) continue;
        const double mu = (rd[l] - rd[k]) / adist[l + natoms * k];
        const double nu = mu + (aij ? aij[k + natoms * l] : 0.0) * (1.0 - mu * mu);
        if (nu < 0.0) cut = true;
      }
      if (__any_sync(0xffffffffu, cut) && lane == 0) w[p] = 0.0;
      continue;
    } else {  // SSF
      const double rk = rd[k];
      if (!(rk >= 0.5 * (1.0 - 0.64) * min_dist[k])) continue;  // Eq. (15): the atomic weight stands
      bool out = false;
      for (int j = lane; j < natoms; j += 32)
        if (j != k && (rk - rd[j]) / adist[k + natoms * j] >= 0.64) out = true;
      if (__any_sync(0xffffffffu, out)) {
        if (lane == 0) w[p] = 0.0;
        continue;
      }
      // nearest atom to the point: one test against it settles most cells (P_i = 0 exactly as soon as ANY nu_ij >= 0.64,
      // whichever j the reference's loop meets first), so only the few atoms around the point walk the full j loop
      double dmin = 1.0e300;
      int amin = 0;
      for (int a = lane; a < natoms; a += 32)
        if (rd[a] < dmin) {
          dmin = rd[a];
          amin = a;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, dmin, o);
        const int oa = __shfl_xor_sync(0xffffffffu, amin, o);
        if (od < dmin || (od == dmin && oa < amin)) {
          dmin = od;
          amin = oa;
        }
      }
      for (int i = lane; i < natoms; i += 32) {
        const double ri = rd[i];
        if (i != amin && (ri - dmin) / adist[i + natoms * amin] >= 0.64) continue;  // P_i = 0
        double cell = 1.0;
        for (int j = 0; j < natoms; ++j) {
          if (i == j) continue;
          double nu = (ri - rd[j]) / adist[i + natoms * j];
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
        if (i == k) cell_k = cell;
        sum += cell;
      }
    }
    sum = warp_sum(sum);
    cell_k = warp_sum(cell_k);  // exactly one lane holds a non-zero value
    if (lane == 0) w[p] = weight * cell_k / sum;
  }
}

}  // namespace sxc
