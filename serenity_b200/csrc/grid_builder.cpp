// grid_builder.cpp - molecular integration grid behind the C ABI (the part that was Python in round 1).
//
//   AtomGridFactory::produce            src/grid/construction/AtomGridFactory.cpp:78-198   radial count, pruning zones, Lebedev shells
//   AtomGridFactory::_radialGrid        :200-255   Becke / Ahlrichs (M3, exponent 0.6, Chebyshev 2nd kind) radial maps
//   GridFactory::produce                src/grid/construction/GridFactory.cpp:52-321       atom grids shifted to the nuclei, partition
//                                       weights (-> sxc_partition_weights on the device), weight cut (:264), Hilbert sort (:287)
//   HilbertRTreeSorting::sort           src/grid/HilbertRTreeSorting.cpp:29-214             integer coordinates, two lookup tables,
//                                       points in DESCENDING index order
// Host C++ except for the O(N n_atoms^2) weight step, which is the device kernel k_partition_weights.  The Lebedev rules come from
// the generated table lebedev_gen.h (the reference ships Burkardt's sphere_lebedev_rule.cpp: the same rules in another point
// order).  Element data: Treutler-Ahlrichs alpha values and Clementi radii for H..Kr (J. Chem. Phys. 102, 346 (1995), Table; J.
// Chem. Phys. 47, 1300 (1967)), Bragg-Slater radii (J. Chem. Phys. 41, 3199 (1964)) for the BECKE flavour where Slater lists one.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/serenity_xc_b200.h"
#include "lebedev_gen.h"

namespace {

constexpr int ZMAX = 36;
// index = nuclear charge (entry 0: dummy atom = hydrogen-like), AtomGridFactory.cpp:39-63
const double AHLRICHS_ALPHA[ZMAX + 1] = {1.1, 0.8, 0.9, 1.8, 1.4, 1.3, 1.1, 0.9, 0.9, 0.9, 0.9, 1.4, 1.3, 1.3, 1.2, 1.1, 1.0, 1.0, 1.0,
                                         1.5, 1.4, 1.3, 1.2, 1.2, 1.2, 1.2, 1.2, 1.2, 1.1, 1.1, 1.1, 1.1, 1.0, 0.9, 0.9, 0.9, 0.9};
const double CLEMENTI[ZMAX + 1] = {1.00, 1.00, 0.59, 3.16, 2.12, 1.64, 1.27, 1.06, 0.91, 0.79, 0.72, 3.59, 2.74, 2.23, 2.10, 1.85, 1.66,
                                   1.49, 1.34, 4.59, 3.67, 3.48, 3.33, 3.23, 3.14, 3.04, 2.95, 2.87, 2.82, 2.74, 2.68, 2.57, 2.36, 2.15,
                                   1.95, 1.78, 1.66};
// Angstrom; 0 = no value in Slater's table (noble gases)
const double BRAGG_SLATER[ZMAX + 1] = {0.0,  0.25, 0.0,  1.45, 1.05, 0.85, 0.70, 0.65, 0.60, 0.50, 0.0,  1.80, 1.50, 1.25, 1.10, 1.00, 1.00, 1.00, 0.0,
                                       2.20, 1.80, 1.60, 1.40, 1.35, 1.40, 1.40, 1.40, 1.35, 1.35, 1.35, 1.35, 1.30, 1.25, 1.15, 1.15, 1.15, 0.0};
constexpr double ANGSTROM_TO_BOHR = 1.0 / (5.29177210544e-11 * 1.0e10);  // src/parameters/Constants.h:78-81
const int RAD_ACC[7] = {13, 13, 13, 14, 15, 16, 17};
const int LDVAL[7][5] = {{4, 4, 4, 4, 4}, {4, 4, 4, 7, 4}, {4, 4, 7, 10, 7}, {4, 7, 10, 13, 10}, {7, 10, 13, 15, 13},
                         {10, 13, 15, 16, 15}, {13, 15, 16, 17, 16}};
const double RANGES[3][4] = {{0.25, 0.5, 1.0, 4.5}, {0.1667, 0.5, 0.9, 3.5}, {0.1, 0.4, 0.8, 2.5}};

thread_local std::string g_error;
int gfail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

int row_of(int z) { return z <= 2 ? 1 : z <= 10 ? 2 : z <= 18 ? 3 : 4; }

// AtomGridFactory.cpp:236-255 (Ahlrichs) and :203-212 (Becke); points with increasing radius (radPoints[nRadial - i])
void radial_grid(int radial, double alpha, int n, std::vector<double>& r, std::vector<double>& w) {
  r.assign(n, 0.0);
  w.assign(n, 0.0);
  const double pi = 3.14159265358979323846;
  for (int i = 1; i <= n; ++i) {
    const double xi = std::cos(i * pi / (n + 1.0));
    double ri, wi;
    if (radial == 0) {
      const double tmp = alpha / std::log(2.0);
      ri = tmp * std::pow(xi + 1.0, 0.6) * std::log(2.0 / (1.0 - xi));
      const double ln = std::log((1.0 - xi) / 2.0);
      const double sq = std::sqrt((1.0 + xi) / (1.0 - xi));
      wi = (pi / (n + 1.0)) * std::pow(1.0 + xi, 1.8) * tmp * tmp * tmp * (sq * ln * ln - 0.6 * ln * ln * ln / sq);
    } else {
      wi = std::sqrt(std::pow(1.0 + xi, 5) / std::pow(1.0 - xi, 7)) * (2.0 * pi) * alpha * alpha * alpha / (n + 1);
      ri = alpha * (1.0 + xi) / (1.0 - xi);
    }
    r[n - i] = ri;
    w[n - i] = wi;
  }
}

struct AtomGrid {
  std::vector<double> xyz, w;  // points relative to the nucleus (interleaved), weights incl. the radial quadrature
};

int atom_grid(int z, int acc, int radial, AtomGrid& out) {
  if (z < 1 || z > ZMAX) return gfail(SXC_ERR_UNSUPPORTED, "atom grids are tabulated for H..Kr (nuclear charge " + std::to_string(z) + ")");
  if (acc < 1 || acc > 7) return gfail(SXC_ERR_INVALID, "grid accuracy must be 1..7");
  const int row = row_of(z);
  const int nrad = (int)(5.0 * (RAD_ACC[acc - 1] + row - 8));
  double alpha = AHLRICHS_ALPHA[z];
  if (radial != 0) {
    if (BRAGG_SLATER[z] == 0.0) return gfail(SXC_ERR_UNSUPPORTED, "no Bragg-Slater radius for nuclear charge " + std::to_string(z));
    const double bs = BRAGG_SLATER[z] * ANGSTROM_TO_BOHR;
    alpha = z == 1 ? bs : 0.5 * bs;
  }
  std::vector<double> rp, rw;
  radial_grid(radial, alpha, nrad, rp, rw);
  const int zone_row = row > 3 ? 2 : row - 1;
  const int redp1 = (row == 1 && acc > 1) ? 2 : 1;
  double zones[4];
  for (int k = 0; k < 4; ++k) zones[k] = RANGES[zone_row][k] * CLEMENTI[z];
  const double four_pi = 4.0 * 3.14159265358979323846;
  int sph = 0;
  out.xyz.clear();
  out.w.clear();
  for (int i = 0; i < nrad; ++i) {
    if (sph < 4 && rp[i] > zones[sph]) ++sph;
    const int rule = LDVAL[acc - redp1][sph];
    const int off = sxc::lebedev::OFFSET[rule], np = sxc::lebedev::NPOINTS[rule];
    for (int k = 0; k < np; ++k) {
      const double* p = sxc::lebedev::POINTS[off + k];
      out.xyz.push_back(p[0] * rp[i]);
      out.xyz.push_back(p[1] * rp[i]);
      out.xyz.push_back(p[2] * rp[i]);
      out.w.push_back(p[3] * rw[i] * four_pi);
    }
  }
  return SXC_OK;
}

// HilbertRTreeSorting.cpp:57-78: cube transformation to the next level of depth, number of a point inside the first cube
const int HRT_TRANS[8][8] = {{0, 7, 6, 1, 2, 5, 4, 3}, {0, 3, 4, 6, 7, 5, 2, 1}, {0, 3, 4, 6, 7, 5, 2, 1}, {2, 3, 0, 1, 6, 7, 4, 5},
                             {2, 3, 0, 1, 6, 7, 4, 5}, {6, 5, 2, 1, 0, 3, 4, 7}, {6, 5, 2, 1, 0, 3, 4, 7}, {4, 3, 2, 5, 6, 1, 0, 7}};
const int HRT_VAL[2][2][2] = {{{5, 6}, {4, 7}}, {{2, 1}, {3, 0}}};

int hilbert_rtree_order(const std::vector<double>& xyz, std::vector<int64_t>& order) {
  const int64_t n = (int64_t)(xyz.size() / 3);
  order.resize(n);
  std::iota(order.begin(), order.end(), (int64_t)0);
  if (n < 2) return SXC_OK;
  int depth = 1;
  int64_t length = 2, nvert = 8;
  while (n * 8 > nvert) {  // :32-39
    ++depth;
    length *= 2;
    nvert *= 8;
  }
  if (depth > 20) return gfail(SXC_ERR_UNSUPPORTED, "grid too large for a 63-bit Hilbert index");
  double lo[3] = {xyz[0], xyz[1], xyz[2]}, hi[3] = {xyz[0], xyz[1], xyz[2]};
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) {
      lo[c] = std::min(lo[c], xyz[3 * i + c]);
      hi[c] = std::max(hi[c], xyz[3 * i + c]);
    }
  double spread[3];
  for (int c = 0; c < 3; ++c) spread[c] = (double)length / (hi[c] - lo[c]);
  std::vector<int64_t> idx(n);
  for (int64_t i = 0; i < n; ++i) {
    int64_t pt[3];
    for (int c = 0; c < 3; ++c) pt[c] = (int64_t)((xyz[3 * i + c] - lo[c]) * spread[c]);  // int(): truncation (:83-86)
    int64_t l = length / 2;
    int g[3] = {pt[0] > l, pt[1] > l, pt[2] > l};
    int v = HRT_VAL[g[0]][g[1]][g[2]];
    int64_t id = v;
    while (l > 1) {  // :88-108
      id *= 8;
      for (int c = 0; c < 3; ++c) pt[c] -= g[c] * l;
      l /= 2;
      for (int c = 0; c < 3; ++c) g[c] = pt[c] > l;
      v = HRT_TRANS[v][HRT_VAL[g[0]][g[1]][g[2]]];
      id += v;
    }
    idx[i] = id;
  }
  // descending index, ties keep their input order (:134-140, :171-181 with one sorting node)
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return idx[a] > idx[b]; });
  return SXC_OK;
}

}  // namespace

struct sxc_grid_points {
  std::vector<double> xyz, w;
};

extern "C" {

const char* sxc_grid_last_error(void) { return g_error.c_str(); }

int sxc_atom_grid(int nuclear_charge, int accuracy, int radial_type, sxc_grid_points** out) {
  if (!out) return gfail(SXC_ERR_INVALID, "sxc_atom_grid: bad arguments");
  AtomGrid ag;
  const int rc = atom_grid(nuclear_charge, accuracy, radial_type, ag);
  if (rc != SXC_OK) return rc;
  auto* g = new sxc_grid_points();
  g->xyz = std::move(ag.xyz);
  g->w = std::move(ag.w);
  *out = g;
  return SXC_OK;
}

int sxc_hilbert_rtree_order(int64_t npts, const double* xyz, int64_t* order) {
  if (npts < 0 || (npts > 0 && (!xyz || !order))) return gfail(SXC_ERR_INVALID, "sxc_hilbert_rtree_order: bad arguments");
  std::vector<double> p(xyz, xyz + 3 * npts);
  std::vector<int64_t> o;
  const int rc = hilbert_rtree_order(p, o);
  if (rc != SXC_OK) return rc;
  std::copy(o.begin(), o.end(), order);
  return SXC_OK;
}

int sxc_molecular_grid(sxc_ctx* ctx, int natoms, const int* nuclear_charges, const double* coords_bohr, int accuracy, int flavour,
                       int radial_type, int becke_smoothing, double weight_threshold, int hilbert_sort, sxc_grid_points** out) {
  if (!ctx || natoms < 1 || !nuclear_charges || !coords_bohr || !out) return gfail(SXC_ERR_INVALID, "sxc_molecular_grid: bad arguments");
  // 1. every atom's reference grid shifted to its nucleus (GridFactory.cpp:117-137)
  std::vector<AtomGrid> cache(ZMAX + 1);
  std::vector<double> xyz, w;
  std::vector<int> parent;
  for (int a = 0; a < natoms; ++a) {
    const int z = nuclear_charges[a];
    if (z < 1 || z > ZMAX) return gfail(SXC_ERR_UNSUPPORTED, "atom grids are tabulated for H..Kr");
    if (cache[z].w.empty()) {
      const int rc = atom_grid(z, accuracy, radial_type, cache[z]);
      if (rc != SXC_OK) return rc;
    }
    const AtomGrid& ag = cache[z];
    for (size_t k = 0; k < ag.w.size(); ++k) {
      for (int c = 0; c < 3; ++c) xyz.push_back(ag.xyz[3 * k + c] + coords_bohr[3 * a + c]);
      w.push_back(ag.w[k]);
      parent.push_back(a);
    }
  }
  // 2. partition weights on the device (GridFactory.cpp:139-266); Becke / Voronoi flavours need the size adjustments (:95-113)
  std::vector<double> aij;
  if (flavour != 1) {
    aij.assign((size_t)natoms * natoms, 0.0);
    for (int i = 0; i < natoms; ++i)
      for (int j = 0; j < natoms; ++j) {
        const double bi = BRAGG_SLATER[nuclear_charges[i]], bj = BRAGG_SLATER[nuclear_charges[j]];
        if (bi == 0.0 || bj == 0.0) return gfail(SXC_ERR_UNSUPPORTED, "no Bragg-Slater radius for an atom of this molecule (BECKE / VORONOI flavour)");
        const double q = std::sqrt(bj / bi), u = (q - 1.0) / (q + 1.0), a = u / (u * u - 1.0);
        aij[(size_t)i * natoms + j] = std::min(0.5, std::max(-0.5, a));
      }
  }
  const int64_t n0 = (int64_t)w.size();
  const int rc = sxc_partition_weights(ctx, flavour, becke_smoothing, natoms, coords_bohr, aij.empty() ? nullptr : aij.data(), n0,
                                       xyz.data(), parent.data(), w.data());
  if (rc != SXC_OK) return gfail(rc, std::string("sxc_partition_weights: ") + sxc_last_error(ctx));
  // 3. weight cut (:264)
  auto* g = new sxc_grid_points();
  for (int64_t i = 0; i < n0; ++i)
    if (w[i] > weight_threshold) {
      g->xyz.insert(g->xyz.end(), xyz.begin() + 3 * i, xyz.begin() + 3 * i + 3);
      g->w.push_back(w[i]);
    }
  // 4. locality sort (:287, HilbertRTreeSorting)
  if (hilbert_sort) {
    std::vector<int64_t> order;
    const int rs = hilbert_rtree_order(g->xyz, order);
    if (rs != SXC_OK) {
      delete g;
      return rs;
    }
    std::vector<double> sx(g->xyz.size()), sw(g->w.size());
    for (size_t k = 0; k < order.size(); ++k) {
      const int64_t i = order[k];
      sx[3 * k] = g->xyz[3 * i];
      sx[3 * k + 1] = g->xyz[3 * i + 1];
      sx[3 * k + 2] = g->xyz[3 * i + 2];
      sw[k] = g->w[i];
    }
    g->xyz.swap(sx);
    g->w.swap(sw);
  }
  *out = g;
  return SXC_OK;
}

int64_t sxc_grid_points_size(const sxc_grid_points* g) { return g ? (int64_t)g->w.size() : 0; }
const double* sxc_grid_points_xyz(const sxc_grid_points* g) { return g ? g->xyz.data() : nullptr; }
const double* sxc_grid_points_weights(const sxc_grid_points* g) { return g ? g->w.data() : nullptr; }
void sxc_grid_points_free(sxc_grid_points* g) { delete g; }

}  // extern "C"
