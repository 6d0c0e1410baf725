// basis_provider.cpp - basis-set file -> shell table, host-only part of libserenity_xc_b200.so.
//
// What BasisFunctionOnGridController needs from Serenity's basis front end, so that the adapter can be fed directly from a
// geometry and a basis-set file of the reference's data directory (Turbomole format, /root/reference/data/basis/*):
//   BasisFunctionProvider::provideAtomWithBasisFunctions   src/basis/BasisFunctionProvider.cpp:32-140   (file parsing)
//   BasisFunctionProvider::resolveAngularMomentumChar      :199-247
//   Shell::Shell                                           src/basis/Shell.cpp:29-47   (libint2::Shell + Cartesian norm factors)
//   libint2::Shell::renorm (libint 2.7.0-beta.6, not vendored; algorithm restated)
//   BasisController extended indices                       src/basis/BasisController.cpp:62-68
//   AtomCenteredBasisController: atom-major, file order    src/basis/AtomCenteredBasisController.cpp
// ECP sections of the files ($ecp) belong to the integral code and are not read.  Errors are reported the way the reference
// words them (SerenityError text -> sxc_host_last_error()).
// Deliberately stricter than the reference's search (std::regex_search of "<element>\\s+<label>", icase, unanchored, :61-80), and
// identical on the Turbomole files it ships: the element symbol must start a line and the label must end at white space, so a
// label that is a prefix of another label (def2-SVP / def2-SVPD) or a symbol that ends another word cannot select the wrong
// entry; every '*' / '#' line after the header is skipped (the reference skips one optional '#' line and then 3 characters,
// :84-91); lower-case Fortran exponents (d+01) are converted as well as D+01 (:93).  tests/test_basis_provider.py pins each.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/serenity_xc_b200.h"

struct sxc_shell_table {
  std::vector<int> l, pure, nprim, first_bf, atom_of_bf;
  std::vector<double> centre, alpha, coeff, normfac;
  int nbf = 0;
};

namespace {

thread_local std::string g_host_error;

int host_fail(int code, const std::string& msg) {
  g_host_error = msg;
  return code;
}

double dfact(int n) {  // n!! with (-1)!! = 1
  double r = 1.0;
  for (; n > 1; n -= 2) r *= n;
  return r;
}

// libint2::Shell::renorm: coefficients times the primitive normalisation, then the contraction scaled to unit norm
void renorm(int l, const std::vector<double>& a, std::vector<double>& c) {
  const double pi32 = std::pow(M_PI, 1.5), df = dfact(2 * l - 1);
  for (size_t p = 0; p < a.size(); ++p) c[p] *= std::sqrt(std::pow(2.0, l) * std::pow(2.0 * a[p], l + 1.5) / (pi32 * df));
  double norm = 0.0;
  for (size_t p = 0; p < a.size(); ++p)
    for (size_t q = 0; q < a.size(); ++q) norm += c[p] * c[q] * df * pi32 / (std::pow(2.0, l) * std::pow(a[p] + a[q], l + 1.5));
  const double s = 1.0 / std::sqrt(norm);
  for (double& x : c) x *= s;
}

int resolve_angular_momentum(char type) {  // BasisFunctionProvider.cpp:199-247
  static const char* order = "spdfghikmno";
  const char* p = std::strchr(order, type);
  return (p && type) ? (int)(p - order) : -1;
}

std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
  return s;
}

struct RawShell {
  int l;
  std::vector<double> exps, coefs;
};

// the entry "<element>  <LABEL>" ... "*" of one element (BasisFunctionProvider.cpp:56-96)
int parse_element(const std::string& file_lower, const std::string& file, const std::string& path, const std::string& element,
                  const std::string& label, std::vector<RawShell>& out) {
  const std::string el = lower(element), lab = lower(label);
  const std::string err = "Error while parsing basis set file " + path + " for element " + el + ". (Basis type: " + label + ")\n";
  // element symbol at the start of a line, white space, basis label (case-insensitive)
  size_t pos = 0, start = std::string::npos;
  while ((pos = file_lower.find(el, pos)) != std::string::npos) {
    const bool line_start = pos == 0 || file_lower[pos - 1] == '\n';
    size_t q = pos + el.size();
    if (line_start && q < file_lower.size() && std::isspace((unsigned char)file_lower[q])) {
      while (q < file_lower.size() && (file_lower[q] == ' ' || file_lower[q] == '\t')) ++q;
      if (file_lower.compare(q, lab.size(), lab) == 0 &&
          (q + lab.size() == file_lower.size() || std::isspace((unsigned char)file_lower[q + lab.size()]))) {
        start = q + lab.size();
        break;
      }
    }
    ++pos;
  }
  if (start == std::string::npos) return host_fail(SXC_ERR_INVALID, err + "The used basis (file) is not defined for this element.");
  // to the end of the header line, over the "*" separator line and over an optional "# o (7s4p1d) / [3s2p1d]" comment
  size_t cur = file.find('\n', start);
  if (cur == std::string::npos) return host_fail(SXC_ERR_INVALID, err + "Unexpected end of file.");
  ++cur;
  for (;;) {
    size_t eol = file.find('\n', cur);
    if (eol == std::string::npos) eol = file.size();
    size_t first = cur;
    while (first < eol && std::isspace((unsigned char)file[first])) ++first;
    if (first < eol && (file[first] == '*' || file[first] == '#')) {
      cur = std::min(eol + 1, file.size());
      continue;
    }
    break;
  }
  const size_t end = file.find('*', cur);  // the data of one element is terminated by a star
  std::string block = file.substr(cur, end == std::string::npos ? std::string::npos : end - cur);
  for (size_t i = 0; i + 1 < block.size(); ++i)  // Fortran exponents: D+ / D- -> E+ / E-  (:93)
    if ((block[i] == 'D' || block[i] == 'd') && (block[i + 1] == '+' || block[i + 1] == '-')) block[i] = 'E';
  std::stringstream work(block);
  int nprim = 0;
  while (work >> nprim) {
    char type = 0;
    if (!(work >> type)) return host_fail(SXC_ERR_INVALID, err + "Number of primitives for a contraction could not be parsed.");
    RawShell sh;
    sh.l = resolve_angular_momentum((char)std::tolower((unsigned char)type));
    if (sh.l < 0)
      return host_fail(SXC_ERR_INVALID, err + "The BasisFunctionProvider tries to read in a basis with an unknown symbol for the "
                                              "angular momentum.");
    for (int i = 0; i < nprim; ++i) {
      double e = 0.0, c = 0.0;
      if (!(work >> e) || !(work >> c))
        return host_fail(SXC_ERR_INVALID, err + "A possible source of error is a mismatch between the number of primitives and the "
                                                "number of exponent-contraction entries.");
      sh.exps.push_back(e);
      sh.coefs.push_back(c);
    }
    out.push_back(std::move(sh));
  }
  if (out.empty()) return host_fail(SXC_ERR_INVALID, err + "No contraction found.");
  return SXC_OK;
}

}  // namespace

extern "C" {

const char* sxc_host_last_error(void) { return g_host_error.c_str(); }

int sxc_shell_table_from_file(const char* path, const char* basis_label, int natoms, const char* const* elements,
                              const double* coords_bohr, int spherical, sxc_shell_table** out) {
  if (!path || !basis_label || natoms <= 0 || !elements || !coords_bohr || !out)
    return host_fail(SXC_ERR_INVALID, "sxc_shell_table_from_file: bad arguments");
  std::ifstream f(path);
  if (!f.good())
    return host_fail(SXC_ERR_INVALID, std::string("Error while parsing basis file ") + path +
                                          "\n Make sure the directory exists, change the path in the input, or set $SERENITY_RESOURCES.");
  std::stringstream buf;
  buf << f.rdbuf();
  const std::string file = buf.str(), file_lower = lower(file);
  std::map<std::string, std::vector<RawShell>> cache;  // one parse per element
  auto tab = new sxc_shell_table();
  for (int a = 0; a < natoms; ++a) {
    const std::string el = lower(elements[a] ? elements[a] : "");
    auto it = cache.find(el);
    if (it == cache.end()) {
      std::vector<RawShell> shells;
      const int rc = parse_element(file_lower, file, path, el, basis_label, shells);
      if (rc != SXC_OK) {
        delete tab;
        return rc;
      }
      it = cache.emplace(el, std::move(shells)).first;
    }
    for (const RawShell& sh : it->second) {
      if (sh.l > 6) {  // AM_MAX of the grid path (src/parameters/Constants.h:31)
        delete tab;
        return host_fail(SXC_ERR_UNSUPPORTED, "angular momentum above 6 is not supported on the grid");
      }
      std::vector<double> c = sh.coefs;
      renorm(sh.l, sh.exps, c);
      tab->l.push_back(sh.l);
      tab->pure.push_back(spherical ? 1 : 0);
      tab->nprim.push_back((int)sh.exps.size());
      tab->first_bf.push_back(tab->nbf);
      for (int k = 0; k < 3; ++k) tab->centre.push_back(coords_bohr[3 * a + k]);
      tab->alpha.insert(tab->alpha.end(), sh.exps.begin(), sh.exps.end());
      tab->coeff.insert(tab->coeff.end(), c.begin(), c.end());
      int nf = 0;
      if (spherical) {
        nf = 2 * sh.l + 1;
        tab->normfac.insert(tab->normfac.end(), nf, 1.0);
      } else {  // Shell.cpp:37-47: sqrt((2l-1)!! / ((2ax-1)!!(2ay-1)!!(2az-1)!!)), ax = l..0, ay = l-ax..0
        for (int ax = sh.l; ax >= 0; --ax)
          for (int ay = sh.l - ax; ay >= 0; --ay, ++nf) {
            const int az = sh.l - ax - ay;
            tab->normfac.push_back(std::sqrt(dfact(2 * sh.l - 1) / (dfact(2 * ax - 1) * dfact(2 * ay - 1) * dfact(2 * az - 1))));
          }
      }
      tab->atom_of_bf.insert(tab->atom_of_bf.end(), nf, a);
      tab->nbf += nf;
    }
  }
  *out = tab;
  return SXC_OK;
}

int sxc_shell_table_sizes(const sxc_shell_table* t, int* nshell, int* nprim_total, int* nbf) {
  if (!t) return host_fail(SXC_ERR_INVALID, "null shell table");
  if (nshell) *nshell = (int)t->l.size();
  if (nprim_total) *nprim_total = (int)t->alpha.size();
  if (nbf) *nbf = t->nbf;
  return SXC_OK;
}

int sxc_shell_table_copy(const sxc_shell_table* t, int* l, int* pure, int* nprim, int* first_bf, double* centre, double* alpha,
                         double* coeff, double* normfac, int* atom_of_bf) {
  if (!t) return host_fail(SXC_ERR_INVALID, "null shell table");
  auto cp = [](auto* dst, const auto& v) {
    if (dst) std::copy(v.begin(), v.end(), dst);
  };
  cp(l, t->l);
  cp(pure, t->pure);
  cp(nprim, t->nprim);
  cp(first_bf, t->first_bf);
  cp(centre, t->centre);
  cp(alpha, t->alpha);
  cp(coeff, t->coeff);
  cp(normfac, t->normfac);
  cp(atom_of_bf, t->atom_of_bf);
  return SXC_OK;
}

void sxc_shell_table_free(sxc_shell_table* t) { delete t; }

int sxc_add_basis_from_table(sxc_ctx* ctx, const sxc_shell_table* t, double radial_threshold, int* basis) {
  if (!t) return host_fail(SXC_ERR_INVALID, "null shell table");
  return sxc_add_basis(ctx, (int)t->l.size(), t->l.data(), t->pure.data(), t->nprim.data(), t->first_bf.data(), t->centre.data(),
                       t->alpha.data(), t->coeff.data(), t->normfac.data(), radial_threshold, basis);
}

}  // extern "C"
