// group.cpp - sxc_group: N GPUs of one node driven from ONE host process (sxc_create(ctx, ngpu, devices)).
//
// The reference calls FuncPotential::getMatrix() from its single SCF driver thread and parallelises inside with OpenMP
// (MatrixOperatorToGridTransformer.cpp:103, ScalarOperatorToMatrixAdder.cpp:66,101).  The drop-in equivalent for a single-process
// host is a group: one device context (sxc_ctx) and one host worker thread per GPU.  A group call hands the same C-ABI call to
// every worker; each context evaluates its shard of the grid blocks and the library's ncclAllReduce (comm.h) sums [V | E | N]
// over NVLink, each rank's collective issued from its own thread on its own stream - the standard one-thread-per-device NCCL
// pattern.  Only the first context copies the matrix back to the caller.  Host-only C++ on top of the public C ABI.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/serenity_xc_b200.h"

namespace {

struct Worker {
  sxc_ctx* ctx = nullptr;
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int(sxc_ctx*)> job;
  bool has_job = false, done = false, stop = false;
  int rc = 0;

  void loop() {
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv.wait(lk, [&] { return has_job || stop; });
      if (stop) return;
      std::function<int(sxc_ctx*)> j = std::move(job);
      has_job = false;
      lk.unlock();
      const int r = j(ctx);
      lk.lock();
      rc = r;
      done = true;
      cv.notify_all();
    }
  }
  void submit(std::function<int(sxc_ctx*)> j) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(j);
    has_job = true;
    done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};

}  // namespace

struct sxc_group {
  std::vector<Worker*> workers;
  std::string err;
  int first_error = 0;

  // run `fn(rank, ctx)` on every worker thread; returns the first non-zero status (and remembers that rank's message)
  int all(const std::function<int(int, sxc_ctx*)>& fn) {
    for (size_t r = 0; r < workers.size(); ++r) workers[r]->submit([fn, r](sxc_ctx* c) { return fn((int)r, c); });
    int rc = SXC_OK;
    for (size_t r = 0; r < workers.size(); ++r) {
      const int x = workers[r]->wait();
      if (x != SXC_OK && rc == SXC_OK) {
        rc = x;
        err = "rank " + std::to_string(r) + ": " + sxc_last_error(workers[r]->ctx);
      }
    }
    return rc;
  }
};

extern "C" {

int sxc_group_create(sxc_group** out, int ngpu, const int* devices) {
  if (!out || ngpu < 1) return SXC_ERR_INVALID;
  *out = nullptr;
  auto* g = new sxc_group();
  for (int r = 0; r < ngpu; ++r) {
    auto* w = new Worker();
    const int rc = sxc_create(&w->ctx, devices ? devices[r] : r);
    if (rc != SXC_OK) {
      delete w;
      for (Worker* x : g->workers) {
        sxc_destroy(x->ctx);
        delete x;
      }
      delete g;
      return rc;
    }
    g->workers.push_back(w);
  }
  for (Worker* w : g->workers) w->th = std::thread([w] { w->loop(); });
  if (ngpu > 1) {
    unsigned char id[SXC_COMM_ID_BYTES];
    int rc = sxc_comm_unique_id(id);
    if (rc == SXC_OK)  // every rank joins from its own thread (ncclCommInitRank blocks until all have arrived)
      rc = g->all([&](int r, sxc_ctx* c) {
        const int rr = sxc_comm_init_rank(c, r, ngpu, id);
        // every rank ends a build with the same all-reduced matrix: each copies 1 / ngpu of it into the caller's buffer
        return rr != SXC_OK ? rr : sxc_set_output_slice(c, r, ngpu);
      });
    if (rc != SXC_OK) {
      sxc_group_destroy(g);
      return rc;
    }
  }
  *out = g;
  return SXC_OK;
}

void sxc_group_destroy(sxc_group* g) {
  if (!g) return;
  for (Worker* w : g->workers) {
    {
      std::lock_guard<std::mutex> lk(w->m);
      w->stop = true;
      w->cv.notify_all();
    }
    if (w->th.joinable()) w->th.join();
  }
  // communicators first (ncclCommDestroy may synchronise with the peers), then the contexts
  for (Worker* w : g->workers) sxc_comm_destroy(w->ctx);
  for (Worker* w : g->workers) {
    sxc_destroy(w->ctx);
    delete w;
  }
  delete g;
}

int sxc_group_size(const sxc_group* g) { return g ? (int)g->workers.size() : 0; }
sxc_ctx* sxc_group_ctx(sxc_group* g, int rank) {
  return (g && rank >= 0 && rank < (int)g->workers.size()) ? g->workers[rank]->ctx : nullptr;
}
const char* sxc_group_last_error(const sxc_group* g) { return g ? g->err.c_str() : "null group"; }

// Handles are allocated in the same order on every context, so one integer names the object on all of them.
static int same_handle(sxc_group* g, const std::vector<int>& h, int* out) {
  for (int x : h)
    if (x != h[0]) {
      g->err = "handles diverged between the contexts of the group (objects must be created through the group only)";
      return SXC_ERR_INVALID;
    }
  *out = h[0];
  return SXC_OK;
}

int sxc_group_set_grid(sxc_group* g, int64_t npts, const double* xyz, const double* w, int blocksize, int* grid) {
  if (!g || !grid) return SXC_ERR_INVALID;
  std::vector<int> h(g->workers.size(), -1);
  const int rc = g->all([&](int r, sxc_ctx* c) { return sxc_set_grid(c, npts, xyz, w, blocksize, &h[r]); });
  return rc != SXC_OK ? rc : same_handle(g, h, grid);
}

int sxc_group_add_basis(sxc_group* g, int nshell, const int* l, const int* pure, const int* nprim, const int* first_bf,
                        const double* centre, const double* alpha, const double* coeff, const double* normfac,
                        double radial_threshold, int* basis) {
  if (!g || !basis) return SXC_ERR_INVALID;
  std::vector<int> h(g->workers.size(), -1);
  const int rc = g->all([&](int r, sxc_ctx* c) {
    return sxc_add_basis(c, nshell, l, pure, nprim, first_bf, centre, alpha, coeff, normfac, radial_threshold, &h[r]);
  });
  return rc != SXC_OK ? rc : same_handle(g, h, basis);
}

int sxc_group_set_functional(sxc_group* g, int ncomp, const int* basic_id, const double* mix, int* func) {
  if (!g || !func) return SXC_ERR_INVALID;
  std::vector<int> h(g->workers.size(), -1);
  const int rc = g->all([&](int r, sxc_ctx* c) { return sxc_set_functional(c, ncomp, basic_id, mix, &h[r]); });
  return rc != SXC_OK ? rc : same_handle(g, h, func);
}

int sxc_group_release_grid(sxc_group* g, int grid) {
  return g ? g->all([&](int, sxc_ctx* c) { return sxc_release_grid(c, grid); }) : SXC_ERR_INVALID;
}
int sxc_group_release_basis(sxc_group* g, int basis) {
  return g ? g->all([&](int, sxc_ctx* c) { return sxc_release_basis(c, basis); }) : SXC_ERR_INVALID;
}

int sxc_group_build_xc(sxc_group* g, int grid, int basis, int func, int nspin, const double* P, double thr, double* V, double* E,
                       double* nelec) {
  if (!g || !P || !V || !E) return SXC_ERR_INVALID;
  std::vector<double> e(g->workers.size(), 0.0), n(g->workers.size(), 0.0);
  const int rc = g->all([&](int r, sxc_ctx* c) {  // every rank ends with the all-reduced [V | E | N] and copies its part of V back
    return sxc_build_xc(c, grid, basis, func, nspin, P, thr, V, &e[r], &n[r]);
  });
  if (rc != SXC_OK) return rc;
  *E = e[0];
  if (nelec) *nelec = n[0];
  return SXC_OK;
}

int sxc_group_build_nadd_multi(sxc_group* g, int grid, int nfunc, const int* funcs, int nspin, int basis_act, const double* P_act,
                               int nenv, const int* basis_env, const double* const* P_env, int env_frozen, double thr,
                               int sum_matrices, double* V_act, double* E) {
  if (!g || !funcs || !P_act || !V_act || !E || nfunc < 1 || nenv < 0) return SXC_ERR_INVALID;
  const size_t ne = (size_t)nfunc * (2 + nenv);
  std::vector<std::vector<double>> e(g->workers.size(), std::vector<double>(ne, 0.0));
  const int rc = g->all([&](int r, sxc_ctx* c) {
    return sxc_build_nadd_multi(c, grid, nfunc, funcs, nspin, basis_act, P_act, nenv, basis_env, P_env, env_frozen, thr,
                                sum_matrices, V_act, e[r].data());
  });
  if (rc != SXC_OK) return rc;
  std::memcpy(E, e[0].data(), ne * sizeof(double));
  return SXC_OK;
}

int sxc_group_xc_gradient(sxc_group* g, int grid, int basis, int func, int nspin, const double* P, int natoms,
                          const int* atom_of_bf, double* grad) {
  if (!g || !P || !grad || natoms <= 0) return SXC_ERR_INVALID;
  std::vector<std::vector<double>> gr(g->workers.size(), std::vector<double>((size_t)natoms * 3, 0.0));
  const int rc = g->all([&](int r, sxc_ctx* c) {
    return sxc_xc_gradient(c, grid, basis, func, nspin, P, natoms, atom_of_bf, gr[r].data());
  });
  if (rc != SXC_OK) return rc;
  std::memcpy(grad, gr[0].data(), gr[0].size() * sizeof(double));
  return SXC_OK;
}

}  // extern "C"
