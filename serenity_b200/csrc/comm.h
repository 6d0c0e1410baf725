// comm.h - the single collective of the multi-GPU build: NCCL all-reduce of [V | E | N] over NVLink 5 / NVSwitch, inside the library.
//
// grid blocks are sharded over the GPUs of one node; the only coupling is the additive reduction of
// V_xc, E_xc and the electron count - what ScalarOperatorToMatrixAdder.cpp:73-75 / :108-110 does serially over the OpenMP
// threads' accumulators.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process if the host
// application loaded one, else the system library), so a single-GPU host never needs it; every entry point that needs it fails
// with a clear message when it is missing.  Only the handful of calls below are used.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>

#include <string>

namespace sxc {

// nccl.h essentials (stable since NCCL 2.0): opaque communicator, 128-byte unique id, result / type / op codes
typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id {
  char internal[128];
};
constexpr int NCCL_SUCCESS = 0;
constexpr int NCCL_DOUBLE = 8;  // ncclFloat64
constexpr int NCCL_SUM = 0;

struct NcclApi {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;

  bool load() {
    if (handle) return true;
    // 1. an explicit choice (SXC_NCCL_LIBRARY: a host that will load its own NCCL later, e.g. PyTorch's bundled copy, names that
    //    file so that both end up with the same library), 2. the copy already in the process, 3. the system library
    if (const char* path = std::getenv("SXC_NCCL_LIBRARY"))
      if (*path) handle = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (handle) break;
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!handle) {
      error = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror();
      return false;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(handle, name);
      if (!p && error.empty()) error = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    if (!error.empty()) {
      dlclose(handle);
      handle = nullptr;
      return false;
    }
    return true;
  }
};

inline NcclApi& nccl() {
  static NcclApi api;
  return api;
}

}  // namespace sxc
