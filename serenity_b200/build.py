"""In-tree build of the native pieces (the built .so files travel to the GPU box with the repo snapshot).

  libserenity_xc_b200.so   CUDA kernels + C ABI, sm_100a only  (serenity_b200/csrc/sxc_api.cu)
  inputs/libsxc_inputs.so  host helper of the synthetic-input producer (partition weights)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]


def _newer(target, sources):
    return (not os.path.exists(target)) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def build_cuda(force=False, verbose=False):
    csrc = os.path.join(HERE, "csrc")
    gen = os.path.join(csrc, "harmonics_gen.cuh")
    if not os.path.exists(gen):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_harmonics.py")])
    out = os.path.join(HERE, "libserenity_xc_b200.so")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "serenity_xc_b200.h")]
    if force or _newer(out, srcs):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-ccbin", "/usr/bin/g++", "-o", out,
                                                                              os.path.join(csrc, "sxc_api.cu"),
                                                                              os.path.join(csrc, "basis_provider.cpp"),
                                                                              os.path.join(csrc, "group.cpp"),
                                                                              os.path.join(csrc, "grid_builder.cpp"), "-lpthread", "-ldl"]
        subprocess.check_call(cmd)
    return out


def build_inputs_helper(force=False):
    src = os.path.join(HERE, "inputs", "gridweights.c")
    out = os.path.join(HERE, "inputs", "libsxc_inputs.so")
    if force or _newer(out, [src]):
        subprocess.check_call(["/usr/bin/gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", out, src,
                               "-lm"])
    return out


def build_host_test(force=False):
    """C++ adapter (serenity_b200/host) + its driver, linked against the C-ABI library."""
    src = os.path.join(ROOT, "tests", "cpp", "host_adapter_test.cpp")
    hdr = os.path.join(HERE, "host", "serenity_xc_adapter.h")
    out = os.path.join(ROOT, "tests", "cpp", "host_adapter_test")
    lib = os.path.join(HERE, "libserenity_xc_b200.so")
    if force or _newer(out, [src, hdr, lib]):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++14", "-Wall", "-Wextra", "-o", out, src, "-L" + HERE,
                               "-lserenity_xc_b200", "-Wl,-rpath,$ORIGIN/../../serenity_b200"])
    return out


def build_bridge_test(force=False):
    """serenity_b200/host/B200Bridge.h (the reference-side glue of INTEGRATION.md) compiled against stand-in controller classes."""
    src = os.path.join(ROOT, "tests", "cpp", "b200_bridge_test.cpp")
    hdr = os.path.join(HERE, "host", "B200Bridge.h")
    out = os.path.join(ROOT, "tests", "cpp", "b200_bridge_test")
    lib = os.path.join(HERE, "libserenity_xc_b200.so")
    if force or _newer(out, [src, hdr, lib]):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++14", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), "-o", out,
                               src, "-L" + HERE, "-lserenity_xc_b200", "-Wl,-rpath,$ORIGIN/../../serenity_b200"])
    return out


def build_jet_probe(force=False):
    """Host-compiled probe of the device functional source (functionals.cuh, kernel2.cuh) for the CPU test suite; the
    arithmetic is __host__ __device__, the probe never touches a GPU and is not part of the product library."""
    src = os.path.join(ROOT, "tests", "cpp", "jet_host_probe.cu")
    csrc = os.path.join(HERE, "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("functionals.cuh", "kernel2.cuh", "sxc_common.cuh")]
    out = os.path.join(ROOT, "tests", "cpp", "libjet_host_probe.so")
    if force or _newer(out, deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                               "-shared", "-ccbin", "/usr/bin/g++", "-o", out, src])
    return out


def build_host_copy_test(force=False):
    """csrc/host_copy.h (pageable <-> page-locked staging with worker threads) exercised without a GPU."""
    src = os.path.join(ROOT, "tests", "cpp", "host_copy_test.cpp")
    hdr = os.path.join(HERE, "csrc", "host_copy.h")
    out = os.path.join(ROOT, "tests", "cpp", "host_copy_test")
    if force or _newer(out, [src, hdr]):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", "-o", out, src])
    return out


def build_all(force=False, verbose=False):
    return [build_cuda(force, verbose), build_inputs_helper(force), build_host_test(force), build_bridge_test(force),
            build_jet_probe(force), build_host_copy_test(force)]


if __name__ == "__main__":
    print("\n".join(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)))
