"""The CPU oracle's optional OpenBLAS back end (bench.py's cpu_baseline / --impl reference legs) and its AVX-512 build give
the results of the built-in loops up to summation order - the timed CPU arm computes the same thing as the pinned oracle."""
import numpy as np
import pytest


def _build(orc, cfg, name):
    from serenity_b200.inputs.configs import FUNCTIONALS
    sub = cfg.subsystems[0]
    ids, mix = FUNCTIONALS[name]
    V, E, ne, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), sub.P)
    return V, E, ne


def test_openblas_backend_matches_builtin_loops():
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    cfg = make_config("h2o", 2)
    info = orc.use_openblas(True)
    if "OpenBLAS" not in info["gemm"]:
        orc.use_openblas(False)
        pytest.skip("no loadable OpenBLAS with scipy_cblas_dgemm in this environment")
    try:
        assert orc.lib().orc_has_dgemm() == 1
        Vb, Eb, nb = _build(orc, cfg, "PBE")
    finally:
        orc.use_openblas(False)
    assert orc.lib().orc_has_dgemm() == 0
    V, E, n = _build(orc, cfg, "PBE")
    assert abs(E - Eb) < 1e-12 and abs(n - nb) < 1e-12 and np.abs(V - Vb).max() < 1e-13


def test_both_vector_variants_are_built_and_agree():
    import ctypes as C
    import os
    from oracle import pyoracle as orc
    here = os.path.dirname(orc.__file__)
    orc.build()
    assert os.path.exists(os.path.join(here, "liboracle.so")) and os.path.exists(os.path.join(here, "liboracle_avx512.so"))
    assert orc.variant() in ("avx2", "avx512")
    for name in ("liboracle.so",) + (("liboracle_avx512.so",) if orc._cpu_has_avx512() else ()):
        lib = C.CDLL(os.path.join(here, name))
        for sym in ("orc_build_xc", "orc_set_dgemm", "orc_probe_get", "orc_build_nadd"):
            assert hasattr(lib, sym), (name, sym)
