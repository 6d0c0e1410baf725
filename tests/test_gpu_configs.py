"""Round-2 GPU tests: every BASELINE.json config at its full size against the CPU oracle, plus the C-ABI additions
(handle release, environment tags, the one-pass multi-functional NAdd build, the in-library communicator).

Bars (BASELINE.json north_star): |dE_xc| <= 1e-9 Eh, max|dV_xc| <= 1e-8, FP64.
Reference path: potentials/FuncPotential.cpp:74-111 (configs 1, 2, 3, 5), tasks/FDETask.cpp:344-347 ->
potentials/NAddFuncPotential.cpp:192-300 (config 4)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_TOL = 1e-9
V_TOL = 1e-8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from serenity_b200.xc import XCContext
    c = XCContext(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    pyoracle.set_threads(len(os.sched_getaffinity(0)))
    pyoracle.use_openblas(True)   # (1e-16-level summation-order differences only; tests/test_oracle_blas.py)
    return pyoracle


@pytest.fixture(scope="module")
def water64_grid():
    from serenity_b200.inputs import make_config
    cfg = make_config("water64")
    return cfg


def _functional(name):
    from serenity_b200.inputs.configs import FUNCTIONALS
    return FUNCTIONALS[name]


def _ks_parity(ctx, orc, cfg):
    sub = cfg.subsystems[0]
    ids, mix = _functional(cfg.functional)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    try:
        _, _, ne = ctx.build_xc(g, b, f, sub.P)
        P = np.asfortranarray(sub.P * (sub.n_electrons / ne))
        V, E, ne = ctx.build_xc(g, b, f, P)
        st = ctx.stats()
    finally:
        ctx.release_grid(g)
        ctx.release_basis(b)
    V_ref, E_ref, ne_ref, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), P)
    assert np.isfinite(V).all() and np.array_equal(V, V.T)
    assert abs(ne - sub.n_electrons) < 1e-7
    assert abs(E - E_ref) <= E_TOL, (cfg.name, E, E_ref)
    assert np.abs(V - V_ref).max() <= V_TOL, (cfg.name, np.abs(V - V_ref).max())
    assert abs(ne - ne_ref) <= 1e-10 * abs(ne_ref)
    return st


def test_config3_water64_full_size(ctx, orc, water64_grid):
    """BASELINE configs[2]: (H2O)64 PBE/def2-SVP, 1536 basis functions, 7.7e5 grid points."""
    st = _ks_parity(ctx, orc, water64_grid)
    assert st["nbf"] == 1536 and st["s_max"] < 1536  # screening prunes: no block sees the whole basis


def test_config5_peptide_full_size(ctx, orc):
    """BASELINE configs[4]: 216-atom peptide stand-in PBE/def2-SVP on the accuracy-6 grid (2.9e6 points, 25 GB of tiles)."""
    from serenity_b200.inputs import make_config
    st = _ks_parity(ctx, orc, make_config("peptide"))
    assert st["nchunks"] == 1


def test_config4_fde_water64_full_size_one_pass(ctx, orc, water64_grid):
    """BASELINE configs[3]: freeze-and-thaw FDE of two (H2O)32 subsystems, NAdd XC (PBE) + NAdd kinetic (PW91k) of the active
    one on the supersystem grid.  One device pass (sxc_build_nadd_multi) must return, per object, the matrix and energies of
    the oracle's NAddFuncPotential restatement, and the summed-matrix mode their sum."""
    from serenity_b200.inputs import make_config
    cfg = make_config("fde_water64", grid=(water64_grid.xyz, water64_grid.w))
    act, env = cfg.subsystems
    names = [cfg.functional, cfg.nadd_kin]
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA, bE = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    fh = [ctx.set_functional(*_functional(n)) for n in names]
    try:
        Vs, Es = ctx.build_nadd_multi(g, fh, bA, act.P, [bE], [env.P], env_frozen=3, sum_matrices=False)
        Vsum, Esum = ctx.build_nadd_multi(g, fh, bA, act.P, [bE], [env.P], env_frozen=3, sum_matrices=True)
        V1, E1 = ctx.build_nadd(g, fh[1], bA, act.P, [bE], [env.P], env_frozen=3)  # the per-object call, served by the cache
    finally:
        ctx.release_grid(g)
        ctx.release_basis(bA)
        ctx.release_basis(bE)
    og, oa, oe = orc.Grid(cfg.xyz, cfg.w, 128), orc.Basis(act.basis), orc.Basis(env.basis)
    V_tot = 0.0
    for k, n in enumerate(names):
        V_ref, E_ref, parts = orc.build_nadd(oa, act.P, [(oe, env.P)], og, orc.Functional(*_functional(n)))
        assert np.abs(Vs[k] - V_ref).max() <= V_TOL, (n, np.abs(Vs[k] - V_ref).max())
        assert np.allclose(Es[k], parts, rtol=0, atol=E_TOL), (n, Es[k], parts)
        assert abs((Es[k][0] - Es[k][1] - Es[k][2:].sum()) - E_ref) <= E_TOL
        V_tot = V_tot + V_ref
    assert np.abs(Vsum - V_tot).max() <= V_TOL
    assert np.abs(Esum - Es).max() <= 1e-12
    assert np.abs(V1 - Vs[1]).max() <= 1e-12 and np.abs(E1 - Es[1]).max() <= 1e-12


def test_fused_scatter_variant_matches_oracle(orc, monkeypatch):
    """SXC_VMAT=24 (opt-in): k_vmat_fg forms G inside the persistent scatter kernel (helper warpgroups, per-block flags).  Whole-block
    work items need a shard of >= 3 waves of CTAs (tetracene at accuracy 4: 1083 blocks); a small grid takes its fallback."""
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.xc import XCContext
    monkeypatch.setenv("SXC_VMAT", "24")
    c = XCContext(0)
    try:
        for name, acc, fn in (("tetracene", 4, "B3LYP"), ("water8", 3, "PBE"), ("h2o", 2, "LDA")):
            cfg = make_config(name, acc)
            sub = cfg.subsystems[0]
            ids, mix = FUNCTIONALS[fn]
            g = c.set_grid(cfg.xyz, cfg.w, 128)
            b = c.add_basis(sub.basis, 1e-9)
            f = c.set_functional(ids, mix)
            for _ in range(2):  # (the second build reuses the plan and the per-block flags)
                V, E, ne = c.build_xc(g, b, f, sub.P)
            Vr, Er, ner, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), sub.P)
            assert abs(E - Er) <= 1e-9 and np.abs(V - Vr).max() <= 1e-8 and abs(ne - ner) <= 1e-10 * abs(ner), name
            assert np.array_equal(V, V.T)
    finally:
        c.close()


def test_functional_kernel_occupancy_variants_agree(monkeypatch):
    """SXC_FUNC = 0 / 4 (default) / 5: k_functional compiled for 3 / 4 / 5 resident CTAs per SM (168 / 128 / 96 registers, the
    smaller ones with a few spilled doubles).  Register allocation must not change E_xc, V_xc or the electron count beyond
    rounding (B3LYP = four basic functionals; the scatter's red.global order is run-to-run noise at the 1e-15 level)."""
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.xc import XCContext
    cfg = make_config("water8", 3)
    sub = cfg.subsystems[0]
    out = {}
    for variant in ("0", "4", "5"):
        monkeypatch.setenv("SXC_FUNC", variant)
        c = XCContext(0)
        try:
            g = c.set_grid(cfg.xyz, cfg.w, 128)
            b = c.add_basis(sub.basis, 1e-9)
            for fn in ("B3LYP", "PBE"):
                f = c.set_functional(*FUNCTIONALS[fn])
                out[variant, fn] = c.build_xc(g, b, f, sub.P)
        finally:
            c.close()
    for fn in ("B3LYP", "PBE"):
        V0, E0, n0 = out["0", fn]
        for variant in ("4", "5"):
            V, E, n = out[variant, fn]
            assert abs(E - E0) <= 1e-11 and n == n0, (variant, fn, E - E0)
            assert np.abs(V - V0).max() <= 1e-12, (variant, fn)


def test_config1_h2o_accuracy4(ctx, orc):
    """BASELINE configs[0]: H2O PBE/def2-SVP on the accuracy-4 grid (the reference's CPU-runnable case)."""
    from serenity_b200.inputs import make_config
    _ks_parity(ctx, orc, make_config("h2o", 4))


# ------------------------------------------------------------------------------------------- C-ABI additions
def test_release_recycles_handles_and_memory(ctx):
    """sxc_release_grid / sxc_release_basis: a geometry step uploads a new grid and basis; handles are reused, results do not
    depend on what lived in the slot before, released handles are refused."""
    import torch
    from serenity_b200._lib import SerenityError
    from serenity_b200.inputs import make_config
    cfg = make_config("water8")
    sub = cfg.subsystems[0]
    f = ctx.set_functional(*_functional("PBE"))
    assert ctx.set_functional(*_functional("PBE")) == f      # equal definitions share one handle
    g0 = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b0 = ctx.add_basis(sub.basis, 1e-9)
    V0, E0, n0 = ctx.build_xc(g0, b0, f, sub.P)
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):                                       # would leak ~ 6 x (grid + work arrays + plan) without release
        ctx.release_grid(g0)
        ctx.release_basis(b0)
        g1 = ctx.set_grid(cfg.xyz, cfg.w, 128)
        b1 = ctx.add_basis(sub.basis, 1e-9)
        assert (g1, b1) == (g0, b0)
        V1, E1, n1 = ctx.build_xc(g1, b1, f, sub.P)
        assert E1 == E0 and np.abs(V1 - V0).max() < 1e-12
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20
    ctx.release_grid(g0)
    with pytest.raises(SerenityError):
        ctx.build_xc(g0, b0, f, sub.P)
    with pytest.raises(SerenityError):
        ctx.release_grid(g0)
    ctx.release_basis(b0)


def test_environment_tags_keep_two_nadd_objects_apart(ctx, orc):
    """the frozen-environment cache lives on the grid.  Two NAdd objects whose environments share a basis handle but
    hold different densities must not see each other's cache: distinct tags keep them apart, and a re-used tag with a changed
    density is the caller's error this test documents (it returns the cached environment)."""
    from serenity_b200.inputs import make_config
    cfg = make_config("fde_dimer")
    act, env = cfg.subsystems
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA, bE = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    f = ctx.set_functional(*_functional("PBE"))
    P1, P2 = env.P, np.asfortranarray(env.P * 0.8)
    ref1 = ctx.build_nadd(g, f, bA, act.P, [bE], [P1], env_frozen=0)
    ref2 = ctx.build_nadd(g, f, bA, act.P, [bE], [P2], env_frozen=0)
    assert np.abs(ref1[0] - ref2[0]).max() > 1e-6
    for _ in range(2):  # alternating objects, each with its own tag: always its own environment
        V, E = ctx.build_nadd(g, f, bA, act.P, [bE], [P1], env_frozen=11)
        assert np.abs(V - ref1[0]).max() < 1e-12 and np.abs(E - ref1[1]).max() < 1e-12
        V, E = ctx.build_nadd(g, f, bA, act.P, [bE], [P2], env_frozen=12)
        assert np.abs(V - ref2[0]).max() < 1e-12 and np.abs(E - ref2[1]).max() < 1e-12
    V, E = ctx.build_nadd(g, f, bA, act.P, [bE], [P1], env_frozen=12)   # same tag, other density: the cache answers
    assert np.abs(V - ref2[0]).max() < 1e-12
    ctx.release_grid(g)
    ctx.release_basis(bA)
    ctx.release_basis(bE)


def test_nadd_multi_unrestricted_and_lda_mix(ctx, orc):
    """sxc_build_nadd_multi with an LDA kinetic functional next to a GGA XC functional (TF + PBE, the reference's FDE test
    combination FDETask_test.cpp:58) restricted and unrestricted: one pass == the single-functional builds."""
    from serenity_b200.inputs import make_config
    cfg = make_config("fde_dimer")
    act, env = cfg.subsystems
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA, bE = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    fh = [ctx.set_functional(*_functional(n)) for n in ("TF", "PBE")]
    for nspin in (1, 2):
        Pa = act.P if nspin == 1 else (act.P * 0.55, act.P * 0.45)
        Pe = env.P if nspin == 1 else (env.P * 0.5, env.P * 0.5)
        single = [ctx.build_nadd(g, f, bA, Pa, [bE], [Pe], nspin=nspin) for f in fh]
        Vs, Es = ctx.build_nadd_multi(g, fh, bA, Pa, [bE], [Pe], sum_matrices=False, nspin=nspin)
        Vsum, Esum = ctx.build_nadd_multi(g, fh, bA, Pa, [bE], [Pe], sum_matrices=True, nspin=nspin)
        for k in range(2):
            for s in range(nspin):
                a = Vs[k][s] if nspin == 2 else Vs[k]
                r = single[k][0][s] if nspin == 2 else single[k][0]
                assert np.abs(a - r).max() < 1e-12
            assert np.abs(Es[k] - single[k][1]).max() < 1e-12
        for s in range(nspin):
            a = Vsum[s] if nspin == 2 else Vsum
            r = (single[0][0][s] + single[1][0][s]) if nspin == 2 else (single[0][0] + single[1][0])
            assert np.abs(a - r).max() < 1e-11
        assert np.abs(Esum - Es).max() < 1e-12
    ctx.release_grid(g)
    ctx.release_basis(bA)
    ctx.release_basis(bE)


def test_single_rank_communicator_is_transparent():
    """sxc_comm_init_rank with world = 1: NCCL is bound (dlopen), grids become shard 0 of 1, results are those of a plain context."""
    from serenity_b200.inputs import make_config
    from serenity_b200.xc import XCContext
    cfg = make_config("h2o", 2)
    sub = cfg.subsystems[0]
    ids, mix = _functional("PBE")
    out = []
    for with_comm in (False, True):
        c = XCContext(0)
        if with_comm:
            c.comm_init_rank(0, 1, XCContext.comm_unique_id())
            info = c.comm_info()
            assert info["world"] == 1 and info["rank"] == 0 and info["nccl_version"] >= 20000
        g = c.set_grid(cfg.xyz, cfg.w, 128)
        out.append(c.build_xc(g, c.add_basis(sub.basis, 1e-9), c.set_functional(ids, mix), sub.P))
        c.close()
    assert out[0][1] == out[1][1] and np.abs(out[0][0] - out[1][0]).max() < 1e-13


_TWO_RANK = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from serenity_b200.inputs import make_config
from serenity_b200.inputs.configs import FUNCTIONALS
from serenity_b200.xc import XCContext
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")                      # bootstrap only: ships the 128-byte NCCL id
box = [XCContext.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx = XCContext(rank)
ctx.comm_init_rank(rank, world, box[0])
cfg = make_config("water8")
sub = cfg.subsystems[0]
g = ctx.set_grid(cfg.xyz, cfg.w, 128)                # becomes this rank's shard
b = ctx.add_basis(sub.basis, 1e-9)
f = ctx.set_functional(*FUNCTIONALS["PBE"])
V, E, ne = ctx.build_xc(g, b, f, sub.P)              # sum over ranks on every rank (ncclAllReduce inside the library)
npts = ctx.stats()["npts"]
act, env = make_config("fde_water8", grid=(cfg.xyz, cfg.w)).subsystems
bA, bE = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
fk = ctx.set_functional(*FUNCTIONALS["PW91K"])
Vn, En = ctx.build_nadd_multi(g, [f, fk], bA, act.P, [bE], [env.P], env_frozen=1, sum_matrices=True)
info = ctx.comm_info()
np.savez(os.path.join(%(out)r, "rank%%d.npz" %% rank), V=V, E=E, ne=ne, npts=npts, Vn=Vn, En=En, coll=info["collectives"])
ctx.close()
dist.destroy_process_group()
"""


def test_two_rank_nccl_allreduce_inside_the_library(tmp_path, orc):
    """Two processes, two GPUs: sxc_comm_init_rank + sxc_build_xc / sxc_build_nadd_multi return the all-reduced result on both
    ranks (no torch.distributed collective on the data path).  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from serenity_b200.inputs import make_config
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK % {"root": ROOT, "out": str(tmp_path)})
    env = dict(os.environ, OMP_NUM_THREADS="4")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                    "127.0.0.1", "--master-port", "29533", str(script)], check=True, env=env, timeout=600)
    r = [np.load(tmp_path / ("rank%d.npz" % k)) for k in range(2)]
    cfg = make_config("water8")
    sub = cfg.subsystems[0]
    ids, mix = _functional("PBE")
    V_ref, E_ref, ne_ref, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), sub.P)
    assert int(r[0]["npts"]) + int(r[1]["npts"]) == cfg.npts and min(int(r[0]["npts"]), int(r[1]["npts"])) > 0
    for k in range(2):
        assert abs(float(r[k]["E"]) - E_ref) <= E_TOL and np.abs(r[k]["V"] - V_ref).max() <= V_TOL
        assert int(r[k]["coll"]) == 2
    assert np.array_equal(r[0]["V"], r[1]["V"]) and np.array_equal(r[0]["Vn"], r[1]["Vn"])
    act, env_s = make_config("fde_water8", grid=(cfg.xyz, cfg.w)).subsystems
    og, oa, oe = orc.Grid(cfg.xyz, cfg.w, 128), orc.Basis(act.basis), orc.Basis(env_s.basis)
    V_tot = sum(orc.build_nadd(oa, act.P, [(oe, env_s.P)], og, orc.Functional(*_functional(n)))[0] for n in ("PBE", "PW91K"))
    assert np.abs(r[0]["Vn"] - V_tot).max() <= V_TOL
