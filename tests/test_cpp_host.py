"""The C++ host adapter (serenity_b200/host/serenity_xc_adapter.h) driven like the reference's potential tests
(potentials/FuncPotential_test.cpp, NAddFuncPotential_test.cpp), checked against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_adapter_test")


def _exe():
    if not os.path.exists(EXE):
        from serenity_b200 import build
        build.build_host_test()
    return EXE


def test_adapter_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_exe(), "--expect-no-device"], capture_output=True, text=True)
    assert r.returncode == 0 and "SerenityError" in r.stdout and "no CPU fallback" in r.stdout


def test_host_copier_streams_every_byte_once():
    """csrc/host_copy.h: the streamed download job of staged_d2h (open / publish / help / close) for 0...7 workers, ragged sizes."""
    from serenity_b200 import build
    r = subprocess.run([build.build_host_copy_test()], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host_copy_test: ok" in r.stdout, r.stdout


def test_resolve_functional_matches_the_python_table():
    """CompositeFunctionals::resolveFunctional of the adapter (host only) against serenity_b200.inputs.configs.FUNCTIONALS, both
    restating dft/functionals/functional_definitions.dat."""
    from serenity_b200.inputs.configs import FUNCTIONALS
    for name, (ids, mix) in FUNCTIONALS.items():
        r = subprocess.run([_exe(), "--resolve", name.lower()], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout
        parts = r.stdout.split()
        got = [(int(p.split(":")[0]), float(p.split(":")[1])) for p in parts[2:]]
        assert got == list(zip(ids, mix)), (name, got)
    assert "hfx=0.2 " in subprocess.run([_exe(), "--resolve", "B3LYP"], capture_output=True, text=True).stdout
    r = subprocess.run([_exe(), "--resolve", "PW91"], capture_output=True, text=True)
    assert r.returncode == 3 and "SerenityError" in r.stdout


def _w(f, a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
    f.write(struct.pack("<q", a.size))
    f.write(a.tobytes())


def _basis(f, t, coords):
    from serenity_b200.inputs.basis import atom_indices_of_basis
    for a, dt in ((t.l, np.int32), (t.pure, np.int32), (t.nprim, np.int32), (t.first_bf, np.int32), (t.centre, np.float64),
                  (t.alpha, np.float64), (t.coeff, np.float64), (t.normfac, np.float64),
                  (atom_indices_of_basis(t, coords), np.int32)):
        _w(f, a, dt)


def _r(f, shape=None):
    n = struct.unpack("<q", f.read(8))[0]
    a = np.frombuffer(f.read(8 * n), dtype=np.float64)
    return a.reshape(shape, order="F") if shape else float(a[0])


@pytest.mark.gpu
def test_cpp_potentials_match_oracle(tmp_path):
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    cfg = make_config("fde_dimer", 2)
    act, env = cfg.subsystems
    xc, kin = FUNCTIONALS["PBE"], FUNCTIONALS["PW91K"]
    PA2 = np.asfortranarray(act.P * 1.05)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        _w(f, cfg.xyz, np.float64)
        _w(f, cfg.w, np.float64)
        for ids, mix in (xc, kin):
            _w(f, ids, np.int32)
            _w(f, mix, np.float64)
        _basis(f, act.basis, act.coords)
        _basis(f, env.basis, env.coords)
        _w(f, act.P.reshape(-1, order="F"), np.float64)
        _w(f, env.P.reshape(-1, order="F"), np.float64)
        _w(f, PA2.reshape(-1, order="F"), np.float64)
        rng = np.random.default_rng(9)
        DA = rng.standard_normal((act.basis.nbf, act.basis.nbf))
        DB = rng.standard_normal((env.basis.nbf, env.basis.nbf))
        _w(f, DA.reshape(-1, order="F"), np.float64)
        _w(f, DB.reshape(-1, order="F"), np.float64)
        BtoA = 0.3 * rng.standard_normal((env.basis.nbf, act.basis.nbf)) / np.sqrt(env.basis.nbf)
        _w(f, BtoA.reshape(-1, order="F"), np.float64)
    import torch
    ngpu = min(2, torch.cuda.device_count())
    r = subprocess.run([_exe(), fin, fout, str(ngpu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nA = act.basis.nbf
    with open(fout, "rb") as f:
        got = [(_r(f, (nA, nA)), _r(f)) for _ in range(5)]
        grad = _r(f, (len(act.symbols), 3))
        Vab = _r(f, (nA, env.basis.nbf))
        Vabn = _r(f, (nA, env.basis.nbf))
        gradN = _r(f, (len(act.symbols), 3))
        nB = env.basis.nbf
        F_fde, F_fde_B, F_iso0, F_iso1 = _r(f, (nA, nA)), _r(f, (nB, nB)), _r(f, (nA, nA)), _r(f, (nA, nA))
        pp_block = _r(f, (128,))
        V_stage, E_stage = _r(f, (nA, nA)), _r(f)
        lin = _r(f)
        V_comb, E_comb = _r(f, (nA, nA)), _r(f)
        F_sum, F_x, F_k, E_x, E_k = _r(f, (nA, nA)), _r(f, (nA, nA)), _r(f, (nA, nA)), _r(f), _r(f)
        pp_mixed = [_r(f, (128,)) for _ in range(3)]
        if ngpu > 1:
            Vg, Eg, Fg, gradg, ng = _r(f, (nA, nA)), _r(f), _r(f, (nA, nA)), _r(f, (len(act.symbols), 3)), _r(f)
    og = orc.Grid(cfg.xyz, cfg.w, 128)
    bA, bE = orc.Basis(act.basis), orc.Basis(env.basis)

    def ks(P):
        V, E, _, _ = orc.build_xc(bA, og, orc.Functional(*xc), P)
        return V, E

    def nadd(P, fn):
        V, E, _ = orc.build_nadd(bA, P, [(bE, env.P)], og, orc.Functional(*fn))
        return V, E

    want = [ks(act.P), nadd(act.P, xc), nadd(act.P, kin), ks(PA2), nadd(PA2, xc)]
    for (V, E), (Vr, Er) in zip(got, want):
        assert np.abs(V - Vr).max() <= 1e-8 and abs(E - Er) <= 1e-9
    from serenity_b200.inputs.basis import atom_indices_of_basis
    grad_ref = orc.xc_gradient(bA, og, orc.Functional(*xc), PA2, atom_indices_of_basis(act.basis, act.coords), len(act.symbols))
    assert np.abs(grad - grad_ref).max() <= 1e-9
    Vab_ref, _ = orc.build_ab(bA, bE, [(bA, PA2), (bE, env.P)], og, orc.Functional(*xc))
    assert np.abs(Vab - Vab_ref).max() <= 1e-8
    Vabn_ref = orc.build_ab_nadd(bA, bE, (bA, PA2), [(bE, env.P)], og, orc.Functional(*xc))
    assert np.abs(Vabn - Vabn_ref).max() <= 1e-8
    gradN_ref = orc.nadd_gradient(bA, PA2, [(bE, env.P)], og, orc.Functional(*xc), atom_indices_of_basis(act.basis, act.coords),
                                  len(act.symbols))
    assert np.abs(gradN - gradN_ref).max() <= 1e-9

    # row f-4: subsystem-TDDFT kernel sigma vector through Kernel / KernelSigmavector (the density is PA2 at that point)
    fx, fk = orc.Functional(*xc), orc.Functional(*kin)
    ra, ga, _, _ = orc.density_on_grid(bA, og, 1e-9, PA2, 1)
    rb, gb, _, _ = orc.density_on_grid(bE, og, 1e-9, env.P, 1)
    rt, gt = ra + rb, [x + y for x, y in zip(ga, gb)]
    tot = orc.kernel_store_r(fk, rt, gt, store=orc.kernel_store_r(fx, rt, gt))
    sub = orc.kernel_store_r(fx, ra, ga)
    sub = orc.kernel_store_r(fx, ra, ga, sign=-1.0, store=sub)
    sub = orc.kernel_store_r(fk, ra, ga, sign=-1.0, store=sub)
    resp_tot = orc.kernel_contract(bA, og, tot, DA, 0, True, block_ave_thr=0.0)
    resp_tot = orc.kernel_contract(bE, og, tot, DB, 0, True, resp=resp_tot, block_ave_thr=0.0)
    resp = orc.kernel_contract(bA, og, sub, DA, 0, True, resp=resp_tot.copy(), block_ave_thr=0.0)
    F_ref = orc.kernel_integrate(bA, og, resp, True, block_ave_thr=0.0)
    assert np.abs(F_fde - F_ref).max() <= 1e-10 * np.abs(F_ref).max()
    subB = orc.kernel_store_r(fx, rb, gb)
    subB = orc.kernel_store_r(fx, rb, gb, sign=-1.0, store=subB)
    subB = orc.kernel_store_r(fk, rb, gb, sign=-1.0, store=subB)
    resp = orc.kernel_contract(bE, og, subB, DB, 0, True, resp=resp_tot.copy(), block_ave_thr=0.0)
    F_ref = orc.kernel_integrate(bE, og, resp, True, block_ave_thr=0.0)
    assert np.abs(F_fde_B - F_ref).max() <= 1e-10 * np.abs(F_ref).max()
    iso = orc.kernel_store_r(fx, ra, ga)
    for F, D in ((F_iso0, DA), (F_iso1, PA2)):
        F_ref = orc.kernel_integrate(bA, og, orc.kernel_contract(bA, og, iso, D, 0, True, block_ave_thr=0.0), True, block_ave_thr=0.0)
        assert np.abs(F - F_ref).max() <= 1e-10 * np.abs(F_ref).max()
    assert np.abs(pp_block - iso[0, 256:384]).max() <= 1e-6 * np.abs(iso[0, 256:384]).max()
    # DensityOnGridCalculator -> FunctionalLibrary -> ScalarOperatorToMatrixAdder stand-ins = FuncPotential of the same density
    assert np.abs(V_stage - want[3][0]).max() <= 1e-8 and abs(E_stage - want[3][1]) <= 1e-9

    # round 2: getLinearizedEnergy, the projected-density constructor, the one-pass FDE bundle, and the multi-GPU group
    assert abs(lin - 0.5 * float((want[4][0] * PA2).sum())) <= 1e-9
    P_comb = PA2 + BtoA.T @ env.P @ BtoA       # (the combination is formed when the object is constructed: P_A is PA2 by then)
    Vc_ref, Ec_ref, _ = orc.build_nadd(bA, np.asfortranarray(P_comb), [(bE, env.P)], og, orc.Functional(*xc))
    assert np.abs(V_comb - Vc_ref).max() <= 1e-8 and abs(E_comb - Ec_ref) <= 1e-9
    Vx_ref, Ex_ref = nadd(PA2, xc)
    Vk_ref, Ek_ref = nadd(PA2, kin)
    assert np.abs(F_x - Vx_ref).max() <= 1e-8 and np.abs(F_k - Vk_ref).max() <= 1e-8
    assert abs(E_x - Ex_ref) <= 1e-9 and abs(E_k - Ek_ref) <= 1e-9
    assert np.abs(F_sum - (F_x + F_k)).max() <= 1e-14
    # mixed exact / approximate embedding kernel (A exact with naddXCExact = kin, B approximate with naddXCApprox = xc, naddKin = kin):
    #   tot = xc[rho_A + rho_B] + kin[rho_A + rho_B];  sub_A = xc[rho_A] - kin[rho_A];  sub_B = xc[rho_B] - xc[rho_B] - kin[rho_B];
    #   exact = kin[rho_A] - xc[rho_A] - kin[rho_A];   getPP(0,0) = tot + exact + sub_A,  getPP(0,1) = tot,  getPP(1,1) = tot + sub_B
    pp = lambda fn, r, g_: orc.kernel_store_r(fn, r, g_)[0, 256:384]
    tot_pp = pp(fx, rt, gt) + pp(fk, rt, gt)
    want_mixed = [tot_pp + (pp(fk, ra, ga) - pp(fx, ra, ga) - pp(fk, ra, ga)) + (pp(fx, ra, ga) - pp(fk, ra, ga)),
                  tot_pp, tot_pp + (pp(fx, rb, gb) - pp(fx, rb, gb) - pp(fk, rb, gb))]
    for got_pp, ref_pp in zip(pp_mixed, want_mixed):
        assert np.abs(got_pp - ref_pp).max() <= 1e-6 * max(np.abs(ref_pp).max(), 1e-30)
    if ngpu > 1:
        assert ng == ngpu
        assert np.abs(Vg - want[3][0]).max() <= 1e-8 and abs(Eg - want[3][1]) <= 1e-9
        assert np.abs(Fg - (Vx_ref + Vk_ref)).max() <= 1e-8
        assert np.abs(gradg - grad_ref).max() <= 1e-9


# ------------------------------------------------------------------------------------------------ B200Bridge.h
BRIDGE = os.path.join(ROOT, "tests", "cpp", "b200_bridge_test")


def _bridge():
    if not os.path.exists(BRIDGE):
        from serenity_b200 import build
        build.build_bridge_test()
    return BRIDGE


def test_bridge_header_compiles_against_the_reference_accessor_names():
    """serenity_b200/host/B200Bridge.h (the file INTEGRATION.md section 3 binds through) compiles against classes that expose
    exactly the GridController / BasisController / Shell / Functional accessors it cites, and links against the C ABI."""
    from serenity_b200 import build
    build.build_bridge_test()
    r = subprocess.run([_bridge(), "--compile-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "C ABI version" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("ngpu", [1, 2])
def test_bridge_builds_vxc_like_funcpotential(tmp_path, ngpu):
    """FuncPotential::getMatrix as INTEGRATION.md section 3 writes it (B200Bridge: handle caches + sxc_build_xc, or the sxc_group
    of a multi-GPU host) against the oracle; a Grid notify() (forgetGrid) re-uploads and reproduces the matrix."""
    import torch
    if torch.cuda.device_count() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    cfg = make_config("water8", 3)
    sub = cfg.subsystems[0]
    ids, mix = FUNCTIONALS["PBE"]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        _w(f, cfg.xyz, np.float64)
        _w(f, cfg.w, np.float64)
        _w(f, ids, np.int32)
        _w(f, mix, np.float64)
        _basis(f, sub.basis, sub.coords)
        _w(f, np.asfortranarray(sub.P).reshape(-1, order="F"), np.float64)
    r = subprocess.run([_bridge(), fin, fout, str(ngpu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    nb = sub.basis.nbf
    with open(fout, "rb") as f:
        V, en, V2, ng = _r(f, (nb, nb)), _r(f, (2,)), _r(f, (nb, nb)), _r(f)
    V_ref, E_ref, ne_ref, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), sub.P)
    assert ng == ngpu
    assert np.abs(V - V_ref).max() <= 1e-8 and abs(en[0] - E_ref) <= 1e-9 and abs(en[1] - ne_ref) <= 1e-10 * abs(ne_ref)
    assert np.abs(V2 - V).max() <= 1e-12
