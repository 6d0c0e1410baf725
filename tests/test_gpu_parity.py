"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's golden vectors.

Bars (BASELINE.json north_star): |dE_xc| <= 1e-9 Eh, max|dV_xc| <= 1e-8, FP64.  Stage-level quantities are
compared much tighter (1e-12 relative) because they only differ by summation order / FMA contraction.
"""
import numpy as np
import pytest

from conftest import grid_arrays, load_golden

pytestmark = pytest.mark.gpu

E_TOL = 1e-9   # Eh
V_TOL = 1e-8   # max-abs


@pytest.fixture(scope="module")
def ctx():
    from serenity_b200.xc import XCContext
    c = XCContext(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    return pyoracle


def _cfg(name, acc=None):
    from serenity_b200.inputs import make_config
    return make_config(name, acc)


def _functional(name):
    from serenity_b200.inputs.configs import FUNCTIONALS
    return FUNCTIONALS[name]


# ------------------------------------------------------------------------------------------- golden vectors
def test_golden_basis_functions(ctx, fixtures, small_mixed):
    """BasisFunctionOnGridController_test.cpp:43-483 (phi and grad phi; Cartesian s, p, d shells)."""
    gold = load_golden("basis_functions_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    g = ctx.set_grid(xyz, w, gold["block_size"])
    b = ctx.add_basis(small_mixed, gold["radial_threshold"])
    arrs, neg, n = ctx.basis_on_grid(g, b, 0, small_mixed.nbf, gold["block_size"])
    assert n == 4 and not neg.any()
    for name, arr in zip(["values", "dx", "dy", "dz"], arrs):
        for p, mu, ref in gold["entries"][name]:
            assert abs(arr[p, mu] - ref) < gold["tolerance"], (name, p, mu, arr[p, mu], ref)


def test_golden_density(ctx, fixtures, small_mixed):
    """DensityOnGridCalculator_test.cpp:43-254 (rho, grad rho; block size 3 -> two ragged blocks)."""
    gold = load_golden("density_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    g = ctx.set_grid(xyz, w, gold["block_size"])
    b = ctx.add_basis(small_mixed, 1e-300)  # the test uses threshold 0: nothing is screened
    rho, grad = ctx.density_on_grid(g, b, np.asarray(gold["P"]), 4)
    exp = gold["expected"]
    assert np.allclose(rho, exp["rho"], rtol=0, atol=gold["tolerance"])
    for k, c in enumerate("xyz"):
        assert np.allclose(grad[k], exp["d" + c], rtol=0, atol=gold["tolerance"])


def test_golden_basis_function_hessians(ctx, fixtures, small_mixed):
    """BasisFunctionOnGridController_test.cpp:84-483, derivative level 2: the 240 second-derivative values on the device."""
    gold = load_golden("basis_functions_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    g = ctx.set_grid(xyz, w, gold["block_size"])
    b = ctx.add_basis(small_mixed, gold["radial_threshold"])
    arrs, n = ctx.basis_hessian_on_grid(g, b, 0, small_mixed.nbf, gold["block_size"])
    assert n == 4
    checked = 0
    for name, arr in zip(["hxx", "hxy", "hxz", "hyy", "hyz", "hzz"], arrs):
        for p, mu, ref in gold["entries"][name]:
            assert abs(arr[p, mu] - ref) < gold["tolerance"], (name, p, mu, arr[p, mu], ref)
            checked += 1
    assert checked == 240


def test_golden_density_hessian(ctx, fixtures, small_mixed):
    """DensityOnGridCalculator_test.cpp:174-253: the 24 second derivatives of the density (block size 3, threshold 0)."""
    gold = load_golden("density_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    g = ctx.set_grid(xyz, w, gold["block_size"])
    b = ctx.add_basis(small_mixed, 1e-300)
    hess = ctx.density_hessian_on_grid(g, b, np.asarray(gold["P"]), 4)
    exp = gold["expected"]
    for name, h in zip(["hxx", "hxy", "hxz", "hyy", "hyz", "hzz"], hess):
        assert np.allclose(h, exp[name], rtol=0, atol=gold["tolerance"]), (name, h, exp[name])


def test_hessians_vs_oracle_all_l(ctx, orc):
    """Second derivatives of spherical l = 0..6 and Cartesian l = 0..4 shells and of a density against the oracle (level 2)."""
    from serenity_b200.inputs.basis import shell_table_from_list
    rng = np.random.default_rng(12)
    pts = rng.uniform(-1.5, 1.5, size=(200, 3))
    w = rng.uniform(0.1, 1.0, size=len(pts))
    shells = [{"l": l, "pure": True, "exps": [0.3 + 0.1 * l, 1.1], "coefs": [0.7, 0.4], "centre": [0.1 * l, -0.2, 0.3]}
              for l in range(7)]
    shells += [{"l": l, "pure": False, "exps": [0.4], "coefs": [1.0], "centre": [0.0, 0.1, -0.1]} for l in range(5)]
    tab = shell_table_from_list(shells)
    g = ctx.set_grid(pts, w, 128)
    b = ctx.add_basis(tab, 1e-9)
    og, ob = orc.Grid(pts, w, 128), orc.Basis(tab)
    for blk in range(2):
        arrs, n = ctx.basis_hessian_on_grid(g, b, blk, tab.nbf, 128)
        ref, _, _ = orc.basis_block(ob, og, 1e-9, 2, blk)
        for a, r in zip(arrs, ref[4:10]):
            assert np.allclose(a, r, rtol=1e-11, atol=1e-13)
    P = rng.standard_normal((tab.nbf, tab.nbf))
    P = P + P.T
    hess = ctx.density_hessian_on_grid(g, b, P, len(pts))
    _, _, want, _ = orc.density_on_grid(ob, og, 1e-9, P, 2)
    for h, r in zip(hess, want):
        assert np.allclose(h, r, rtol=1e-11, atol=1e-11)


def test_supersystem_density_is_the_exact_sum(ctx, fixtures, small_mixed):
    """SupersystemDensityOnGridController_test.cpp:43-90: a supersystem of twice the same subsystem has EXACTLY rho + rho and
    grad rho + grad rho on every point (EXPECT_EQ in the reference: bit-for-bit), and a third copy adds once more."""
    gold = load_golden("density_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    g = ctx.set_grid(xyz, w, gold["block_size"])
    b = ctx.add_basis(small_mixed, 1e-300)
    P = np.asarray(gold["P"])
    rho, grad = ctx.density_on_grid(g, b, P, 4)
    rho2, grad2 = ctx.supersystem_density_on_grid(g, [b, b], [P, P], 4)
    assert np.array_equal(rho2, rho + rho)
    for k in range(3):
        assert np.array_equal(grad2[k], grad[k] + grad[k])
    rho3, grad3 = ctx.supersystem_density_on_grid(g, [b, b, b], [P, P, P], 4)
    assert np.array_equal(rho3, (rho + rho) + rho)
    # two different subsystems on a molecular grid: the sum of the separately evaluated densities, in the order given
    from serenity_b200.inputs import make_config
    cfg = make_config("fde_dimer", 2)
    act, env = cfg.subsystems
    gg = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba, be = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    ra, ga = ctx.density_on_grid(gg, ba, act.P, cfg.npts)
    re_, ge = ctx.density_on_grid(gg, be, env.P, cfg.npts)
    rs, gs = ctx.supersystem_density_on_grid(gg, [ba, be], [act.P, env.P], cfg.npts)
    assert np.array_equal(rs, ra + re_)
    for k in range(3):
        assert np.array_equal(gs[k], ga[k] + ge[k])


def test_density_is_bitwise_reproducible(ctx):
    """k_density sums in a fixed order (its mbarrier-ordered cp.async ring is what compute-sanitizer's racecheck cannot follow,
    profiles/r02_sanitizer.md): the same build twice gives the same bits, on a grid large enough for many CTAs per SM."""
    cfg = _cfg("tetracene", 4)
    assert (cfg.npts + 127) // 128 >= 900  # whole-block work items (a smaller shard is cut into segments that accumulate with atomics)
    sub = cfg.subsystems[0]
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    runs = [ctx.density_on_grid(g, b, sub.P, cfg.npts) for _ in range(3)]
    for rho, grad in runs[1:]:
        assert np.array_equal(rho, runs[0][0])
        for k in range(3):
            assert np.array_equal(grad[k], runs[0][1][k])


def test_golden_scalar_to_matrix(ctx, fixtures, small_mixed):
    """ScalarOperatorToMatrixAdder_test.cpp:41-148 (55 elements, GGA variant, block size 3)."""
    gold = load_golden("scatter_ref.json")
    xyz, w = grid_arrays(fixtures, "VERY_SMALL")
    g = ctx.set_grid(xyz, w, gold["block_size"])
    b = ctx.add_basis(small_mixed, gold["radial_threshold"])
    pot = gold["potential"]
    V = ctx.scalar_to_matrix(g, b, small_mixed.nbf, np.asarray(pot["pot"]), np.asarray(pot["gx"]), np.asarray(pot["gy"]),
                             np.asarray(pot["gz"]), gold["block_ave_threshold"])
    for i, j, ref in gold["entries"]:
        assert abs(V[i, j] - ref) < gold["tolerance"], (i, j, V[i, j], ref)
    assert np.array_equal(V, V.T)


def test_harmonics_all_l(ctx, orc):
    """Spherical shells l = 0..6 (straight-line code for l <= 3, table-driven above) vs the oracle."""
    from serenity_b200.inputs.basis import shell_table_from_list
    rng = np.random.default_rng(11)
    pts = rng.uniform(-1.5, 1.5, size=(200, 3))
    w = np.ones(len(pts))
    shells = [{"l": l, "pure": True, "exps": [0.3 + 0.1 * l, 1.1], "coefs": [0.7, 0.4], "centre": [0.1 * l, -0.2, 0.3]}
              for l in range(7)]
    shells += [{"l": l, "pure": False, "exps": [0.4], "coefs": [1.0], "centre": [0.0, 0.1, -0.1]} for l in range(5)]
    tab = shell_table_from_list(shells)
    g = ctx.set_grid(pts, w, 128)
    b = ctx.add_basis(tab, 1e-9)
    og, ob = orc.Grid(pts, w, 128), orc.Basis(tab)
    for blk in range(2):
        arrs, neg, n = ctx.basis_on_grid(g, b, blk, tab.nbf, 128)
        ref, rneg, _ = orc.basis_block(ob, og, 1e-9, 1, blk)
        assert np.array_equal(neg, rneg)
        for a, r in zip(arrs, ref):
            assert np.allclose(a, r, rtol=1e-12, atol=1e-14)


# ------------------------------------------------------------------------------------------- functionals
@pytest.mark.parametrize("fid", [2, 45, 66, 80, 81, 135, 184, 193, 197, 283, 286])
def test_functional_kernels_vs_oracle(ctx, orc, fid):
    rng = np.random.default_rng(fid)
    n = 1000
    rho = 10.0 ** rng.uniform(-9, 2.3, size=n)
    s = 10.0 ** rng.uniform(-3, 1.0, size=n)
    gnorm = s * 2.0 * (3 * np.pi ** 2) ** (1 / 3) * rho ** (4 / 3)
    u = rng.normal(size=(3, n))
    u /= np.linalg.norm(u, axis=0)
    gx, gy, gz = (np.ascontiguousarray(gnorm * u[k]) for k in range(3))
    rho[::97] = 3e-15     # below the tiny-density cut
    w = rng.uniform(0.1, 1.0, size=n)
    f = ctx.set_functional([fid], [1.0])
    e, out = ctx.functional_on_grid(f, w, rho, gx, gy, gz)
    e_ref, ref = orc.functional_on_grid(orc.Functional([fid], [1.0]), w, rho, gx, gy, gz)
    for k in range(5):
        scale = np.maximum(np.abs(ref[k]), 1e-300)
        err = np.abs(out[k] - ref[k])
        assert np.all((err / scale < 5e-11) | (err < 1e-16)), (fid, k, float((err / scale).max()))
    assert abs(e - e_ref) <= 1e-11 * max(1.0, abs(e_ref))


def test_functional_block_skip_and_ragged(ctx, orc):
    f = ctx.set_functional([135, 197], [1.0, 1.0])
    n = 300  # 128 + 128 + 44
    rho = np.full(n, 1e-3)
    rho[:128] = 5e-13
    rho[130] = 5e-15
    g = [np.full(n, 1e-4) for _ in range(3)]
    w = np.linspace(0.5, 1.5, n)
    e, out = ctx.functional_on_grid(f, w, rho, *g)
    e_ref, ref = orc.functional_on_grid(orc.Functional([135, 197], [1.0, 1.0]), w, rho, *g)
    assert np.all(out[0][:128] == 0) and np.all(out[1][:128] == 0) and out[1][130] == 0
    for k in range(5):
        assert np.allclose(out[k], ref[k], rtol=1e-11, atol=1e-18)
    assert abs(e - e_ref) < 1e-13


# ------------------------------------------------------------------------------------------- whole builds
def _compare_build(ctx, orc, cfg, func_name, blocksize=128, ws_limit=None):
    sub = cfg.subsystems[0]
    ids, mix = _functional(func_name)
    ob, og, of = orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, blocksize), orc.Functional(ids, mix)
    # scale P so that the grid integrates to N_el
    rho, _, _, _ = orc.density_on_grid(ob, og, 1e-9, sub.P, deriv=0)
    P = sub.P * (sub.n_electrons / float(rho @ cfg.w))
    V_ref, E_ref, ne_ref, _ = orc.build_xc(ob, og, of, P)
    if ws_limit:
        ctx.set_workspace_limit(ws_limit)
    g = ctx.set_grid(cfg.xyz, cfg.w, blocksize)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    V, E, ne = ctx.build_xc(g, b, f, P)
    if ws_limit:
        ctx.set_workspace_limit(0)
    st = ctx.stats()
    assert abs(E - E_ref) <= E_TOL, (E, E_ref)
    assert np.abs(V - V_ref).max() <= V_TOL, np.abs(V - V_ref).max()
    assert abs(ne - ne_ref) <= 1e-10 * abs(ne_ref)
    assert np.array_equal(V, V.T)
    return st, (V, E, ne), P


@pytest.mark.parametrize("func_name", ["PBE", "LDA", "B3LYP", "BP86", "BLYP", "PW91K", "TF"])
def test_h2o_build_matches_oracle(ctx, orc, func_name):
    """cfg 1: H2O / def2-SVP, grid accuracy 4 (SCF grid accuracy 2 in the second test)."""
    _compare_build(ctx, orc, _cfg("h2o"), func_name)


def test_h2o_small_grid_and_odd_blocksize(ctx, orc):
    cfg = _cfg("h2o", 2)
    _compare_build(ctx, orc, cfg, "PBE")
    _compare_build(ctx, orc, cfg, "PBE", blocksize=100)


def test_water8_screening_and_chunking(ctx, orc):
    """(H2O)8: block screening is active; identical negligible sets, multi-chunk pipeline gives the same answer."""
    cfg = _cfg("water8")
    st1, r1, P = _compare_build(ctx, orc, cfg, "PBE")
    assert st1["s_max"] <= st1["nbf"] and st1["sum_s"] < st1["nblocks"] * st1["nbf"]  # something was screened
    st2, r2, _ = _compare_build(ctx, orc, cfg, "PBE", ws_limit=64 << 20)
    assert st2["nchunks"] > 1
    assert abs(r1[1] - r2[1]) < 1e-11 and np.abs(r1[0] - r2[0]).max() < 1e-11
    # negligible flags of a few blocks against the oracle
    sub = cfg.subsystems[0]
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    ob, og = orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128)
    nblocks = og.nblocks
    for blk in [0, nblocks // 3, nblocks // 2, nblocks - 1]:
        arrs, neg, n = ctx.basis_on_grid(g, b, blk, sub.basis.nbf, 128)
        ref, rneg, _ = orc.basis_block(ob, og, 1e-9, 1, blk)
        assert np.array_equal(neg, rneg), blk
        for a, r in zip(arrs, ref):
            assert np.allclose(a, r, rtol=1e-12, atol=1e-15)


def test_far_points_give_empty_blocks(ctx, orc):
    """Blocks without any significant function (s_b = 0) leave rho = 0 and contribute nothing (Appendix E.5)."""
    cfg = _cfg("h2o", 2)
    far = np.tile(np.array([[200.0, 150.0, -300.0]]), (256, 1)) + np.random.default_rng(1).uniform(-1, 1, (256, 3))
    xyz = np.concatenate([cfg.xyz, far])
    w = np.concatenate([cfg.w, np.full(256, 0.01)])
    cfg.xyz, cfg.w = np.ascontiguousarray(xyz), w
    st, _, _ = _compare_build(ctx, orc, cfg, "PBE")


def test_shards_sum_to_full_build(ctx, orc):
    """Two shards (rank 0/2, 1/2) on one GPU: partial V/E/N add up to the unsharded result."""
    from serenity_b200.xc import XCContext
    cfg = _cfg("water8")
    sub = cfg.subsystems[0]
    ids, mix = _functional("PBE")
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    V, E, ne = ctx.build_xc(g, b, f, sub.P)
    tot_V, tot_E, tot_n, npts = 0.0, 0.0, 0.0, 0
    for rank in range(2):
        c2 = XCContext(0)
        g2 = c2.set_grid(cfg.xyz, cfg.w, 128)
        c2.set_grid_shard(g2, rank, 2)
        b2 = c2.add_basis(sub.basis, 1e-9)
        f2 = c2.set_functional(ids, mix)
        Vp, Ep, nep = c2.build_xc(g2, b2, f2, sub.P)
        npts += c2.stats()["npts"]
        tot_V, tot_E, tot_n = tot_V + Vp, tot_E + Ep, tot_n + nep
        c2.close()
    assert npts == cfg.npts
    assert abs(tot_E - E) < 1e-11 and abs(tot_n - ne) < 1e-10 and np.abs(tot_V - V).max() < 1e-11


def test_nadd_dimer_matches_oracle(ctx, orc):
    """cfg 4 (small case): water dimer, NAdd-XC = PBE and NAdd-kin = PW91k on the supersystem grid."""
    cfg = _cfg("fde_dimer")
    act, env = cfg.subsystems
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA = ctx.add_basis(act.basis, 1e-9)
    bE = ctx.add_basis(env.basis, 1e-9)
    og = orc.Grid(cfg.xyz, cfg.w, 128)
    for name in ["PBE", "PW91K", "LDA"]:
        ids, mix = _functional(name)
        f = ctx.set_functional(ids, mix)
        V, E = ctx.build_nadd(g, f, bA, act.P, [bE], [env.P])
        V_ref, E_ref, parts = orc.build_nadd(orc.Basis(act.basis), act.P, [(orc.Basis(env.basis), env.P)], og,
                                             orc.Functional(ids, mix))
        assert np.abs(V - V_ref).max() <= V_TOL, (name, np.abs(V - V_ref).max())
        assert np.allclose(E, parts, rtol=0, atol=E_TOL), (name, E, parts)
        assert abs((E[0] - E[1] - E[2]) - E_ref) <= E_TOL
        # frozen environment: second call reuses the cached environment density and gives the same answer
        V2, E2 = ctx.build_nadd(g, f, bA, act.P, [bE], [env.P], env_frozen=True)
        assert np.abs(V2 - V).max() < 1e-12 and np.abs(E2 - E).max() < 1e-12


def test_potential_classes_lazy_evaluation(ctx):
    from serenity_b200.xc import DensityMatrixController, FuncPotential
    cfg = _cfg("h2o", 2)
    sub = cfg.subsystems[0]
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(*_functional("PBE"))
    dmat = DensityMatrixController(sub.P)
    pot = FuncPotential(ctx, g, b, dmat, f)
    V1 = pot.getMatrix()
    assert pot.getMatrix() is V1              # cached until notify (FuncPotential.h:107-109)
    E1 = pot.getEnergy(sub.P)
    dmat.setDensityMatrix(sub.P * 1.1)        # notifies
    V2 = pot.getMatrix()
    assert V2 is not V1 and pot.getEnergy(None) != E1


# ------------------------------------------------------------------------------------------- full-size properties
def test_tetracene_full_size_properties(ctx, orc):
    """cfg 2 at BASELINE size (B3LYP/def2-TZVP, ~4.9e5 points): oracle parity on E and V plus size-independent
    properties - symmetry, linearity of N_el in P, additivity of V and E over grid halves."""
    cfg = _cfg("tetracene")
    sub = cfg.subsystems[0]
    ids, mix = _functional("B3LYP")
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    V, E, ne = ctx.build_xc(g, b, f, sub.P)
    P = sub.P * (sub.n_electrons / ne)
    V, E, ne = ctx.build_xc(g, b, f, P)
    assert abs(ne - sub.n_electrons) < 1e-8           # linearity of rho in P
    assert np.array_equal(V, V.T) and np.isfinite(V).all()
    half = (cfg.npts // 256) * 128
    parts = []
    for sl in (slice(0, half), slice(half, None)):
        gi = ctx.set_grid(cfg.xyz[sl], cfg.w[sl], 128)
        parts.append(ctx.build_xc(gi, b, f, P))
    assert abs(parts[0][1] + parts[1][1] - E) < 1e-9
    assert np.abs(parts[0][0] + parts[1][0] - V).max() < 1e-9
    V_ref, E_ref, ne_ref, _ = orc.build_xc(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), P)
    assert abs(E - E_ref) <= E_TOL and np.abs(V - V_ref).max() <= V_TOL and abs(ne - ne_ref) < 1e-8


# ------------------------------------------------------------------------------------------- UNRESTRICTED
def _open_shell_P(sub, seed=21):
    """P_alpha with one occupied orbital more than P_beta (random PSD matrices of the closed-shell magnitude)."""
    nb, nocc = sub.basis.nbf, sub.n_electrons // 2
    rng = np.random.default_rng(seed)
    C = rng.normal(size=(nb, nocc + 1)) / np.sqrt(nb)
    Pa = np.asfortranarray(C @ C.T)
    Pb = np.asfortranarray(C[:, : nocc - 1] @ C[:, : nocc - 1].T)
    return Pa, Pb


@pytest.mark.parametrize("fid", [2, 45, 66, 80, 135, 184, 193, 197, 283, 286])
def test_unrestricted_functional_kernels_vs_oracle(ctx, orc, fid):
    rng = np.random.default_rng(100 + fid)
    n = 700
    rho = 10.0 ** rng.uniform(-9, 2.0, size=n)
    zeta = rng.uniform(-1.0, 1.0, size=n)
    zeta[::50] = 1.0                      # fully polarised points: rho_b = 0 is raised to the tiny density
    rho2 = np.stack([0.5 * rho * (1 + zeta), 0.5 * rho * (1 - zeta)])
    rho2[:, 5::97] = 2e-15                # below the tiny-density cut
    grad = rng.normal(size=(2, 3, n)) * (rho2[:, None, :] + 1e-12) ** (4 / 3)
    w = rng.uniform(0.1, 1.0, size=n)
    f = ctx.set_functional([fid], [1.0])
    e, ep, vr, vg = ctx.functional_on_grid_u(f, w, rho2, grad)
    e_ref, ep_r, vr_r, vg_r = orc.functional_on_grid_u(orc.Functional([fid], [1.0]), w, rho2, grad)
    for got, ref in ((ep, ep_r), (vr, vr_r), (vg, vg_r)):
        scale = np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max(axis=tuple(range(ref.ndim - 1)), keepdims=True) + 1e-300)
        err = np.abs(got - ref)
        assert np.all((err / scale < 1e-9) | (err < 1e-15)), (fid, float((err / scale).max()))
    assert abs(e - e_ref) <= 1e-11 * max(1.0, abs(e_ref))


@pytest.mark.parametrize("func_name", ["PBE", "B3LYP", "LDA", "BP86", "PW91K"])
def test_h2o_unrestricted_build_matches_oracle(ctx, orc, func_name):
    """FuncPotential<UNRESTRICTED>: {V_alpha, V_beta}, E_xc for an open-shell density pair."""
    cfg = _cfg("h2o")
    sub = cfg.subsystems[0]
    ids, mix = _functional(func_name)
    Pa, Pb = _open_shell_P(sub)
    (Va_r, Vb_r), E_r, ne_r = orc.build_xc_u(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix),
                                             Pa, Pb)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    (Va, Vb), E, ne = ctx.build_xc(g, b, f, (Pa, Pb), nspin=2)
    assert abs(E - E_r) <= E_TOL and abs(ne - ne_r) <= 1e-10 * abs(ne_r)
    assert np.abs(Va - Va_r).max() <= V_TOL and np.abs(Vb - Vb_r).max() <= V_TOL
    assert np.array_equal(Va, Va.T) and np.array_equal(Vb, Vb.T) and np.abs(Va - Vb).max() > 1e-4


def test_unrestricted_reduces_to_restricted(ctx):
    """P_alpha = P_beta = P/2 gives V_alpha = V_beta = V_restricted and the same E_xc."""
    cfg = _cfg("water8")
    sub = cfg.subsystems[0]
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(*_functional("PBE"))
    V, E, ne = ctx.build_xc(g, b, f, sub.P)
    (Va, Vb), Eu, neu = ctx.build_xc(g, b, f, (0.5 * sub.P, 0.5 * sub.P), nspin=2)
    assert abs(E - Eu) < 1e-10 and abs(ne - neu) < 1e-10
    assert np.abs(Va - V).max() < 1e-10 and np.abs(Vb - V).max() < 1e-10


def test_unrestricted_empty_beta_channel(ctx, orc):
    """One-electron-like case: P_beta = 0.  The beta blocks fail the per-spin block test in the density-dependent part
    (Appendix E.9) but the functional is still evaluated because alpha is not negligible."""
    cfg = _cfg("h2o", 2)
    sub = cfg.subsystems[0]
    ids, mix = _functional("PBE")
    Pa = np.asfortranarray(0.5 * sub.P)
    Pb = np.zeros_like(Pa)
    (Va_r, Vb_r), E_r, _ = orc.build_xc_u(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), Pa, Pb)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    (Va, Vb), E, _ = ctx.build_xc(g, b, f, (Pa, Pb), nspin=2)
    assert abs(E - E_r) <= E_TOL and np.abs(Va - Va_r).max() <= V_TOL and np.abs(Vb - Vb_r).max() <= V_TOL


def test_nadd_unrestricted_dimer_matches_oracle(ctx, orc):
    """NAddFuncPotential<UNRESTRICTED> on the water dimer: open-shell active system, closed-shell environment."""
    cfg = _cfg("fde_dimer")
    act, env = cfg.subsystems
    PA = _open_shell_P(act)
    PE = (np.asfortranarray(0.5 * env.P), np.asfortranarray(0.5 * env.P))
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA = ctx.add_basis(act.basis, 1e-9)
    bE = ctx.add_basis(env.basis, 1e-9)
    og = orc.Grid(cfg.xyz, cfg.w, 128)
    for name in ["PBE", "PW91K"]:
        ids, mix = _functional(name)
        f = ctx.set_functional(ids, mix)
        (Va, Vb), E = ctx.build_nadd(g, f, bA, PA, [bE], [PE], nspin=2)
        (Va_r, Vb_r), E_ref, parts = orc.build_nadd_u(orc.Basis(act.basis), PA, [(orc.Basis(env.basis), PE)], og,
                                                      orc.Functional(ids, mix))
        assert np.abs(Va - Va_r).max() <= V_TOL and np.abs(Vb - Vb_r).max() <= V_TOL, name
        assert np.allclose(E, parts, rtol=0, atol=E_TOL) and abs((E[0] - E[1] - E[2]) - E_ref) <= E_TOL
        (Va2, Vb2), E2 = ctx.build_nadd(g, f, bA, PA, [bE], [PE], env_frozen=True, nspin=2)
        assert np.abs(Va2 - Va).max() < 1e-12 and np.abs(E2 - E).max() < 1e-12


def test_tile_cache_reproduces_uncached_builds_and_is_invalidated_by_other_work(ctx, orc):
    """sxc_set_tile_cache: the second and later builds of the same (grid, basis) skip k_screen / k_basis and must reproduce
    the uncached result (E_xc and N_el bit for bit - their sums are ordered; V to the 1e-13 of its FP64 red.global
    accumulation order) for RESTRICTED, UNRESTRICTED and NAdd builds; a build with another basis in between refills the
    workspace, so the next cached-mode build re-evaluates its own tiles."""
    cfg = _cfg("fde_dimer")
    act, env = cfg.subsystems
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA, bE = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    f = ctx.set_functional(*_functional("PBE"))
    V0, E0, n0 = ctx.build_xc(g, bA, f, act.P)
    launches_full = ctx.stats()["kernel_launches"]
    Vu0, Eu0, _ = ctx.build_xc(g, bA, f, (0.6 * act.P, 0.4 * act.P), nspin=2)
    Vn0, En0 = ctx.build_nadd(g, f, bA, act.P, [bE], [env.P])
    ctx.set_tile_cache(True)
    same = lambda a, b: np.abs(a - b).max() <= 1e-13  # noqa: E731
    try:
        V1, E1, n1 = ctx.build_xc(g, bA, f, act.P)            # fills the cache
        V2, E2, n2 = ctx.build_xc(g, bA, f, act.P)            # served from it
        assert ctx.stats()["kernel_launches"] == launches_full - 1   # no k_basis (k_screen runs once, with the plan)
        assert same(V1, V0) and same(V2, V0) and E1 == E0 == E2 and n2 == n0
        P2 = act.P * 1.03
        Vp, Ep, _ = ctx.build_xc(g, bA, f, P2)
        ctx.set_tile_cache(False)
        Vq, Eq, _ = ctx.build_xc(g, bA, f, P2)
        ctx.set_tile_cache(True)
        assert same(Vp, Vq) and Ep == Eq
        Vu, Eu, _ = ctx.build_xc(g, bA, f, (0.6 * act.P, 0.4 * act.P), nspin=2)
        assert same(Vu[0], Vu0[0]) and same(Vu[1], Vu0[1]) and Eu == Eu0
        # another basis takes the workspace; the active system's next build must not trust stale tiles
        Ve, Ee, _ = ctx.build_xc(g, bE, f, env.P)
        V3, E3, _ = ctx.build_xc(g, bA, f, act.P)
        assert ctx.stats()["kernel_launches"] == launches_full
        assert same(V3, V0) and E3 == E0
        # freeze-and-thaw: frozen environment, the active tiles stay valid from call to call
        Vn1, En1 = ctx.build_nadd(g, f, bA, act.P, [bE], [env.P], env_frozen=True)
        Vn2, En2 = ctx.build_nadd(g, f, bA, act.P, [bE], [env.P], env_frozen=True)
        assert same(Vn1, Vn0) and same(Vn2, Vn0) and np.array_equal(En1, En0) and np.array_equal(En2, En0)
    finally:
        ctx.set_tile_cache(False)


def test_nadd_frozen_environment_cache_serves_alternating_functionals(ctx, orc):
    """One freeze-and-thaw iteration calls the XC and the kinetic NAddFuncPotential in turn (FDEPotentials.cpp:43-61): the frozen
    environment's summed density is shared, its energies are cached per functional - both objects must keep returning their own
    E[rho_env] and the same V as an uncached call."""
    cfg = _cfg("fde_dimer")
    act, env = cfg.subsystems
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    bA, bE = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    fx, fk = ctx.set_functional(*_functional("PBE")), ctx.set_functional(*_functional("PW91K"))
    ref = {f: ctx.build_nadd(g, f, bA, act.P, [bE], [env.P], env_frozen=False) for f in (fx, fk)}
    assert abs(ref[fx][1][2] - ref[fk][1][2]) > 1e-3       # E_xc[rho_env] and T_s[rho_env] differ
    for _ in range(3):
        for f in (fx, fk):
            V, E = ctx.build_nadd(g, f, bA, act.P, [bE], [env.P], env_frozen=True)
            assert np.array_equal(E, ref[f][1]) and np.abs(V - ref[f][0]).max() <= 1e-13
    launches = ctx.stats()["kernel_launches"]
    V, E = ctx.build_nadd(g, fx, bA, act.P, [bE], [env.P], env_frozen=False)
    assert ctx.stats()["kernel_launches"] > launches        # an unfrozen call recomputes the environment
    assert np.array_equal(E, ref[fx][1])
