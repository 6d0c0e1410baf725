"""world_size-2 gloo test of the multi-GPU host layer (serenity_b200/sharded.py) on CPU.

Each rank evaluates its contiguous block range with the CPU oracle standing in for the CUDA library (the local builder is
injectable); ShardedBuild's single all-reduce of [V | E | N] must reproduce the unsharded oracle build."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        from serenity_b200.inputs import make_config
        from serenity_b200.inputs.configs import FUNCTIONALS
        from serenity_b200.sharded import ShardedBuild, shard_bounds
        cfg = make_config("h2o", 2)
        sub = cfg.subsystems[0]
        ids, mix = FUNCTIONALS["PBE"]
        nbf = sub.basis.nbf
        nblk = (cfg.npts + 127) // 128
        bounds = shard_bounds(np.ones(nblk), world)
        lo, hi = int(bounds[rank]) * 128, min(int(bounds[rank + 1]) * 128, cfg.npts)
        ob, of = orc.Basis(sub.basis), orc.Functional(ids, mix)

        def local_build(d_P, d_VEN, p_ready=None):
            P = d_P.numpy().reshape(nbf, nbf, order="F")
            V, E, ne, _ = orc.build_xc(ob, orc.Grid(cfg.xyz[lo:hi], cfg.w[lo:hi], 128), of, P)
            d_VEN[: nbf * nbf] = torch.from_numpy(V.reshape(-1, order="F").copy())
            d_VEN[nbf * nbf] = E
            d_VEN[nbf * nbf + 1] = ne

        sb = ShardedBuild(nbf, local_build, "cpu")
        assert sb.world == world
        V, E, ne = sb.build(sub.P)
        V2, E2, ne2 = sb.build(sub.P)  # buffers are reused: a second build must not accumulate
        assert abs(E2 - E) < 1e-12 and np.abs(V - V2).max() < 1e-12  # (oracle sums in OpenMP order)
        assert np.array_equal(sb.d_P.numpy(), sub.P.reshape(-1, order="F"))  # slices + all-gather rebuilt P everywhere
        if rank == 0:
            V_ref, E_ref, ne_ref, _ = orc.build_xc(ob, orc.Grid(cfg.xyz, cfg.w, 128), of, sub.P)
            np.savez(out_path, dV=np.abs(V - V_ref).max(), dE=abs(E - E_ref), dn=abs(ne - ne_ref), sym=np.abs(V - V.T).max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_build_allreduce_gloo(tmp_path, world):
    port = 29600 + os.getpid() % 300
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r = np.load(out)
    assert r["dE"] < 1e-11 and r["dV"] < 1e-11 and r["dn"] < 1e-11 and r["sym"] < 1e-13


def _worker_u(rank, world, port, out_path):
    """UNRESTRICTED: ShardedBuild(nspin=2) moves {alpha, beta} pairs of P and V (2 nb^2 + 2 doubles all-reduced)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        from serenity_b200.inputs import make_config
        from serenity_b200.inputs.configs import FUNCTIONALS
        from serenity_b200.sharded import ShardedBuild, shard_bounds
        cfg = make_config("h2o", 2)
        sub = cfg.subsystems[0]
        ids, mix = FUNCTIONALS["PBE"]
        nbf = sub.basis.nbf
        n2 = nbf * nbf
        nblk = (cfg.npts + 127) // 128
        bounds = shard_bounds(np.ones(nblk), world)
        lo, hi = int(bounds[rank]) * 128, min(int(bounds[rank + 1]) * 128, cfg.npts)
        ob, of = orc.Basis(sub.basis), orc.Functional(ids, mix)
        Pa, Pb = 0.6 * sub.P, 0.4 * sub.P + 0.01 * np.eye(nbf)

        def local_build(d_P, d_VEN, p_ready=None):
            P = d_P.numpy()
            (Va, Vb), E, ne = orc.build_xc_u(ob, orc.Grid(cfg.xyz[lo:hi], cfg.w[lo:hi], 128), of,
                                             P[:n2].reshape(nbf, nbf, order="F"), P[n2:].reshape(nbf, nbf, order="F"))
            d_VEN[:n2] = torch.from_numpy(Va.reshape(-1, order="F").copy())
            d_VEN[n2:2 * n2] = torch.from_numpy(Vb.reshape(-1, order="F").copy())
            d_VEN[2 * n2] = E
            d_VEN[2 * n2 + 1] = ne

        sb = ShardedBuild(nbf, local_build, "cpu", nspin=2)
        V, E, ne = sb.build(np.stack([Pa, Pb]))
        assert V.shape == (2, nbf, nbf) and sb.d_VEN.numel() == 2 * n2 + 2
        if rank == 0:
            (Va, Vb), E_ref, ne_ref = orc.build_xc_u(ob, orc.Grid(cfg.xyz, cfg.w, 128), of, Pa, Pb)
            np.savez(out_path, dV=max(np.abs(V[0] - Va).max(), np.abs(V[1] - Vb).max()), dE=abs(E - E_ref), dn=abs(ne - ne_ref))
    finally:
        dist.destroy_process_group()


def test_sharded_build_unrestricted_gloo(tmp_path):
    port = 29950 + os.getpid() % 300
    out = str(tmp_path / "res_u.npz")
    mp.spawn(_worker_u, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert r["dE"] < 1e-11 and r["dV"] < 1e-11 and r["dn"] < 1e-11


def _sigma_worker(rank, world, port, out_path):
    """row f-4: ShardedSigma - every rank contracts / integrates its block range, one all-reduce of the nvec matrices."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        from serenity_b200.inputs import make_config
        from serenity_b200.inputs.configs import FUNCTIONALS
        from serenity_b200.sharded import ShardedSigma, shard_bounds
        cfg = make_config("h2o", 2)
        sub = cfg.subsystems[0]
        func = orc.Functional(*FUNCTIONALS["PBE"])
        nbf, nvec = sub.basis.nbf, 3
        nblk = (cfg.npts + 127) // 128
        bounds = shard_bounds(np.ones(nblk), world)
        lo, hi = int(bounds[rank]) * 128, min(int(bounds[rank + 1]) * 128, cfg.npts)
        ob = orc.Basis(sub.basis)

        def sigma_on(grid, D):
            rho, g, _, _ = orc.density_on_grid(ob, grid, 1e-9, sub.P, 1)
            store = orc.kernel_store_r(func, rho, g)
            return orc.kernel_integrate(ob, grid, orc.kernel_contract(ob, grid, store, D, 0, True), True)

        def local_sigma(d_D, d_F):
            mine = orc.Grid(cfg.xyz[lo:hi], cfg.w[lo:hi], 128)
            for v in range(nvec):
                D = d_D.numpy()[v * nbf * nbf:(v + 1) * nbf * nbf].reshape(nbf, nbf, order="F")
                d_F[v * nbf * nbf:(v + 1) * nbf * nbf] = torch.from_numpy(sigma_on(mine, D).reshape(-1, order="F").copy())

        ss = ShardedSigma(nbf, nvec, local_sigma, "cpu")
        rng = np.random.default_rng(3)
        D = [rng.standard_normal((nbf, nbf)) for _ in range(nvec)]
        F = ss.sigma(D)
        F2 = ss.sigma(D)  # reused buffers must not accumulate
        assert all(np.abs(a - b).max() < 1e-12 for a, b in zip(F, F2))
        if rank == 0:
            full = orc.Grid(cfg.xyz, cfg.w, 128)
            err = max(np.abs(F[v] - sigma_on(full, D[v])).max() / np.abs(F[v]).max() for v in range(nvec))
            np.savez(out_path, err=err, sym=max(np.abs(f - f.T).max() for f in F))
    finally:
        dist.destroy_process_group()


def test_sharded_sigma_allreduce_gloo(tmp_path):
    port = 29950 + os.getpid() % 300
    out = str(tmp_path / "sig.npz")
    mp.spawn(_sigma_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert r["err"] < 1e-12 and r["sym"] < 1e-13


def _nadd_worker(rank, world, port, out_path):
    """ShardedBuild with ntail = 2 + nenv: the non-additive potential of config 4 sharded over ranks (every term of
    [V_nadd | E_tot | E_act | E_env] is a sum over grid blocks)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        from serenity_b200.inputs import make_config
        from serenity_b200.inputs.configs import FUNCTIONALS
        from serenity_b200.sharded import ShardedBuild, shard_bounds
        cfg = make_config("fde_dimer", 2)
        act, env = cfg.subsystems
        func = orc.Functional(*FUNCTIONALS["PW91K"])
        nbf = act.basis.nbf
        nblk = (cfg.npts + 127) // 128
        bounds = shard_bounds(np.ones(nblk), world)
        lo, hi = int(bounds[rank]) * 128, min(int(bounds[rank + 1]) * 128, cfg.npts)
        oa, oe = orc.Basis(act.basis), orc.Basis(env.basis)

        def local_build(d_P, d_VE, p_ready=None):
            P = d_P.numpy().reshape(nbf, nbf, order="F")
            V, _, parts = orc.build_nadd(oa, P, [(oe, env.P)], orc.Grid(cfg.xyz[lo:hi], cfg.w[lo:hi], 128), func)
            d_VE[: nbf * nbf] = torch.from_numpy(V.reshape(-1, order="F").copy())
            d_VE[nbf * nbf:] = torch.from_numpy(np.asarray(parts))

        sb = ShardedBuild(nbf, local_build, "cpu", ntail=3)
        V, E_tot, E_act = sb.build(act.P)
        E_env = float(sb.h_VEN[nbf * nbf + 2])
        if rank == 0:
            V_ref, E_nadd_ref, parts = orc.build_nadd(oa, act.P, [(oe, env.P)], orc.Grid(cfg.xyz, cfg.w, 128), func)
            np.savez(out_path, dV=np.abs(V - V_ref).max(), dE=abs((E_tot - E_act - E_env) - E_nadd_ref),
                     dparts=np.abs(np.array([E_tot, E_act, E_env]) - parts).max())
    finally:
        dist.destroy_process_group()


def test_sharded_nadd_build_allreduce_gloo(tmp_path):
    port = 29300 + os.getpid() % 300
    out = str(tmp_path / "nadd.npz")
    mp.spawn(_nadd_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert r["dV"] < 1e-12 and r["dE"] < 1e-12 and r["dparts"] < 1e-12
