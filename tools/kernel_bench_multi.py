#!/usr/bin/env python
"""Row f-4 on N GPUs: the LR-TDDFT kernel sigma build sharded over the ranks of one node (serenity_b200/sharded.py:
ShardedSigma - grid blocks per rank, ONE NCCL all-reduce of the nvec Fock-like matrices).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
      tools/kernel_bench_multi.py [workload] [nvec] [steps]
Rank 0 prints one JSON line: device time per sigma build (CUDA events, max over ranks, barrier on both sides), the same
through pinned host buffers, and the deviation from the unsharded single-GPU result computed on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from serenity_b200.inputs import make_config  # noqa: E402
from serenity_b200.inputs.configs import FUNCTIONALS  # noqa: E402
from serenity_b200.sharded import ShardedSigma, cuda_local_sigma  # noqa: E402
from serenity_b200.xc import XCContext  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tetracene"
    nvec = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    real = os.dup(1)  # keep stdout for the JSON line (NCCL prints its banner to fd 1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_config(name)
    sub = cfg.subsystems[0]
    nb = sub.basis.nbf
    ids, mix = FUNCTIONALS[cfg.functional]

    def setup(shard):
        ctx = XCContext(local)
        g = ctx.set_grid(cfg.xyz, cfg.w, 128)
        if shard and world > 1:
            ctx.set_grid_shard(g, rank, world)
        b = ctx.add_basis(sub.basis, 1e-9)
        k = ctx.kernel_create(g, 1, True)
        ctx.kernel_add(k, ctx.set_functional(ids, mix), [b], [sub.P])
        return ctx, g, b, k

    ctx, g, b, k = setup(True)
    ss = ShardedSigma(nb, nvec, cuda_local_sigma(ctx, g, b, [k], nvec), dev)
    rng = np.random.default_rng(0)
    D = [rng.standard_normal((nb, nb)) * 1e-2 for _ in range(nvec)]
    F = ss.sigma(D)  # warm-up: plans, workspace, NCCL communicator

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev = timed(ss.sigma_device)
    ms_e2e = timed(lambda: ss.sigma(D))
    line = None
    if rank == 0:
        line = {"workload": cfg.description, "name": cfg.name, "functional": cfg.functional, "grid_points": cfg.npts, "nbf": nb,
                "nvec": nvec, "n_gpus": world, "steps": steps, "sigma_device_ms": ms_dev, "sigma_e2e_ms": ms_e2e,
                "grid_pts_x_vectors_per_s": cfg.npts * nvec / (ms_dev * 1e-3), "allreduce_bytes": nvec * nb * nb * 8,
                "h2d_bytes_per_build": nvec * nb * nb * 8 * world, "d2h_bytes_per_build": nvec * nb * nb * 8 * world}
        if world > 1:  # the unsharded build on this GPU as the checker
            c1, g1, b1, k1 = setup(False)
            F1 = c1.kernel_sigma(g1, b1, nb, [k1], D, 0)
            line["max_rel_dev_vs_unsharded"] = float(max(np.abs(a - c).max() / np.abs(c).max() for a, c in zip(F, F1)))
            c1.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if line is not None:
        os.write(real, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
