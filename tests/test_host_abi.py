"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares,
fails loudly without a GPU (no CPU fallback), and the host-only helpers work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "serenity_xc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sxc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from serenity_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libserenity_xc_b200.so lacks %s" % name
    assert sorted(_lib.SYMBOLS) == declared  # the ctypes binding covers the whole header
    assert lib.sxc_abi_version() == 5


def test_stats_struct_matches_header():
    """ctypes mirror of sxc_stats has the layout of the C struct (10 int64, 2 int32, 8 float, 8 int32, 1 float)."""
    from serenity_b200._lib import Stats
    assert C.sizeof(Stats) == 10 * 8 + 2 * 4 + 8 * 4 + 8 * 4 + 4 + 4  # + tail padding to 8


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from serenity_b200.xc import XCContext
    from serenity_b200._lib import SerenityError
    with pytest.raises(SerenityError):
        XCContext(0)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "serenity_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "sharded.py" and "import oracle" not in src and "from oracle" not in src, \
                    os.path.join(dirpath, f)


def test_balance_ranges_contiguous_and_balanced():
    from serenity_b200.sharded import shard_bounds
    rng = np.random.default_rng(3)
    cost = rng.uniform(1.0, 100.0, size=1000)
    for world in (1, 2, 3, 4, 8):
        b = shard_bounds(cost, world)
        assert b[0] == 0 and b[-1] == len(cost) and np.all(np.diff(b) >= 0)
        sums = [cost[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(sums) - min(sums) <= 2 * cost.max() + 1e-9
    # degenerate inputs: fewer blocks than ranks, empty list
    b = shard_bounds(np.ones(3), 8)
    assert b[0] == 0 and b[-1] == 3 and np.all(np.diff(b) >= 0) and np.diff(b).sum() == 3
    b = shard_bounds(np.zeros(0), 4)
    assert list(b) == [0, 0, 0, 0, 0]


def test_scatter_schedule_covers_the_upper_triangle_once():
    """Host logic of k_vmat (sxc_api.cu: build_scatter_schedule): for every block size the rounds hold each upper-triangle
    32 x 32 warp tile exactly once (or twice with complementary k-step masks when a round has <= 4 tiles), stage at most 6
    distinct row groups, give every warp at most one tile, and refer only to staged groups."""
    from serenity_b200 import _lib
    lib = _lib.load()
    for n32 in list(range(1, 40)) + [64, 100, 255]:
        nr = lib.sxc_debug_scatter_schedule(n32, None, 0)
        assert nr > 0
        buf = np.zeros((nr, 40), dtype=np.uint8)
        assert lib.sxc_debug_scatter_schedule(n32, buf.ctypes.data_as(C.c_void_p), nr) == nr
        seen = {}
        for r in buf:
            ng, group, ta, tb, km = int(r[0]), r[8:16], r[16:24], r[24:32], r[32:40]
            assert 1 <= ng <= 6 and len(set(group[:ng].tolist())) == ng and all(g < n32 for g in group[:ng])
            for w in range(8):
                if ta[w] == 0xFF:
                    assert tb[w] == 0xFF
                    continue
                assert ta[w] < ng and tb[w] < ng and km[w] in (1, 2, 3)
                i, j = int(group[ta[w]]), int(group[tb[w]])
                assert i <= j
                seen[(i, j)] = seen.get((i, j), 0) | int(km[w])
                if km[w] != 3:  # a split tile: the other k-step must be in the same round
                    partner = [v for v in range(8) if v != w and ta[v] == ta[w] and tb[v] == tb[w]]
                    assert len(partner) == 1 and km[partner[0]] == 3 - km[w]
        assert sorted(seen) == [(i, j) for i in range(n32) for j in range(i, n32)]
        assert all(v == 3 for v in seen.values())
        assert nr <= (n32 * (n32 + 1) // 2 + 7) // 8 + n32  # never far from the slot bound


def test_scatter_schedule_v2_partitions_every_tile_over_warps_and_k_steps():
    """Host logic of k_vmat_tma (sxc_api.cu: build_scatter_schedule2): every upper-triangle warp tile sits in exactly one round;
    inside it the k-steps of a chunk are partitioned (no gap, no overlap) over the warps that share the tile; a round stages at
    most 6 row groups and uses at most 8 warps; the per-warp group indices repeat the staged slots; the schedule is never
    longer (in DMMA time units: 16 per full tile and k-step share, 10 per diagonal tile) than the v1 schedule."""
    from serenity_b200 import _lib
    lib = _lib.load()

    def units_v1(n32):
        nr = lib.sxc_debug_scatter_schedule(n32, None, 0)
        buf = np.zeros((nr, 40), dtype=np.uint8)
        lib.sxc_debug_scatter_schedule(n32, buf.ctypes.data_as(C.c_void_p), nr)
        tot = 0.0
        for r in buf:
            ta, tb, km = r[16:24], r[24:32], r[32:40]
            tot += max((10 if ta[w] == tb[w] else 16) * bin(int(km[w])).count("1") / 2 for w in range(8) if ta[w] != 0xFF)
        return tot

    for ks in (2, 4):
        full = (1 << ks) - 1
        for n32 in list(range(1, 40)) + [64, 100, 255]:
            nr = lib.sxc_debug_scatter_schedule2(n32, ks, None, 0)
            assert nr > 0
            buf = np.zeros((nr, 64), dtype=np.uint8)
            assert lib.sxc_debug_scatter_schedule2(n32, ks, buf.ctypes.data_as(C.c_void_p), nr) == nr
            seen, units = {}, 0.0
            for ri, r in enumerate(buf):
                ng, group, ta, tb, km, ga, gb = int(r[0]), r[8:16], r[16:24], r[24:32], r[32:40], r[40:48], r[48:56]
                assert 1 <= ng <= 6 and len(set(group[:ng].tolist())) == ng and all(g < n32 for g in group[:ng])
                in_round, cost = {}, 0.0
                for w in range(8):
                    if ta[w] == 0xFF:
                        assert tb[w] == 0xFF
                        continue
                    assert ta[w] < ng and tb[w] < ng and 0 < km[w] <= full
                    i, j = int(group[ta[w]]), int(group[tb[w]])
                    assert i <= j and (int(ga[w]), int(gb[w])) == (i, j)
                    assert in_round.get((i, j), 0) & int(km[w]) == 0      # k-steps of a tile are not done twice
                    in_round[(i, j)] = in_round.get((i, j), 0) | int(km[w])
                    cost = max(cost, (10 if i == j else 16) * bin(int(km[w])).count("1") / ks)
                assert all(v == full for v in in_round.values())         # ... and none is left out
                for t in in_round:
                    assert t not in seen, "tile in two rounds"
                    seen[t] = ri
                units += cost
            assert sorted(seen) == [(i, j) for i in range(n32) for j in range(i, n32)]
            assert units <= units_v1(n32) + 1e-9, (ks, n32, units, units_v1(n32))
