"""CPU test of the GEMM's host-side work decomposition (csrc/gemm.cu gemm_make_plan + the SplitMap arithmetic the consumer
kernels evaluate on the device): the persistent CTAs' unit ranges must tile the (tile, k-block) space exactly, and the
number of partial-sum slices a consumer adds for a column must equal the number of CTAs whose range touches that column's
tile -- otherwise a consumer would read an unwritten slice or drop a written one."""
import ctypes as C

import numpy as np
import pytest

SHAPES = [(4096, (4096, 4096, 4096)), (4096, (4096,)), (4096, (11008, 11008)), (11008, (4096,)), (4096, (32859,)),
          (768, (768, 768, 768)), (768, (3072, 3072)), (3072, (768,)), (768, (33014,)), (64, (64, 64, 64)), (128, (64,)),
          (256, (512, 512)), (192, (200, 72))]


@pytest.mark.parametrize("K,rows", SHAPES)
@pytest.mark.parametrize("T", [1, 10, 16, 50, 90, 128, 130, 220, 256, 289, 512])
@pytest.mark.parametrize("cut", [1, 0])
@pytest.mark.parametrize("pair", [0, 1])
def test_plan_tiles_the_work_and_slice_counts_agree(K, rows, T, cut, pair, monkeypatch):
    """pair = 1: the default -- the CTA-pair kernel serves T > 256; pair = 0: ATSPEED_GEMM_2CTA=0 keeps the single-CTA kernel."""
    from atspeed_b200 import _lib
    lib = _lib.load()
    if pair:
        monkeypatch.delenv("ATSPEED_GEMM_2CTA", raising=False)
    else:
        monkeypatch.setenv("ATSPEED_GEMM_2CTA", "0")
    r = list(rows) + [0] * (3 - len(rows))
    info = (C.c_int32 * 16)()
    cols = sum(rows)
    sl = (C.c_int32 * cols)()
    for sms in (148, 132, 7):
        rc = lib.atspeed_gemm_plan(T, K, r[0], r[1], r[2], sms, cut, info, sl)
        assert rc == 0, lib.atspeed_last_error()
        BM, KB, tiles, U, grid, max_slices, stages, tmem, bufs, T_pad = (int(info[i]) for i in range(10))
        tl = [int(info[10 + i]) for i in range(3)]
        two_cta, n_mma, N_mma = (int(info[i]) for i in (13, 14, 15))
        assert BM in (128, 256) and KB == -(-K // 64) and T_pad == -(-T // 16) * 16
        assert tiles == sum(-(-x // BM) for x in rows) == sum(tl)
        units = tiles * KB
        assert two_cta == int(pair and T_pad > 256), "the CTA-pair kernel serves exactly the forwards of more than 256 tokens"
        n_workers = grid // 2 if two_cta else grid          # CTAs, or CTA pairs, that own a unit range
        assert (n_workers - 1) * U < units <= n_workers * U, "every worker owns at least one unit and the ranges cover all units"
        assert 2 <= stages <= 12 and tmem <= 512 and tmem & (tmem - 1) == 0
        if two_cta:
            assert BM == 256 and grid % 2 == 0 and (grid <= sms or not cut)
            assert n_mma in (1, 2) and N_mma % 16 == 0 and N_mma <= 256 and n_mma * N_mma >= T > n_mma * N_mma - 32
            assert stages * (16384 + (n_mma * N_mma // 2) * 128) <= 220 * 1024
            assert bufs * n_mma * N_mma <= 512
        else:
            assert grid <= sms or not cut        # persistent: at most one CTA per SM when tiles may be cut
            assert stages * (BM * 128 + T_pad * 128) <= 220 * 1024
            assert (BM // 128) * bufs * T_pad <= 512
        if not cut:
            assert U % KB == 0 and max_slices == 1
        # slices per column == number of unit ranges [c*U, (c+1)*U) that intersect the tile's units [t*KB, (t+1)*KB)
        slices = np.frombuffer(sl, dtype=np.int32)
        col, t = 0, 0
        worst = 0
        for w in rows:
            for i in range(-(-w // BM)):
                lo, hi = t * KB, (t + 1) * KB - 1
                n = hi // U - lo // U + 1
                worst = max(worst, n)
                c0, c1 = col + i * BM, col + min(w, (i + 1) * BM)
                assert (slices[c0:c1] == n).all(), (t, n, slices[c0:c1][:4])
                t += 1
            col += w
        assert worst == max_slices


@pytest.mark.parametrize("pair", [0, 1])
def test_every_token_count_has_a_sane_plan(pair, monkeypatch):
    """All T in 1..512 for the 7B and 68M projection shapes (cohort forwards pack arbitrary token counts): unit ranges
    tile the work, rings fit in shared memory, accumulators fit in TMEM -- with the CTA-pair kernel for T > 256 (default) and
    with the single-CTA kernel everywhere (ATSPEED_GEMM_2CTA=0)."""
    from atspeed_b200 import _lib
    lib = _lib.load()
    if pair:
        monkeypatch.delenv("ATSPEED_GEMM_2CTA", raising=False)
    else:
        monkeypatch.setenv("ATSPEED_GEMM_2CTA", "0")
    shapes = {"qkv": (4096, (4096, 4096, 4096)), "o": (4096, (4096,)), "gu": (4096, (11008, 11008)), "down": (11008, (4096,)),
              "lm": (4096, (32859,)), "dqkv": (768, (768, 768, 768)), "do": (768, (768,)), "dgu": (768, (3072, 3072)),
              "ddown": (3072, (768,)), "dlm": (768, (33014,))}
    info = (C.c_int32 * 16)()
    for name, (K, rows) in shapes.items():
        r = list(rows) + [0] * (3 - len(rows))
        cut = 0 if name.endswith("lm") else 1
        for T in range(1, 513):
            assert lib.atspeed_gemm_plan(T, K, r[0], r[1], r[2], 148, cut, info, None) == 0, (name, T)
            BM, KB, tiles, U, grid, ms, stages, tmem, bufs, T_pad, _, _, _, two, n_mma, N_mma = (int(x) for x in info)
            units = tiles * KB
            assert two == int(pair and T_pad > 256)
            workers = grid // 2 if two else grid
            assert (workers - 1) * U < units <= workers * U and grid <= 148, (name, T)
            stage = 16384 + (n_mma * N_mma // 2) * 128 if two else BM * 128 + T_pad * 128
            assert 2 <= stages and stages * stage + 1024 <= 226 * 1024, (name, T)
            assert tmem <= 512 and (bufs * n_mma * N_mma <= 512 if two else (BM // 128) * bufs * T_pad <= 512), (name, T)


@pytest.mark.parametrize("K,rows", [(4096, (4096, 4096, 4096)), (4096, (4096,)), (4096, (11008, 11008)), (11008, (4096,)),
                                    (4096, (32859,)), (768, (768, 768, 768)), (768, (3072, 3072)), (3072, (768,))])
@pytest.mark.parametrize("T", [257, 289, 300, 400, 481, 512])
@pytest.mark.parametrize("cut", [1, 0])
def test_cluster_of_four_plans(K, rows, T, cut, monkeypatch):
    """ATSPEED_GEMM_CLUSTER=4 (opt-in): the worker is a cluster of two CTA pairs on adjacent tiles, the work unit a
    (super-tile, k-block).  Without a device the plan takes num_sms / 4 workers; the unit ranges must tile the super-tile x
    k-block space and a column's slice count must be the number of workers whose range touches its SUPER-tile."""
    from atspeed_b200 import _lib
    lib = _lib.load()
    monkeypatch.delenv("ATSPEED_GEMM_2CTA", raising=False)
    monkeypatch.setenv("ATSPEED_GEMM_CLUSTER", "4")
    r = list(rows) + [0] * (3 - len(rows))
    info = (C.c_int32 * 16)()
    cols = sum(rows)
    sl = (C.c_int32 * cols)()
    for sms in (148, 132):
        assert lib.atspeed_gemm_plan(T, K, r[0], r[1], r[2], sms, cut, info, sl) == 0, lib.atspeed_last_error()
        BM, KB, tiles, U, grid, max_slices, stages, tmem, bufs, T_pad = (int(info[i]) for i in range(10))
        two_cta, n_mma, N_mma = (int(info[i]) for i in (13, 14, 15))
        assert two_cta == 1 and BM == 256 and tiles == sum(-(-x // 256) for x in rows)
        if tiles < 2:
            continue                                   # a single tile keeps the pair kernel (nothing to share)
        assert grid % 4 == 0 and (grid <= sms or not cut)
        n_super = -(-tiles // 2)
        units, workers = n_super * KB, grid // 4
        assert (workers - 1) * U < units <= workers * U
        # token padding: a CTA's multicast box holds N_mma / 4 tokens, a whole number of 8-row swizzle atoms
        assert N_mma % 32 == 0 and n_mma * N_mma >= T > n_mma * N_mma - 64
        assert stages * (16384 + (n_mma * N_mma // 2) * 128) <= 220 * 1024 and bufs * n_mma * N_mma <= 512
        slices = np.frombuffer(sl, dtype=np.int32)
        col, t = 0, 0
        for w in rows:
            for i in range(-(-w // 256)):
                st = t // 2
                lo, hi = st * KB, (st + 1) * KB - 1
                expect = hi // U - lo // U + 1
                assert (slices[col: col + min(256, w - i * 256)] == expect).all(), (K, rows, T, sms, t)
                col += min(256, w - i * 256)
                t += 1
        assert max_slices == max(int(slices.max()), 1) or not cut
