"""CPU test of the cohort scheduler's packing policy (csrc/cohort.cu plan_packs through atspeed_debug_plan_packs): which ready
users share a target forward, and which packs are held back a scheduler step for a better fill.  Pure host arithmetic -- the
GPU tiers (tests/test_gpu_cohort*.py) check that who shares a forward never changes a user's result."""
import ctypes as C

import numpy as np
import pytest

T_MAX, R_MAX, MAX_USERS = 512, 512, 16


def plan(T, R=None, waited=None, defer=1, no_more_work=0, T_max=T_MAX, R_max=R_MAX):
    from atspeed_b200 import _lib
    lib = _lib.load()
    n = len(T)
    R = list(R) if R is not None else list(T)
    waited = list(waited) if waited is not None else [0] * n
    a = lambda v: (C.c_int32 * max(1, n))(*v)
    pack_of = (C.c_int32 * max(1, n))()
    run_now = (C.c_uint8 * max(1, n))()
    n_packs = C.c_int32(0)
    rc = lib.atspeed_debug_plan_packs(a(T), a(R), a(waited), n, T_max, R_max, defer, no_more_work, pack_of, run_now, C.byref(n_packs))
    assert rc == 0, lib.atspeed_last_error()
    return [int(pack_of[i]) for i in range(n)], [int(run_now[b]) for b in range(n_packs.value)]


def check_invariants(T, R, pack_of, run_now, T_max=T_MAX, R_max=R_MAX):
    n_packs = len(run_now)
    assert sorted(set(pack_of)) == list(range(n_packs)), "every pack holds an item and every item sits in exactly one pack"
    for b in range(n_packs):
        members = [i for i, p in enumerate(pack_of) if p == b]
        assert sum(T[i] for i in members) <= T_max and sum(R[i] for i in members) <= R_max and len(members) <= MAX_USERS
    assert n_packs == 0 or any(run_now), "progress: some pack always runs"


def test_first_round_and_later_round_trees_are_packed_best_fit_decreasing():
    # one scheduler step at the benchmark configuration (K=10, N=40, gamma=3): first rounds are prompt + 120 tokens, later
    # rounds 90 / 50 / 10
    T = [300, 330, 90, 50, 10, 90, 260, 50]
    pack_of, run_now = plan(T)
    check_invariants(T, T, pack_of, run_now)
    packs = {b: sorted((T[i] for i, p in enumerate(pack_of) if p == b), reverse=True) for b in set(pack_of)}
    assert packs[0] == [330, 90, 90], "opened by the largest item, filled with the largest that fit"
    assert packs[1] == [300, 50, 50, 10]
    assert packs[2] == [260]
    assert run_now == [1, 0, 0], "510 tokens run; 410 and 260 (< 7/8 of 512) wait a step for the small trees to come back"
    # nothing else can arrive: waiting cannot help
    assert plan(T, no_more_work=1)[1] == [1, 1, 1]
    # deferral off (ATSPEED_COHORT_DEFER=0): everything runs
    assert plan(T, defer=0)[1] == [1, 1, 1]
    # the user with 260 tokens has already waited twice: its pack runs
    assert plan(T, waited=[0, 0, 0, 0, 0, 0, 2, 0])[1] == [1, 0, 1]


def test_progress_when_no_pack_is_full():
    pack_of, run_now = plan([200, 150, 90])
    assert pack_of == [0, 0, 0] and run_now == [1], "440 < 448 tokens, but holding back the only pack would stall the step"
    pack_of, run_now = plan([300, 290])
    assert pack_of == [0, 1] and run_now == [1, 0], "the fullest pack runs"
    assert plan([]) == ([], [])


def test_row_and_user_limits():
    T = [20] * 40
    pack_of, run_now = plan(T)
    check_invariants(T, T, pack_of, run_now)
    assert max(np.bincount(pack_of)) == MAX_USERS and len(run_now) == 3
    T, R = [100, 100, 100, 100], [300, 300, 100, 100]
    pack_of, run_now = plan(T, R)
    check_invariants(T, R, pack_of, run_now)
    assert pack_of[0] != pack_of[1], "two users whose logit rows do not fit together get separate forwards"


@pytest.mark.parametrize("seed", range(20))
def test_random_steps_keep_the_invariants(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 17))
    first = rng.random(n) < 0.4
    T = [int(rng.integers(186, 290)) if f else int(rng.choice([90, 50, 10])) for f in first]
    waited = [int(w) for w in rng.integers(0, 3, n)]
    for defer in (0, 1):
        for nmw in (0, 1):
            pack_of, run_now = plan(T, waited=waited, defer=defer, no_more_work=nmw)
            check_invariants(T, T, pack_of, run_now)
            for b, r in enumerate(run_now):
                members = [i for i, p in enumerate(pack_of) if p == b]
                fill = sum(T[i] for i in members)
                must = (not defer) or nmw or fill >= T_MAX - T_MAX // 8 or any(waited[i] >= 2 for i in members)
                assert r == 1 if must else True
                if r == 0:
                    assert not must
            # packs are opened in order of their largest item
            heads = [max(T[i] for i, p in enumerate(pack_of) if p == b) for b in range(len(run_now))]
            assert heads == sorted(heads, reverse=True)
