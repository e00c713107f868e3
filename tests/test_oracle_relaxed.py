"""AtSpeed-R (sampling / relaxed acceptance) on the CPU:
  * the oracle's sampling branch reproduces the UNMODIFIED reference draw for draw under the same torch seed
    (tests/golden/bssd_relaxed.json, written by tools/make_golden_relaxed.py);
  * the exponential-race identity the CUDA path relies on: torch.multinomial(p, n) without replacement ==
    top-n of p / Exp(1) noise drawn from the same generator state;
  * the counter-based generator of the CUDA path (csrc/noise.cuh, evaluated on the host through the C ABI) is
    Philox4x32-10: published known-answer vector + an independent numpy implementation;
  * noise mode of the oracle: statistically sane acceptance, deterministic, defined fallback."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, golden, oracle_model
from oracle import bssd_ref


def _cases():
    return golden("bssd_relaxed.json")["cases"]


def _id(c):
    return (f"{c['dataset']}-{c['constraint']}-{c['draft']}-K{c['K']}N{c['N']}g{c['gamma']}T{c['temperature']}"
            f"u{c['user']}s{c['seed']}")


@pytest.mark.parametrize("case", _cases(), ids=_id)
def test_oracle_relaxed_reproduces_reference(case):
    ds = dataset(case["dataset"])
    fn = constraint_fn(case["dataset"], case["constraint"])
    prompt = ds.prompt_ids(case["user"])
    tgt = oracle_model("ref_bf16", case["dataset"], "target")
    dft = oracle_model("ref_bf16", case["dataset"], case["draft"])
    cfg = bssd_ref.SamplingCfg(temperature=case["temperature"], top_k=case["top_k"])
    torch.manual_seed(case["seed"])
    if case["raises"]:
        with pytest.raises((RuntimeError, ValueError, IndexError)):
            bssd_ref.bssd(tgt, dft, prompt, case["K"], case["N"], case["gamma"], 4, fn, sampling=cfg)
        return
    res = bssd_ref.bssd(tgt, dft, prompt, case["K"], case["N"], case["gamma"], 4, fn, sampling=cfg)
    P = len(prompt)
    assert res.sequences[:, P:].tolist() == case["bssd"]["items"]
    # scores: the bf16-contract model rounds activations per batch composition (the reference batches the verify tree
    # differently from the oracle), which moves an occasional score by a few 1e-3; items and decisions are exact
    np.testing.assert_allclose(res.scores, case["bssd"]["scores"], rtol=0, atol=1e-2)
    assert res.accept_steps == [r["n_matches"] for r in case["rounds"]]
    assert res.n_run == case["n_run"]
    st = res.stats(case["K"])
    assert st["total_accept_steps"] == case["total_accept_steps"]
    assert abs(st["ave_accept_tokens"] - case["ave_accept_tokens"]) < 1e-9
    for tr, g in zip(res.rounds, case["rounds"]):
        for lv, gt, gp in zip(tr.draft_levels, g["draft_tokens"], g["draft_parents"]):
            # G4: with fewer than N positive-probability candidates the reference's multinomial pads
            # with zero-probability picks and its token-range filter keeps those landing on id 2 / >= 32000; the
            # restatement drops them, so ours is a prefix of theirs, shorter only when the level is not full.
            n = len(lv)
            assert [t for _, t, _ in lv] == gt[:n] and [p for p, _, _ in lv] == gp[:n]
            assert n == len(gt) or n < case["N"]


def test_relaxed_golden_has_accepting_and_rejecting_cases():
    ok = [c for c in _cases() if not c["raises"]]
    hist = np.bincount([c["total_accept_steps"] for c in ok])
    assert len(ok) >= 40 and hist[0] > 0 and hist[1:].sum() > 0, hist


@pytest.mark.parametrize("n,k", [(50, 10), (2000, 40), (32859, 40)])
def test_multinomial_is_an_exponential_race(n, k):
    g = torch.Generator().manual_seed(123 + n)
    p = torch.rand(n, generator=g) ** 4
    p[::7] = 0
    p = p / p.sum()
    g1 = torch.Generator().manual_seed(99)
    a = torch.multinomial(p, k, replacement=False, generator=g1)
    g2 = torch.Generator().manual_seed(99)
    noise = torch.empty_like(p).exponential_(1, generator=g2)
    b = torch.topk(p / noise, k)[1]
    assert a.tolist() == b.tolist()


def _philox_np(seed, stream, idx):
    """Independent Philox4x32-10 (Salmon et al. 2011), word 0; counter = (idx, 0, stream_lo, stream_hi)."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    idx = np.asarray(idx, dtype=np.uint64)
    c = [idx & np.uint64(0xFFFFFFFF), np.zeros_like(idx), np.full_like(idx, stream & 0xFFFFFFFF), np.full_like(idx, stream >> 32)]
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c[0].astype(np.uint32)


def test_noise_generator_is_philox4x32_10():
    from atspeed_b200 import _lib
    lib = _lib.load()
    assert lib.atspeed_noise_host_u32(0, 0, 0) == 0x6627E8D5          # Random123 known-answer vector (all-zero input)
    rng = np.random.default_rng(0)
    for _ in range(20):
        seed, stream = int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 63))
        idx = rng.integers(0, 2 ** 32, size=64, dtype=np.uint64)
        want = _philox_np(seed, stream, idx)
        got = np.array([lib.atspeed_noise_host_u32(seed, stream, int(i)) for i in idx], dtype=np.uint32)
        assert (got == want).all()
    # stream ids keep user / round / level / site apart
    ids = {lib.atspeed_noise_stream(u, r, l, s) for u in range(3) for r in range(4) for l in range(5) for s in range(6)}
    assert len(ids) == 3 * 4 * 5 * 6
    # uniform transform: 23 bits + half an ulp is in (0,1) and unbiased enough
    bits = _philox_np(12345, 678, np.arange(200000))
    u = ((bits >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 3e-3 and abs(np.var(u) - 1 / 12) < 2e-3
    e = -np.log(u)
    assert abs(e.mean() - 1) < 1e-2


def _numpy_noise_fn(seed, user_seq=0):
    def fn(kind, site, rnd, level, n):
        stream = (user_seq << 16) | ((rnd & 0xFF) << 8) | ((level & 0xF) << 4) | (site & 0xF)
        bits = _philox_np(seed, stream, np.arange(n))
        if kind == "bits":
            return torch.from_numpy(bits.astype(np.int64))
        u = ((bits >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
        return torch.from_numpy(u if kind == "uniform" else (-np.log(u)).astype(np.float32))
    return fn


@pytest.mark.parametrize("draft,constraint", [("correlated", "positional"), ("correlated", "strict"), ("independent", "strict")])
def test_noise_mode_is_deterministic_and_well_formed(draft, constraint):
    ds = dataset("beauty")
    fn = constraint_fn("beauty", constraint)
    tgt, dft = oracle_model("ref_bf16", "beauty", "target"), oracle_model("ref_bf16", "beauty", draft)
    prompt = ds.prompt_ids(3)
    runs = []
    for _ in range(2):
        cfg = bssd_ref.SamplingCfg(temperature=1.0, top_k=50, noise_fn=_numpy_noise_fn(42), defined_fallback=True)
        runs.append(bssd_ref.bssd(tgt, dft, prompt, 10, 40, 3, 4, fn, sampling=cfg))
    a, b = runs
    assert a.sequences.tolist() == b.sequences.tolist() and a.accept_steps == b.accept_steps
    P = len(prompt)
    items = [tuple(r) for r in a.sequences[:, P:].tolist()]
    assert 1 <= len(items) <= 10 and len(set(items)) == len(items)          # beams are distinct sequences
    assert all(a.scores[i] >= a.scores[i + 1] for i in range(len(items) - 1))   # final sort (beamSD.py:529-531)
    other = bssd_ref.bssd(tgt, dft, prompt, 10, 40, 3, 4, fn,
                          sampling=bssd_ref.SamplingCfg(noise_fn=_numpy_noise_fn(43), defined_fallback=True))
    assert other.sequences.tolist() != a.sequences.tolist()                 # a different key gives different samples
