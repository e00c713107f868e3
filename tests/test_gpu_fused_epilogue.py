"""Fused GEMM epilogues (csrc/gemm.cu FusedEpi): the q|k|v GEMM applies RoPE and appends to the KV cache, the gate|up GEMM
applies SiLU * up, tiles cut across CTAs are completed inside the kernel (owner CTA adds the other CTAs' raw partials in
slice order).  Same rounding points as the row-wise kernels they replace (elementwise.cu qkv_rope_append / silu_mul), so a
forward with ATSPEED_FUSED_EPI=0 (slices + consumer kernels) must give the same logits up to the fp32 summation order of
the differently cut k-ranges -- checked here on every tile path: stacked 256-row tiles (T <= 128), 128-row tiles
(128 < T <= 256), the CTA pair (T > 256) and, with ATSPEED_GEMM_2CTA=0, single-CTA tiles at T > 256; at the small parity
shape (heads of 64) and at the benchmark layer shape (heads of 128, hidden 4096, MLP 11008).  Also against the oracle."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, oracle_model, stack_weights

pytestmark = pytest.mark.gpu


def _session(dm, T_cap):
    from atspeed_b200.constraint import compile_constraint
    from atspeed_b200.engine import DeviceTrie, Session
    ds = dataset("beauty")
    csr = compile_constraint(constraint_fn("beauty", "positional"), ds.prompt_ids(0), 4)
    return Session(dm, None, DeviceTrie(csr, dm.device), K=10, N=10, max_new_tokens=4, max_prompt=T_cap)


def _causal_forward(sess, toks, rows):
    T = len(toks)
    i32 = lambda x: torch.tensor(list(x), dtype=torch.int32, device="cuda")
    vis = torch.zeros(T, 16, dtype=torch.int32, device="cuda")
    return sess.forward_raw(0, i32(toks), i32(range(T)), i32(range(T)), i32(range(1, T + 1)), vis, T, T, i32(rows))


def _toks(T, V, seed):
    return np.random.default_rng(seed).integers(3, V, T).tolist()


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("T", [7, 100, 128, 130, 250, 300, 470])
def test_fused_forward_equals_rowwise_forward_small_shape(T, pair, monkeypatch):
    from atspeed_b200.engine import DeviceModel, ModelSpec
    monkeypatch.setenv("ATSPEED_GEMM_2CTA", pair)
    sh, W = stack_weights("ref_bf16", "beauty", "target")
    spec = ModelSpec(sh.vocab, sh.hidden, sh.n_layers, sh.n_heads, sh.head_dim, sh.mlp, sh.eps, sh.rope_theta)
    dm = DeviceModel(spec, W, "cuda")
    toks = _toks(T, sh.vocab, T)
    rows = sorted({0, T // 2, T - 1})
    monkeypatch.setenv("ATSPEED_FUSED_EPI", "1")
    fused = _causal_forward(_session(dm, 500), toks, rows)
    monkeypatch.setenv("ATSPEED_FUSED_EPI", "0")
    plain = _causal_forward(_session(dm, 500), toks, rows)
    err = float(np.abs(fused - plain).max())
    print(f"T={T} pair={pair}: max |fused - rowwise| logit {err:.5f} (logit std {plain.std():.3f})")
    assert err < 2e-2, err
    if T <= 130:
        ref = oracle_model("ref_bf16", "beauty", "target")
        want = ref.forward(torch.tensor(toks), torch.arange(T), torch.tril(torch.ones(T, T, dtype=torch.bool)), ref.new_cache(),
                           torch.tensor(rows)).numpy()
        assert float(np.abs(fused - want).max()) < 6e-2


@pytest.mark.parametrize("T", [40, 200, 420])
def test_fused_forward_equals_rowwise_forward_benchmark_layer_shape(T, monkeypatch):
    import bench
    from atspeed_b200.engine import DeviceModel, ModelSpec
    V = dataset("beauty").vocab_size
    s = bench.SHAPES["7b"]
    spec = ModelSpec(V, s["hidden"], 2, s["n_heads"], s["hidden"] // s["n_heads"], s["mlp"])
    dm = DeviceModel(spec, bench.gpu_weights(spec, 5, torch.device("cuda", 0)), "cuda")
    toks = _toks(T, V, 100 + T)
    rows = sorted({0, T // 3, T - 1})
    monkeypatch.setenv("ATSPEED_FUSED_EPI", "1")
    fused = _causal_forward(_session(dm, 500), toks, rows)
    monkeypatch.setenv("ATSPEED_FUSED_EPI", "0")
    plain = _causal_forward(_session(dm, 500), toks, rows)
    err = float(np.abs(fused - plain).max())
    print(f"7B layer shape, T={T}: max |fused - rowwise| logit {err:.5f} (logit std {plain.std():.3f})")
    assert err < 3e-2, err
