"""Parity at the BENCHMARK layer shape (VERDICT r01 item 5 / "what's missing" 7): every other parity test runs hidden-64/256
models, so the GEMM tiles, stream-K cuts, attention head size and KV row size of the benchmark (hidden 4096, 32 heads x 128,
MLP 11008, V = 32 859; draft hidden 768, 12 heads x 64, MLP 3072, 2 layers) were only ever checked against the repo itself.

Here a target with the LLaMA-7B LAYER shape -- 4 layers instead of 32, so the CPU oracle finishes in seconds and the host
copy of the weights stays small -- and the full LLaMA-68M-shape draft run speculative beam search on the GPU (bf16, the
production kernels) and through oracle/bssd_ref.py on the host with the SAME weights (bf16 contract of oracle/llama_ref.py).
Ranked lists must be the oracle's, or the difference must be explained at the level where the run left the oracle's
trajectory (tests/test_gpu_e2e.py::_split_level_margin); accepted lengths are reported.  K=10 strict Beauty (configs[1]) and
K=20 positional Games (configs[2]'s sizes).  The full 32-layer check is `bench.py --check-users N` (profiles/)."""
import numpy as np
import pytest
import torch

from _common import BF16_SCORE_TOL, constraint_fn, dataset, lists_match

pytestmark = pytest.mark.gpu

# absolute tolerance on cumulative log-probs of magnitude ~37 (3e-3 relative): contractions are 4096 / 11008 long here, 16-40 x
# those of the small parity models whose tolerance is BF16_SCORE_TOL = 6e-2; measured worst case 0.061 (session B)
TOL_7B = 2 * BF16_SCORE_TOL


class _GC:
    def __init__(self, num_beams):
        self.num_beams, self.do_sample, self.max_new_tokens = num_beams, False, 4


class _Handle:
    def __init__(self, dm, num_beams):
        self._atspeed_device_model = dm
        self.generation_config = _GC(num_beams)
        self.device = dm.device


def _models(ds_name):
    import bench
    from atspeed_b200.engine import DeviceModel, ModelSpec
    from oracle import llama_ref as LR
    V = dataset(ds_name).vocab_size
    dev = torch.device("cuda", 0)
    out = []
    for name, seed, layers in (("7b", 1, 4), ("68m", 2, 2)):
        s = bench.SHAPES[name]
        spec = ModelSpec(V, s["hidden"], layers, s["n_heads"], s["hidden"] // s["n_heads"], s["mlp"])
        dm = DeviceModel(spec, bench.gpu_weights(spec, seed, dev), dev)
        sh = LR.LlamaShape(V, spec.hidden, layers, spec.n_heads, spec.mlp, spec.head_dim, spec.rope_theta, spec.eps)
        ref = LR.RefLlama(sh, {"embed": dm.embed, "norm": dm.norm, "lm_head": dm.lm_head, "layers": dm.layers}, "bf16")
        out.append((dm, ref))
    return out


@pytest.mark.parametrize("ds_name,kind,K,N,gamma,users", [("beauty", "strict", 10, 40, 3, (0, 7, 1234)),
                                                          ("games", "positional", 20, 40, 3, (3, 4000))])
def test_bssd_at_the_benchmark_layer_shape_matches_the_oracle(ds_name, kind, K, N, gamma, users):
    from atspeed_b200 import beamSD
    from oracle import bssd_ref
    from test_gpu_e2e import _split_level_margin
    torch.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0))))
    (tdm, tref), (ddm, dref) = _models(ds_name)
    ds, fn = dataset(ds_name), constraint_fn(ds_name, kind)
    tm, dm = _Handle(tdm, K), _Handle(ddm, N)
    exact = steps = near = 0
    for u in users:
        prompt = ds.prompt_ids(u)
        out = beamSD.BSSD(tm, dm, {"input_ids": torch.tensor([prompt], device="cuda")}, gamma, 4, prefix_allowed_tokens_fn=fn,
                          trace=True)
        ref = bssd_ref.bssd(tref, dref, prompt, K, N, gamma, 4, fn)
        P = len(prompt)
        items, scores = out["beam_sequence"][:, P:].cpu().tolist(), out["beam_scores"].cpu().numpy()
        ok, _, msg = lists_match(items, scores, ref.sequences[:, P:].tolist(), ref.scores, TOL_7B)
        exact += int(items == ref.sequences[:, P:].tolist())
        steps += int(out["accept_steps"] == ref.accept_steps)
        if not ok:
            case = {"dataset": ds_name, "K": K}
            split = _split_level_margin(out, case, prompt, fn, model=tref, tol=TOL_7B)
            assert split["explained"], f"user {u}: {msg} | {split}"
            near += 1
        else:
            np.testing.assert_allclose(np.sort(scores)[::-1], scores, atol=0)          # sorted descending
    print(f"7B-layer-shape parity {ds_name}/{kind} K={K}: {exact}/{len(users)} ranked lists identical to the oracle, "
          f"{near} explained by a near-tie at the split level, {steps}/{len(users)} identical accepted lengths")
    beamSD.clear_sessions()
