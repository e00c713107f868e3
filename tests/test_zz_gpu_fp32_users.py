"""GPU fp32 parity on 64 + 64 more users (tests/golden/bssd_strict_users.json, written by the unmodified reference through
tools/make_golden_users.py: HF fp32 tiny target + correlated draft, K=10, N=40, gamma=3, strict / positional alternating).
Same path as tests/test_gpu_fp32_parity.py (fp32 forward + kernels (a)/(b)/(c) + beam-tree runtime, through
atspeed_b200.beamSD).  A ranked list must equal the reference's; the only admissible exception is a pair of beams whose fp32
scores coincide to 1e-4 absolute (summation order of the fp32 GEMMs), and at least 95 % of the lists must be identical."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, golden, lists_match, stack_weights

pytestmark = pytest.mark.gpu


class _GC:
    def __init__(self, num_beams):
        self.num_beams, self.do_sample, self.max_new_tokens = num_beams, False, 4


class _Handle:
    def __init__(self, dm, num_beams):
        self._atspeed_device_model = dm
        self.generation_config = _GC(num_beams)
        self.device = dm.device


@pytest.mark.parametrize("ds_name", ["beauty", "games"])
def test_fp32_bssd_matches_reference_on_64_users(ds_name):
    from atspeed_b200 import beamSD
    from atspeed_b200.engine import DeviceModel, ModelSpec
    cases = [c for c in golden("bssd_strict_users.json")["cases"] if c["dataset"] == ds_name]
    assert len(cases) == 64
    models = {}
    for which in ("target", "correlated"):
        sh, W = stack_weights("hf_fp32", ds_name, which)
        spec = ModelSpec(sh.vocab, sh.hidden, sh.n_layers, sh.n_heads, sh.head_dim, sh.mlp, sh.eps, sh.rope_theta)
        models[which] = DeviceModel(spec, W, "cuda", dtype=torch.float32)
    tm, dm = _Handle(models["target"], 10), _Handle(models["correlated"], 40)
    ds = dataset(ds_name)
    exact = same_steps = 0
    worst_rel = 0.0
    for case in cases:
        prompt = ds.prompt_ids(case["user"])
        fn = constraint_fn(ds_name, case["constraint"])
        out = beamSD.BSSD(tm, dm, {"input_ids": torch.tensor([prompt], device="cuda")}, case["gamma"], 4,
                          prefix_allowed_tokens_fn=fn)
        P = len(prompt)
        items = out["beam_sequence"][:, P:].cpu().tolist()
        scores = out["beam_scores"].cpu().numpy()
        if items == case["bssd"]["items"]:
            exact += 1
            want = np.asarray(case["bssd"]["scores"], dtype=np.float64)
            worst_rel = max(worst_rel, float(np.max(np.abs(scores - want) / np.abs(want))))
        else:
            ok, _, msg = lists_match(items, scores, case["bssd"]["items"], case["bssd"]["scores"], 1e-4)
            assert ok, f"user {case['user']}: {msg}"
        same_steps += int(out["accept_steps"] == case["accept_steps"] and out["n_run"] == case["n_run"])
    print(f"{ds_name}: {exact}/64 ranked lists identical, {same_steps}/64 accepted-length sequences identical, "
          f"max relative score error {worst_rel:.2e}")
    assert worst_rel < 1e-3
    assert exact >= 61 and same_steps >= 61, (exact, same_steps)
