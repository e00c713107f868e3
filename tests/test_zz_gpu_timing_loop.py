"""GPU test of the reference's per-user inference loop (code/inference.py:162-189) running on the drop-in entry points:
`run_timing` drives atspeed_b200.beamSD.BSSD / target_generate for a handful of Beauty users and the sink produces the
reference's CSV row with consistent timings and search statistics."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, stack_weights

pytestmark = pytest.mark.gpu


class _GC:
    def __init__(self, num_beams):
        self.num_beams, self.do_sample, self.max_new_tokens = num_beams, False, 4


class _Handle:
    def __init__(self, dm, num_beams):
        self._atspeed_device_model = dm
        self.generation_config = _GC(num_beams)
        self.device = dm.device


def test_timing_loop_on_the_drop_in_entry_points(tmp_path):
    from atspeed_b200 import beamSD
    from atspeed_b200.engine import DeviceModel, ModelSpec
    from atspeed_b200.timing import TIMING_COLUMNS, csv_name, run_timing

    ds = dataset("beauty")
    fn = constraint_fn("beauty", "strict")
    models = {}
    for which in ("target", "correlated"):
        sh, W = stack_weights("ref_bf16", "beauty", which)
        spec = ModelSpec(sh.vocab, sh.hidden, sh.n_layers, sh.n_heads, sh.head_dim, sh.mlp, sh.eps, sh.rope_theta)
        models[which] = DeviceModel(spec, W, "cuda")
    tm, dm = _Handle(models["target"], 10), _Handle(models["correlated"], 40)
    users = [2, 5, 9, 11]
    batches = [{"input_ids": torch.tensor([ds.prompt_ids(u)], device="cuda")} for u in users]
    sink = run_timing(beamSD.BSSD, beamSD.target_generate, None, batches, tm, dm, 3, 4, fn, "small-target", "small-draft")
    assert len(sink) == len(users)
    f = sink.frame
    assert list(f.columns) == TIMING_COLUMNS
    assert (f["total_time_cost"] > 0).all() and (f["generalBS_time_cost"] > 0).all()
    assert (f["draft_time_cost"] + f["target_time_cost"] + f["verify_time_cost"] <= f["total_time_cost"] * 1.01 + 1e-4).all()
    np.testing.assert_allclose(f["speedup"].astype(float), f["generalBS_time_cost"].astype(float) / f["total_time_cost"].astype(float))
    assert (f["total_accept_tokens"] == f["total_accept_steps"] * 10).all()
    # search statistics are those of BSSD (code/beamSD.py:532-542): 4 new tokens, gamma = 3 -> 1..3 verify rounds, at most 3
    # accepted draft steps; parity of the decisions themselves is the business of tests/test_gpu_e2e.py
    steps, toks, ave = (f[c].astype(float) for c in ("total_accept_steps", "total_accept_tokens", "ave_accept_tokens"))
    assert ((steps >= 0) & (steps <= 3)).all() and (toks == steps * 10).all()
    assert ((ave >= 0) & (ave <= 30)).all()
    path = sink.write(str(tmp_path / csv_name("Beauty", "small-target", "small-draft", 10, 40, 0, len(users), False, 1.0, 2025)))
    import pandas as pd
    back = pd.read_csv(path)
    assert list(back.columns) == TIMING_COLUMNS and len(back) == 1 and back["beam_size"][0] == 10 and back["gamma"][0] == 3
