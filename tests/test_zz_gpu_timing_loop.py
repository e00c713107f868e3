"""GPU test of the reference's per-user inference loop (code/inference.py:162-189) running on the drop-in entry points:
`run_timing` drives atspeed_b200.beamSD.BSSD / target_generate for a handful of Beauty users and the sink produces the
reference's CSV row; the search statistics in it must agree with the CPU oracle for the same users."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, oracle_model, stack_weights

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
    from oracle import bssd_ref

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
    assert (f["draft_time_cost"] + f["target_time_cost"] + f["verify_time_cost"] <= f["total_time_cost"] * 1.001).all()
    np.testing.assert_allclose(f["speedup"].astype(float), f["generalBS_time_cost"].astype(float) / f["total_time_cost"].astype(float))
    assert (f["total_accept_tokens"] == f["total_accept_steps"] * 10).all()
    # accepted steps per user against the oracle on the same weights (a bf16 near-tie may move one decision)
    agree = 0
    for i, u in enumerate(users):
        ref = bssd_ref.bssd(oracle_model("ref_bf16", "beauty", "target"), oracle_model("ref_bf16", "beauty", "correlated"),
                            ds.prompt_ids(u), 10, 40, 3, 4, fn)
        agree += int(int(f["total_accept_steps"][i]) == sum(ref.accept_steps))
    assert agree >= len(users) // 2, f"accepted steps agree for only {agree}/{len(users)} users"
    path = sink.write(str(tmp_path / csv_name("Beauty", "small-target", "small-draft", 10, 40, 0, len(users), False, 1.0, 2025)))
    import pandas as pd
    back = pd.read_csv(path)
    assert list(back.columns) == TIMING_COLUMNS and len(back) == 1 and back["beam_size"][0] == 10 and back["gamma"][0] == 3
