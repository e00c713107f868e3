"""CPU tests of the result / timing sink: the CSV schema, derived columns and file name of the reference's
inference loop (/root/reference/code/inference.py:152-156, :179-189), driven through `run_timing` with stub searches."""
import types

import numpy as np
import pandas as pd

from atspeed_b200.timing import GROUP_KEYS, TIMING_COLUMNS, TimingSink, csv_name, run_timing, timing_row


def test_columns_are_the_reference_schema():
    assert TIMING_COLUMNS == ["target_model", "draft_model", "beam_size", "gamma", "draft_time_cost", "target_time_cost",
                              "verify_time_cost", "total_time_cost", "generalBS_time_cost", "TF_cache_time_cost", "speedup",
                              "speedupTF", "total_accept_steps", "total_accept_tokens", "ave_accept_tokens", "overhead"]
    assert GROUP_KEYS == TIMING_COLUMNS[:4]


def test_row_derived_quantities():
    out = {"time_cost": 0.020, "draft_time_cost": 0.004, "target_time_cost": 0.012, "verify_time_cost": 0.001, "n_run": 2,
           "total_accept_steps": 3, "total_accept_tokens": 30, "ave_accept_tokens": 15.0}
    tgt = {"time_cost": 0.050}
    row = timing_row("7b", "68m", 10, 3, out, tgt, 0.080, 4)
    d = dict(zip(TIMING_COLUMNS, row))
    assert d["speedup"] == 0.050 / 0.020 and d["speedupTF"] == 0.080 / 0.020
    assert d["overhead"] == (0.020 * 4) / (0.050 * 2)                       # code/inference.py:181
    assert d["total_time_cost"] == 0.020 and d["generalBS_time_cost"] == 0.050 and d["TF_cache_time_cost"] == 0.080
    assert (d["total_accept_steps"], d["total_accept_tokens"], d["ave_accept_tokens"]) == (3, 30, 15.0)


def test_csv_name_matches_the_reference_pattern():
    assert csv_name("Beauty", "llama-7b", "llama-68m", 10, 40, 0, 100, False, 1.0, 2025) == \
        "AnaResult/Beauty/timing_mean_llama-7b_llama-68m_B10-40_0-100__temp1.0_seed2025.csv"
    assert csv_name("Games", "t", "d", 20, 40, 5, 9, True, 0.7, 1).endswith("_B20-40_5-9_do_sample_temp0.7_seed1.csv")


def test_run_timing_loop_and_grouped_mean(tmp_path):
    calls = {"warm": 0, "bssd": [], "tg": [], "tf": 0}
    target = types.SimpleNamespace(generation_config=types.SimpleNamespace(num_beams=5))
    draft = types.SimpleNamespace(generation_config=types.SimpleNamespace(num_beams=40))

    def bssd(t, d, inputs, gamma, max_new_tokens, prefix_allowed_tokens_fn=None):
        assert t is target and d is draft and gamma == 3 and max_new_tokens == 4 and prefix_allowed_tokens_fn == "fn"
        u = inputs["user"]
        calls["bssd"].append(u)
        return {"time_cost": 0.01 * (u + 1), "draft_time_cost": 0.001, "target_time_cost": 0.002 * (u + 1),
                "verify_time_cost": 0.0005, "n_run": 1 + u % 3, "total_accept_steps": u % 4, "total_accept_tokens": 5 * (u % 4),
                "ave_accept_tokens": 5 * (u % 4) / (1 + u % 3)}

    def tgen(t, inputs, max_new_tokens, prefix_allowed_tokens_fn=None):
        calls["tg"].append(inputs["user"])
        return {"time_cost": 0.04}

    def tf(inputs):
        calls["tf"] += 1
        return 0.09

    def warm(inputs):
        calls["warm"] += 1
        assert inputs["user"] == 2                                           # the first user inside [stop_l, stop_r)

    sink = run_timing(bssd, tgen, tf, [{"user": u} for u in range(10)], target, draft, 3, 4, "fn", "T", "D", stop_l=2,
                      stop_r=7, warmup=warm)
    assert calls["warm"] == 1 and calls["bssd"] == calls["tg"] == [2, 3, 4, 5, 6] and calls["tf"] == 5 and len(sink) == 5
    mean = sink.mean()
    assert list(mean.index.names) == GROUP_KEYS and list(mean.columns) == TIMING_COLUMNS[4:]
    assert mean.index[0] == ("T", "D", 5, 3)
    # the same reduction the reference applies: DataFrame.groupby(keys).mean() over per-user rows
    rows = [timing_row("T", "D", 5, 3, bssd(target, draft, {"user": u}, 3, 4, prefix_allowed_tokens_fn="fn"), {"time_cost": 0.04},
                       0.09, 4) for u in range(2, 7)]
    ref = pd.DataFrame(rows, columns=TIMING_COLUMNS).groupby(GROUP_KEYS).mean()
    np.testing.assert_allclose(mean.values.astype(float), ref.values.astype(float), rtol=1e-12)
    path = sink.write(str(tmp_path / csv_name("Beauty", "T", "D", 5, 40, 2, 7, False, 1.0, 2025)))
    back = pd.read_csv(path)
    assert list(back.columns) == TIMING_COLUMNS and len(back) == 1
    np.testing.assert_allclose(back["speedup"][0], np.mean([0.04 / (0.01 * (u + 1)) for u in range(2, 7)]))


def test_sink_rejects_malformed_rows():
    s = TimingSink()
    try:
        s.add([1, 2, 3])
    except ValueError:
        return
    raise AssertionError("short row accepted")
