"""Result / timing sink of the per-user inference loop.

The reference's driver collects one row per user and writes the per-configuration mean as a CSV
(/root/reference/code/inference.py:152-156 columns, :179-187 row, :189 file name).  Downstream analysis reads
those column names (`speedup`, `speedupTF`, `overhead`, `ave_accept_tokens`, ...), so this module keeps the schema,
the derived quantities and the file naming; `run_timing` is the loop of :162-187 over any `BSSD` / `target_generate` /
HF-generate callables (the drop-in ones of atspeed_b200.beamSD, or the reference's own).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List, Optional

import pandas as pd

# code/inference.py:152-156
TIMING_COLUMNS = ["target_model", "draft_model", "beam_size", "gamma",
                  "draft_time_cost", "target_time_cost", "verify_time_cost", "total_time_cost", "generalBS_time_cost",
                  "TF_cache_time_cost",
                  "speedup", "speedupTF", "total_accept_steps", "total_accept_tokens", "ave_accept_tokens", "overhead"]
GROUP_KEYS = ["target_model", "draft_model", "beam_size", "gamma"]


def timing_row(target_model_name: str, draft_model_name: str, beam_size: int, gamma: int, outputs: Dict,
               target_outputs: Dict, tf_time_cost: float, max_new_tokens: int) -> List:
    """One user's row (code/inference.py:179-187): `outputs` = BSSD's dict, `target_outputs` = target_generate's,
    `tf_time_cost` = wall time of HF `generate` (the `TF_target` timer)."""
    speedup = target_outputs["time_cost"] / outputs["time_cost"]
    speedup_tf = tf_time_cost / outputs["time_cost"]
    overhead = (outputs["time_cost"] * max_new_tokens) / (target_outputs["time_cost"] * outputs["n_run"])
    return [target_model_name, draft_model_name, beam_size, gamma,
            outputs["draft_time_cost"], outputs["target_time_cost"], outputs["verify_time_cost"], outputs["time_cost"],
            target_outputs["time_cost"], tf_time_cost,
            speedup, speedup_tf, outputs["total_accept_steps"], outputs["total_accept_tokens"], outputs["ave_accept_tokens"],
            overhead]


def csv_name(dataset: str, target_model_name: str, draft_model_name: str, beam_size: int, draft_beam_size: int,
             stop_l: int, stop_r: int, do_sample: bool, temperature: float, seed: int) -> str:
    """Relative path of the mean-timing CSV (code/inference.py:189)."""
    suffix = "do_sample" if do_sample else ""
    return os.path.join("AnaResult", dataset,
                        f"timing_mean_{target_model_name}_{draft_model_name}_B{beam_size}-{draft_beam_size}_{stop_l}-{stop_r}_"
                        f"{suffix}_temp{temperature}_seed{seed}.csv")


class TimingSink:
    """Per-user rows of one (target, draft, beam_size) sweep and their grouped mean."""

    def __init__(self):
        self.frame = pd.DataFrame(columns=TIMING_COLUMNS)

    def add(self, row: List):
        if len(row) != len(TIMING_COLUMNS):
            raise ValueError(f"a timing row has {len(TIMING_COLUMNS)} fields, got {len(row)}")
        self.frame.loc[len(self.frame)] = row

    def __len__(self):
        return len(self.frame)

    def mean(self) -> pd.DataFrame:
        num = self.frame.copy()
        for c in TIMING_COLUMNS:
            if c not in GROUP_KEYS[:2]:
                num[c] = pd.to_numeric(num[c])
        return num.groupby(GROUP_KEYS).mean()

    def write(self, path: str) -> str:
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self.mean().to_csv(path)
        return path


def run_timing(bssd: Callable[..., Dict], target_generate: Callable[..., Dict], tf_generate: Optional[Callable[[Dict], float]],
               batches: Iterable[Dict], target_model, draft_model, gamma: int, max_new_tokens: int,
               prefix_allowed_tokens_fn, target_model_name: str = "target", draft_model_name: str = "draft",
               stop_l: int = 0, stop_r: Optional[int] = None, warmup: Optional[Callable[[Dict], None]] = None) -> TimingSink:
    """The user loop of code/inference.py:162-187: users [stop_l, stop_r) one at a time; the first processed user triggers
    `warmup(inputs)` (the reference's two untimed `generate` calls); every user runs BSSD, target_generate and -- when
    `tf_generate` is given -- HF beam search, whose wall time it returns.  K is read from the target's generation_config
    like the reference does."""
    sink = TimingSink()
    beam_size = int(target_model.generation_config.num_beams)
    warmed = False
    for step, inputs in enumerate(batches):
        if step < stop_l:
            continue
        if stop_r is not None and step >= stop_r:
            break
        if not warmed and warmup is not None:
            warmup(inputs)
        warmed = True
        outputs = bssd(target_model, draft_model, inputs, gamma, max_new_tokens, prefix_allowed_tokens_fn=prefix_allowed_tokens_fn)
        target_outputs = target_generate(target_model, inputs, max_new_tokens, prefix_allowed_tokens_fn=prefix_allowed_tokens_fn)
        tf_time = float(tf_generate(inputs)) if tf_generate is not None else float("nan")
        sink.add(timing_row(target_model_name, draft_model_name, beam_size, gamma, outputs, target_outputs, tf_time,
                            max_new_tokens))
    return sink
