"""Shared test helpers: golden fixtures, oracle models, constraint functions."""
from __future__ import annotations

import functools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from atspeed_b200.generation_trie import (Trie, positional_prefix_allowed_tokens_fn,   # noqa: E402
                                          suffix_prefix_allowed_tokens_fn)
from atspeed_b200.prompts import RESPONSE_SEP, load_dataset                             # noqa: E402
from oracle import llama_ref as LR                                                      # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


@functools.lru_cache(None)
def golden(name="bssd_strict.json"):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@functools.lru_cache(None)
def dataset(name):
    return load_dataset(name)


@functools.lru_cache(None)
def constraint_fn(ds_name, kind):
    ds = dataset(ds_name)
    if kind == "strict":
        return suffix_prefix_allowed_tokens_fn(Trie(ds.strict_trie_sequences()), RESPONSE_SEP)
    if kind == "positional":
        return positional_prefix_allowed_tokens_fn(ds.positional_allowed(), RESPONSE_SEP)
    raise KeyError(kind)


@functools.lru_cache(None)
def stack_weights(stack, ds_name, which):
    """(shape, weights) exactly as tools/make_golden.py built them."""
    V = dataset(ds_name).vocab_size
    if stack == "hf_fp32":
        sh_t, sh_d = LR.shape_tiny_target(V), LR.shape_tiny_draft(V)
        std = 1.28 / 8
        Wt = LR.make_weights(sh_t, 0, std=std)
        if which == "target":
            return sh_t, Wt
        if which == "independent":
            return sh_d, LR.make_weights(sh_d, 1, std=std)
        return sh_d, LR.make_weights(sh_d, 2, std=std, like=Wt, noise=0.05)
    if stack == "ref_bf16":
        sh_t, sh_d = LR.shape_small_target(V), LR.shape_small_draft(V)
        Wt = LR.make_weights(sh_t, 10, std=1.28 / 16, dtype=torch.bfloat16)
        if which == "target":
            return sh_t, Wt
        if which == "independent":
            return sh_d, LR.make_weights(sh_d, 11, std=1.28 / (128 ** 0.5), dtype=torch.bfloat16)
        sh_dc = LR.LlamaShape(V, sh_t.hidden, 1, sh_t.n_heads, sh_t.mlp)
        return sh_dc, LR.make_weights(sh_dc, 12, std=1.28 / 16, dtype=torch.bfloat16, like=Wt, noise=0.03)
    raise KeyError(stack)


@functools.lru_cache(None)
def oracle_model(stack, ds_name, which):
    sh, W = stack_weights(stack, ds_name, which)
    return LR.RefLlama(sh, W, "fp32" if stack == "hf_fp32" else "bf16")


BF16_SCORE_TOL = 6e-2   # absolute, on cumulative log-probs of magnitude ~40 (= 1.5e-3 relative)


def lists_match(items_a, scores_a, items_b, scores_b, tol):
    """Ranked-list parity modulo numerical near-ties.

    Exact equality of the ranked item lists is required unless the disagreement is explained by
    scores closer than `tol`: an item present in only one list must sit within `tol` of the other
    list's cut-off score, and two lists may order a pair differently only if that pair's scores are
    within `tol`.  Returns (ok, n_exact_positions, message)."""
    a = [tuple(x) for x in items_a]
    b = [tuple(x) for x in items_b]
    sa, sb = dict(zip(a, scores_a)), dict(zip(b, scores_b))
    exact = sum(x == y for x, y in zip(a, b))
    if len(a) != len(b):
        return False, exact, f"lengths differ {len(a)} vs {len(b)}"
    for it in a:
        if it in sb:
            if abs(sa[it] - sb[it]) > tol:
                return False, exact, f"score of {it}: {sa[it]} vs {sb[it]}"
        elif sa[it] - min(scores_b) > tol:
            return False, exact, f"{it} only in first list, {sa[it]} vs cut-off {min(scores_b)}"
    for it in b:
        if it not in sa and sb[it] - min(scores_a) > tol:
            return False, exact, f"{it} only in second list, {sb[it]} vs cut-off {min(scores_a)}"
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y and abs(scores_a[i] - scores_b[i]) > tol:
            return False, exact, f"rank {i}: {x}@{scores_a[i]} vs {y}@{scores_b[i]}"
    return True, exact, ""
