"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of AtSpeed's speculative beam search, file:line relative to /root/reference/code:
  one_step_beam_search  beamSD.py:40-106     -> _expand()
  _draft_beam_search    beamSD.py:108-179    -> draft phase of bssd()
  _target_beam_search   beamSD.py:190-232    -> target phase of bssd()
  verify (greedy)       beamSD.py:278-330,370-380  -> _verify_strict()
  verify (sampling)     beamSD.py:293-321,332-369  -> _verify_relaxed()
  BSSD                  beamSD.py:458-542    -> bssd()
  target_generate       beamSD.py:544-595    -> target_generate()
  PrefixConstrainedLogitsProcessor (transformers, via beamSD.py:62,64,288,291) -> _allowed_mask()

It is written against an explicit beam TREE (nodes with parent pointers, per-model KV slots) instead
of the reference's dense masks + sliced caches: every beam attends the prompt and its own ancestor
chain, which is exactly the set of slots the reference's additive mask leaves at 0.  Pinned against
the unmodified reference (tools/make_golden.py -> tests/golden/*.json, tests/test_oracle_bssd.py).

Defined divergences from the reference:
  * the token-range filter `tok >= 32000 | tok == 2` is restated as "drop non-finite candidates";
  * ties are broken towards the lowest flat index (row * V + token); torch.topk's choice is
    implementation-defined, and exact ties have measure zero with fp32 random weights;
  * gamma = 1 works (the reference crashes on a cache/mask length mismatch).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .llama_ref import RefCache, RefLlama

NEG_INF = float("-inf")
GAP_LOG = None   # set to a list to record, for every greedy top-k, the score margin at the cut-off (near-tie diagnostics)
# set to a list to record, for every greedy top-k, {"cand": {generated-sequence tuple: score} over ALL finite candidates,
# "kept": [sequence tuples in rank order]} -- tests/test_gpu_e2e.py locates the level where a bf16 run leaves the oracle's
# trajectory and checks the margin THERE
LEVEL_LOG = None


# ----------------------------------------------------------------------------------------------
# beam tree
# ----------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("tok", "parent", "depth", "score", "slot")

    def __init__(self, tok, parent, score=0.0):
        self.tok, self.parent, self.score = tok, parent, score
        self.depth = 0 if parent is None else parent.depth + 1
        self.slot = {}  # model key -> KV slot index (prompt node: list of slots)

    def gen(self) -> List[int]:
        out, n = [], self
        while n.parent is not None:
            out.append(n.tok)
            n = n.parent
        return out[::-1]

    def chain(self) -> List["Node"]:
        out, n = [], self
        while n.parent is not None:
            out.append(n)
            n = n.parent
        return out[::-1]


class _ModelState:
    """One model + its append-only KV cache, addressed through tree nodes."""

    def __init__(self, key: str, model: RefLlama, prompt: Sequence[int]):
        self.key, self.model, self.prompt = key, model, list(prompt)
        self.cache = model.new_cache()
        self.P = len(prompt)
        self.n_forward = 0
        self.tokens_forwarded = 0

    def _run(self, tokens, pos, vis, rows):
        self.n_forward += 1
        self.tokens_forwarded += len(tokens)
        return self.model.forward(torch.tensor(tokens), torch.tensor(pos), vis, self.cache,
                                  torch.tensor(rows))

    def forward_prompt(self, root: Node) -> torch.Tensor:
        P = self.P
        vis = torch.tril(torch.ones(P, P, dtype=torch.bool))
        root.slot[self.key] = list(range(P))
        return self._run(self.prompt, list(range(P)), vis, [P - 1])

    def forward_nodes(self, nodes: List[Node], want: List[Node]) -> torch.Tensor:
        """Append `nodes` (every ancestor must already have a slot or precede it in `nodes`),
        return fp32 logits rows for `want` (subset of nodes)."""
        S0 = len(self.cache)
        T = len(nodes)
        for j, n in enumerate(nodes):
            assert self.key not in n.slot
            n.slot[self.key] = S0 + j
        vis = torch.zeros(T, S0 + T, dtype=torch.bool)
        vis[:, : self.P] = True
        for j, n in enumerate(nodes):
            for a in n.chain():
                vis[j, a.slot[self.key]] = True
        idx = {id(n): j for j, n in enumerate(nodes)}
        rows = [idx[id(n)] for n in want]
        pos = [self.P - 1 + n.depth for n in nodes]
        return self._run([n.tok for n in nodes], pos, vis, rows)

    def missing_ancestors(self, nodes: List[Node]) -> List[Node]:
        seen, out = set(), []
        for n in nodes:
            for a in n.chain()[:-1]:
                if self.key not in a.slot and id(a) not in seen:
                    seen.add(id(a))
                    out.append(a)
        out.sort(key=lambda a: a.depth)
        return out


# ----------------------------------------------------------------------------------------------
# constraint + selection
# ----------------------------------------------------------------------------------------------
def _allowed_mask(fn: Callable, prompt: Sequence[int], frontier: List[Node], V: int) -> torch.Tensor:
    """PrefixConstrainedLogitsProcessor restated: additive mask, 0 on allowed ids, -inf elsewhere;
    `fn(batch_id, sentence)` gets the beam's full sequence (prompt + generated)."""
    mask = torch.full((len(frontier), V), NEG_INF)
    if fn is None:
        return torch.zeros(len(frontier), V)
    for r, n in enumerate(frontier):
        allowed = fn(0, torch.tensor(list(prompt) + n.gen()))
        if allowed is None or len(allowed) == 0:
            raise ValueError("empty constraint")  # HF raises the same
        mask[r, torch.tensor(list(allowed), dtype=torch.long)] = 0.0
    return mask


def _warp(scores: torch.Tensor, temperature: float, top_k: Optional[int], num_beams: int) -> torch.Tensor:
    """transformers 4.41 `_get_logits_warper` for a default sampling config: temperature (if != 1)
    then TopKLogitsWarper(top_k, min_tokens_to_keep = 2 if num_beams > 1 else 1)."""
    if temperature is not None and temperature != 1.0:
        scores = scores / temperature
    if top_k:
        k = min(max(top_k, 2 if num_beams > 1 else 1), scores.shape[-1])
        kth = torch.topk(scores, k)[0][..., -1, None]
        scores = scores.masked_fill(scores < kth, NEG_INF)
    return scores


def _topk_lowest_index(flat: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """top-k by value, ties -> lowest flat index; non-finite candidates dropped (G4)."""
    order = torch.sort(flat, descending=True, stable=True)[1][:k]
    vals = flat[order]
    keep = torch.isfinite(vals)
    return vals[keep], order[keep]


# draw sites (the CUDA path's noise streams are named the same way: csrc/noise.cuh)
SITE_DRAFT, SITE_ACCEPT, SITE_PERM, SITE_RESIDUAL, SITE_BONUS, SITE_STEP = 0, 1, 2, 3, 4, 5


class ReferenceUndefined(RuntimeError):
    """The reference's behaviour is undefined here (its torch.multinomial call raises / samples disallowed ids)."""


@dataclass
class SamplingCfg:
    """Randomness of the sampling branch.  Two modes:
      * generator mode (default): the same torch calls, in the same order, on the same generators as the
        reference -> reproduces the reference draw for draw under one torch.manual_seed (tests/golden/bssd_relaxed.json);
      * noise mode (`noise_fn` set): every draw is a pure function of (site, round, level, index), the contract of the
        CUDA path.  noise_fn(kind, site, round, level, n) -> tensor[n]; kind "exp" (Exp(1) floats), "uniform"
        (floats in (0,1)) or "bits" (integers).  A multinomial without replacement is the top-n of p / exp-noise
        (what ATen computes from its generator), restricted to p > 0."""
    temperature: float = 1.0
    top_k: Optional[int] = 50
    generator: Optional[torch.Generator] = None          # device generator of the reference (CPU here)
    cpu_generator: Optional[torch.Generator] = None      # torch.randperm's generator (beamSD.py:343)
    noise_fn: Optional[Callable] = None
    # empty residual distribution: False = behave like the reference (raise / uniform over everything),
    # True = the defined behaviour of the CUDA path (draw the missing beams from p itself)
    defined_fallback: bool = False
    round: int = 0                                        # maintained by bssd()
    fallbacks: int = 0

    def multinomial(self, p: torch.Tensor, n: int, site: int = 0, level: int = 0) -> torch.Tensor:
        if self.noise_fn is not None:
            noise = self.noise_fn("exp", site, self.round, level, p.numel()).to(p.dtype).view(-1)
            order = _topk_lowest_index(p / noise, n)[1]
            return order[p[order] > 0]
        return torch.multinomial(p, n, generator=self.generator)

    def uniform(self, n: int, level: int = 0) -> torch.Tensor:
        if self.noise_fn is not None:
            return self.noise_fn("uniform", SITE_ACCEPT, self.round, level, n).view(-1)
        return torch.rand(n, generator=self.generator)

    def random_subset(self, accepted_pos: torch.Tensor, n_picks: int, k: int, level: int = 0) -> torch.Tensor:
        """Indices (into the accepted list) of a uniformly random k-subset: `torch.randperm(n)[:k]` (beamSD.py:343),
        or in noise mode the k accepted picks with the smallest (bits, position) keys."""
        if self.noise_fn is not None:
            bits = self.noise_fn("bits", SITE_PERM, self.round, level, n_picks).view(-1).to(torch.int64)
            key = bits[accepted_pos] * 1024 + accepted_pos.to(torch.int64)     # (bits, position) order; bits < 2**32
            return torch.argsort(key)[:k]
        return torch.randperm(len(accepted_pos), generator=self.cpu_generator)[:k]


def _expand(logits_rows: torch.Tensor, frontier: List[Node], width: int, fn, prompt, V,
            sampling: Optional[SamplingCfg], num_beams_for_warp: int, site: int = SITE_STEP, level: int = 0):
    """One beam-search step over `frontier` (beamSD.py:57-86): full-vocab log_softmax, constraint
    mask applied afterwards, + parent score, flatten, top-`width` (or multinomial), split into
    (parent row, token). Returns children nodes (rank / sample order) and q (sampling only)."""
    logp = torch.log_softmax(logits_rows.float(), -1)
    scores = logp + _allowed_mask(fn, prompt, frontier, V)
    if sampling is not None:
        scores = _warp(scores, sampling.temperature, sampling.top_k, num_beams_for_warp)
    parent = torch.tensor([n.score for n in frontier], dtype=torch.float32)
    flat = (scores + parent[:, None]).reshape(-1)
    probs = None
    if sampling is not None:
        probs = torch.softmax(flat, -1)
        idx = sampling.multinomial(probs, width, site, level)
        vals = flat[idx]
        keep = torch.isfinite(vals)
        vals, idx = vals[keep], idx[keep]
    else:
        vals, idx = _topk_lowest_index(flat, width + 1)
        if GAP_LOG is not None and len(vals) > width:
            GAP_LOG.append(float(vals[width - 1] - vals[width]))   # margin between the last kept and first dropped
        vals, idx = vals[:width], idx[:width]
        if LEVEL_LOG is not None:
            fin = torch.nonzero(torch.isfinite(flat)).view(-1)
            gens = [tuple(n.gen()) for n in frontier]
            LEVEL_LOG.append({"cand": {gens[int(i // V)] + (int(i % V),): float(flat[i]) for i in fin},
                              "kept": [gens[int(i // V)] + (int(i % V),) for i in idx]})
    kids = [Node(int(i % V), frontier[int(i // V)], float(v)) for v, i in zip(vals, idx)]
    return kids, idx, probs, flat


# ----------------------------------------------------------------------------------------------
# results
# ----------------------------------------------------------------------------------------------
@dataclass
class RoundTrace:
    draft_len: int
    n_roots: int
    # per draft level: (parent position in previous level / roots, token, score)
    draft_levels: List[List[Tuple[int, int, float]]] = field(default_factory=list)
    # per verified level: target picks (parent position in the *draft* level i list, token, score)
    target_picks: List[List[Tuple[int, int, float]]] = field(default_factory=list)
    hits: List[List[int]] = field(default_factory=list)   # per level, hit positions in the draft list
    n_matches: int = 0


@dataclass
class BSSDResult:
    sequences: np.ndarray            # [K, P + L] int64
    scores: np.ndarray               # [K] float32
    n_run: int = 0
    accept_steps: List[int] = field(default_factory=list)
    rounds: List[RoundTrace] = field(default_factory=list)
    n_target_forward: int = 0
    n_draft_forward: int = 0
    target_tokens: int = 0

    @property
    def total_accept_steps(self):
        return int(sum(self.accept_steps))

    def stats(self, K):
        n_run = max(self.n_run, 1)
        return {"n_run": self.n_run, "total_accept_steps": self.total_accept_steps,
                "total_accept_tokens": self.total_accept_steps * K,
                "ave_accept_tokens": self.total_accept_steps * K / n_run}


def _finish(prompt, beams: List[Node], res_kw) -> BSSDResult:
    seqs = np.asarray([list(prompt) + b.gen() for b in beams], dtype=np.int64)
    scores = np.asarray([b.score for b in beams], dtype=np.float32)
    return BSSDResult(seqs, scores, **res_kw)


# ----------------------------------------------------------------------------------------------
# verify
# ----------------------------------------------------------------------------------------------
def _pos_in(level: List[Node], parent: Node, tok: int) -> int:
    for p, n in enumerate(level):
        if n.parent is parent and n.tok == tok:
            return p
    return -1


def _verify_strict(rows_of, roots, levels, K, fn, prompt, V, trace: RoundTrace):
    """AtSpeed-S (beamSD.py:278-330,370-380). rows_of(nodes) -> target logits rows for those nodes."""
    dl = len(levels)
    cur = roots
    m = 0
    for i in range(dl + 1):
        kids, idx, _, _ = _expand(rows_of(cur), cur, K, fn, prompt, V, None, K)
        # record picks with the parent expressed as its position in the draft's level-i list
        parent_pos = []
        for c in kids:
            if i == 0:
                parent_pos.append(roots.index(c.parent))
            else:
                parent_pos.append(levels[i - 1].index(c.parent))
        trace.target_picks.append([(pp, c.tok, c.score) for pp, c in zip(parent_pos, kids)])
        if i == dl:
            break
        level = levels[i]
        pos_of_pick = [_pos_in(level, c.parent, c.tok) for c in kids]
        hits = sorted(p for p in pos_of_pick if p >= 0)
        trace.hits.append(hits)
        if len(hits) == K:
            m += 1
            score_at = {p: c.score for p, c in zip(pos_of_pick, kids)}
            cur = [level[p] for p in hits]
            for n in cur:
                n.score = score_at[level.index(n)]     # carried score = the TARGET's (beamSD.py:296)
        else:
            break
    trace.n_matches = m
    return m, kids


def _verify_relaxed(rows_of, roots, levels, level_q, level_flat_idx, K, fn, prompt, V,
                    trace: RoundTrace, sampling: SamplingCfg):
    """AtSpeed-R sequence-level speculative sampling (beamSD.py:293-321,332-369)."""
    dl = len(levels)
    cur = roots
    m = 0
    for i in range(dl + 1):
        logp = torch.log_softmax(rows_of(cur).float(), -1)
        scores = _warp(logp + _allowed_mask(fn, prompt, cur, V), sampling.temperature, sampling.top_k, K)
        parent = torch.tensor([n.score for n in cur], dtype=torch.float32)
        flat_t = (scores + parent[:, None]).reshape(-1)
        if i == dl:  # bonus level: sample K from the target
            p = torch.softmax(flat_t, -1)
            idx = sampling.multinomial(p, K, SITE_BONUS, i)
            kids = [Node(int(j % V), cur[int(j // V)], float(flat_t[j])) for j in idx]
            break
        prev = roots if i == 0 else levels[i - 1]
        q = level_q[i]
        if i > 0:  # scatter the target rows into the draft's [n_prev, V] index space (:309-321)
            full = torch.full((len(prev), V), NEG_INF)
            rows = torch.tensor([prev.index(n) for n in cur])
            full[rows] = flat_t.view(len(cur), V)
            flat_d = full.reshape(-1)
        else:
            flat_d = flat_t
        p = torch.softmax(flat_d, -1)
        p = torch.nan_to_num(p, nan=0.0)
        q = torch.nan_to_num(q, nan=0.0)
        picks = level_flat_idx[i]
        ratio = p[picks] / q[picks]
        r = sampling.uniform(len(picks), i)
        acc = r <= ratio
        acc_tok = picks[acc]
        trace.hits.append([int(j) for j in torch.nonzero(acc).view(-1)])
        if int(acc.sum()) >= K:
            m += 1
            sel = acc_tok[sampling.random_subset(torch.nonzero(acc).view(-1), len(picks), K, i)].sort()[0]
            level = levels[i]
            pos = [int(torch.nonzero(picks == y).view(-1)[0]) for y in sel]
            cur = [level[pp] for pp in pos]
            for n, y in zip(cur, sel):
                n.score = float(flat_d[y])
            trace.target_picks.append([(int(y // V), int(y % V), float(flat_d[y])) for y in sel])
        else:
            newp = torch.clamp(p - q, min=0)
            newp[acc_tok] = 0
            if float(newp.sum()) == 0:
                sampling.fallbacks += 1
                if sampling.defined_fallback:      # the CUDA path's defined behaviour: draw from p itself
                    newp = p.clone()
                    newp[acc_tok] = 0
                elif i == 0:
                    newp[...] = torch.finfo(newp.dtype).tiny
                else:
                    # the reference's fill (:355-357) writes into a copy and is a no-op: its multinomial then raises
                    raise ReferenceUndefined("empty residual distribution at level %d" % i)
            else:
                newp = newp / newp.sum()
            extra = sampling.multinomial(newp, K - int(acc.sum()), SITE_RESIDUAL, i)
            sel = torch.cat((acc_tok, extra)).sort()[0]
            kids = [Node(int(y % V), prev[int(y // V)], float(flat_d[y])) for y in sel]
            trace.target_picks.append([(int(y // V), int(y % V), float(flat_d[y])) for y in sel])
            break
    trace.n_matches = m
    return m, kids


# ----------------------------------------------------------------------------------------------
# drivers
# ----------------------------------------------------------------------------------------------
def bssd(target: RefLlama, draft: RefLlama, prompt: Sequence[int], K: int, N: int, gamma: int,
         max_new_tokens: int, fn: Optional[Callable], sampling: Optional[SamplingCfg] = None) -> BSSDResult:
    V = target.shape.vocab
    prompt = list(prompt)
    tgt, dft = _ModelState("t", target, prompt), _ModelState("d", draft, prompt)
    root = Node(None, None, 0.0)
    roots = [root]
    first = True
    done = 0
    accept, rounds = [], []
    beams = None
    while done < max_new_tokens:
        dl = min(gamma, max_new_tokens - done - 1)
        if sampling is not None:
            sampling.round = len(accept)
        if dl == 0:  # one plain target step (beamSD.py:505-509)
            rows = tgt.forward_prompt(root) if first else tgt.forward_nodes(roots, roots)
            beams, _, _, _ = _expand(rows, roots, K, fn, prompt, V, sampling, K, SITE_STEP, 0)
            break
        tr = RoundTrace(dl, len(roots))
        # 1. draft: dl beam-search steps of width N
        levels, level_q, level_idx = [], [], []
        frontier = roots
        for j in range(dl):
            if j == 0:
                if first:
                    rows = dft.forward_prompt(root)
                else:
                    batch = dft.missing_ancestors(roots) + roots
                    rows = dft.forward_nodes(batch, roots)
            else:
                rows = dft.forward_nodes(frontier, frontier)
            kids, idx, q, _ = _expand(rows, frontier, N, fn, prompt, V, sampling, K, SITE_DRAFT, j)
            tr.draft_levels.append([(frontier.index(c.parent), c.tok, c.score) for c in kids])
            levels.append(kids), level_q.append(q), level_idx.append(idx)
            frontier = kids
        # 2. target: ONE forward over roots + every draft level (beamSD.py:203-224)
        flat_nodes = [n for lv in levels for n in lv]
        if first:
            r0 = tgt.forward_prompt(root)
            rl = tgt.forward_nodes(flat_nodes, flat_nodes)
            tgt.n_forward -= 1  # prompt + tree are a single forward in the reference
            table = {id(root): r0[0]}
            want = flat_nodes
        else:
            rl = tgt.forward_nodes(roots + flat_nodes, roots + flat_nodes)
            table = {}
            want = roots + flat_nodes
        for n, row in zip(want, rl):
            table[id(n)] = row
        rows_of = lambda nodes: torch.stack([table[id(n)] for n in nodes])
        # 3. verify
        if sampling is None:
            m, beams = _verify_strict(rows_of, roots, levels, K, fn, prompt, V, tr)
        else:
            m, beams = _verify_relaxed(rows_of, roots, levels, level_q, level_idx, K, fn, prompt, V,
                                       tr, sampling)
        rounds.append(tr)
        accept.append(m)
        done += m + 1
        roots, first = beams, False
    if sampling is not None:  # beamSD.py:529-531
        beams = sorted(beams, key=lambda b: -b.score)
    return _finish(prompt, beams, dict(n_run=len(accept), accept_steps=accept, rounds=rounds,
                                       n_target_forward=tgt.n_forward, n_draft_forward=dft.n_forward,
                                       target_tokens=tgt.tokens_forwarded))


def target_generate(target: RefLlama, prompt: Sequence[int], K: int, max_new_tokens: int,
                    fn: Optional[Callable], sampling: Optional[SamplingCfg] = None) -> BSSDResult:
    """Plain tree-mask beam search on the target (beamSD.py:544-595): max_new_tokens x one step."""
    V = target.shape.vocab
    prompt = list(prompt)
    tgt = _ModelState("t", target, prompt)
    root = Node(None, None, 0.0)
    frontier = [root]
    if sampling is not None:
        sampling.round = 0
    for step in range(max_new_tokens):
        rows = tgt.forward_prompt(root) if step == 0 else tgt.forward_nodes(frontier, frontier)
        frontier, _, _, _ = _expand(rows, frontier, K, fn, prompt, V, sampling, K, SITE_STEP, step)
    if sampling is not None:
        frontier = sorted(frontier, key=lambda b: -b.score)
    return _finish(prompt, frontier, dict(n_target_forward=tgt.n_forward, target_tokens=tgt.tokens_forwarded))
