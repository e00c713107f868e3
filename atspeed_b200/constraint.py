"""Compile a `prefix_allowed_tokens_fn(batch_id, sentence) -> List[int]` into a CSR child table.

The reference evaluates the constraint in Python once per beam row per step, each call starting
with a device->host `sentence.tolist()` (transformers PrefixConstrainedLogitsProcessor reached from
/root/reference/code/beamSD.py:62,64,288,291; constraint functions at code/generation_trie.py:92-98,
code/generate_teacher_data.py:174-188 and code/data.py:84-104).  Here the function is compiled ONCE
into a flat table that lives in HBM and every beam only carries a node id:

    child_off [n_nodes + 1] int32     children of node n are edges child_off[n] .. child_off[n+1]-1
    child_tok [n_edges]     int32     token id of the edge, ascending within a node
    child_node[n_edges]     int32     node reached through the edge (-1 beyond `depth`)

Node 0 is the state right after the prompt; nodes are numbered breadth first, so depth d is a
contiguous id range.  Three ways to get there, tried in this order:
  1. a trie reachable from the callable (attribute `candidate_trie` or a closure cell holding an
     object with `.trie_dict`) is flattened directly, then checked against the callable by probes;
  2. a `{depth: tokens}` dict reachable the same way (attribute / closure `allowed_tokens`) gives one
     node per depth (positional constraint);
  3. any other callable is probed breadth first from the given prompt (budgeted); if all nodes of a
     depth answer identically it is collapsed to one node per depth.
Every result is verified for prompt independence on a second prompt when one is supplied.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch


@dataclass
class CSRTrie:
    child_off: np.ndarray
    child_tok: np.ndarray
    child_node: np.ndarray
    depth: int
    kind: str = "trie"

    @property
    def n_nodes(self) -> int:
        return len(self.child_off) - 1

    @property
    def n_edges(self) -> int:
        return len(self.child_tok)

    @property
    def max_fanout(self) -> int:
        return int(np.max(np.diff(self.child_off))) if self.n_nodes else 0

    def children(self, node: int) -> np.ndarray:
        return self.child_tok[self.child_off[node]:self.child_off[node + 1]]

    def walk(self, tokens: Iterable[int]) -> int:
        node = 0
        for t in tokens:
            lo, hi = self.child_off[node], self.child_off[node + 1]
            j = lo + int(np.searchsorted(self.child_tok[lo:hi], t))
            if j >= hi or self.child_tok[j] != t:
                return -1
            node = int(self.child_node[j])
            if node < 0:
                return -1
        return node


def _finalize(levels: List[List[List[int]]], child_ids: List[List[List[int]]], depth: int, kind: str) -> CSRTrie:
    off, tok, nxt = [0], [], []
    for lv_tok, lv_node in zip(levels, child_ids):
        for toks, nodes in zip(lv_tok, lv_node):
            order = np.argsort(np.asarray(toks, dtype=np.int64), kind="stable") if len(toks) else []
            tok += [int(toks[i]) for i in order]
            nxt += [int(nodes[i]) for i in order]
            off.append(len(tok))
    return CSRTrie(np.asarray(off, np.int32), np.asarray(tok, np.int32), np.asarray(nxt, np.int32), depth, kind)


def csr_from_trie_dict(root: Dict[int, dict], depth: int) -> CSRTrie:
    """Breadth-first flatten of a nested-dict trie (reference Trie.trie_dict) below `root`."""
    levels, child_ids = [], []
    frontier = [root]
    next_id = 1
    for d in range(depth + 1):
        lv_tok, lv_node, nxt_frontier = [], [], []
        for node in frontier:
            toks = list(node.keys()) if d < depth else []
            ids = []
            for t in toks:
                if d + 1 <= depth:
                    ids.append(next_id)
                    next_id += 1
                    nxt_frontier.append(node[t])
                else:
                    ids.append(-1)
            lv_tok.append(toks), lv_node.append(ids)
        levels.append(lv_tok), child_ids.append(lv_node)
        frontier = nxt_frontier
        if not frontier:
            break
    return _finalize(levels, child_ids, depth, "trie")


def csr_from_positional(allowed: Dict[int, Iterable[int]], depth: int) -> CSRTrie:
    levels, child_ids = [], []
    for d in range(depth + 1):
        toks = sorted(int(t) for t in allowed.get(d, [])) if d < depth else []
        levels.append([toks])
        child_ids.append([[d + 1] * len(toks)])
    return _finalize(levels, child_ids, depth, "positional")


def _reachable(fn: Callable, name_pred) -> Optional[object]:
    for attr in ("candidate_trie", "allowed_tokens", "trie"):
        if hasattr(fn, attr) and name_pred(getattr(fn, attr)):
            return getattr(fn, attr)
    for cell in getattr(fn, "__closure__", None) or ():
        try:
            v = cell.cell_contents
        except ValueError:
            continue
        if name_pred(v):
            return v
        # one hop: objects such as the reference's SeqRecDataset hold `allowed_tokens`
        for attr in ("allowed_tokens", "trie_dict"):
            inner = getattr(v, attr, None)
            if inner is not None and name_pred(inner):
                return inner
            if attr == "trie_dict" and inner is not None and name_pred(v):
                return v
    return None


def _is_trie(o) -> bool:
    return hasattr(o, "trie_dict") and isinstance(getattr(o, "trie_dict"), dict)


def _is_positional(o) -> bool:
    return isinstance(o, dict) and len(o) > 0 and all(isinstance(k, int) for k in o) and \
        all(isinstance(v, (set, list, tuple, frozenset)) for v in o.values())


def _call(fn, prompt: Sequence[int], suffix: Sequence[int]) -> List[int]:
    out = fn(0, torch.tensor(list(prompt) + list(suffix), dtype=torch.long))
    return sorted(int(t) for t in (out or []))


def _check(csr: CSRTrie, fn, prompt, rng: np.random.Generator, n_paths: int = 24) -> bool:
    """Compare the table with the callable along random root-to-leaf paths."""
    for _ in range(n_paths):
        node, suffix = 0, []
        for d in range(csr.depth):
            kids = csr.children(node)
            if _call(fn, prompt, suffix) != [int(t) for t in kids]:
                return False
            if len(kids) == 0:
                break
            j = int(rng.integers(len(kids)))
            suffix.append(int(kids[j]))
            node = int(csr.child_node[csr.child_off[node] + j])
            if node < 0:
                break
    return True


def _probe(fn, prompt, depth: int, budget: int) -> CSRTrie:
    levels, child_ids = [], []
    frontier: List[List[int]] = [[]]
    next_id, calls = 1, 0
    collapsed = False
    for d in range(depth + 1):
        lv_tok, lv_node, nxt_frontier = [], [], []
        answers = []
        for suffix in frontier:
            if d < depth:
                calls += 1
                if calls > budget:
                    raise RuntimeError(f"constraint probe exceeded {budget} calls; pass a Trie or a positional dict")
                answers.append(_call(fn, prompt, suffix))
            else:
                answers.append([])
        if d < depth and len(answers) > 8 and all(a == answers[0] for a in answers):
            # every node of this depth allows the same set: positional from here on
            collapsed = True
            rest = {0: answers[0]}
            suffix = list(frontier[0])
            for dd in range(1, depth - d):
                suffix = suffix + [rest[dd - 1][0]]
                rest[dd] = _call(fn, prompt, suffix)
            n_here = len(frontier)
            base = next_id  # ids of the shared chain nodes
            for _ in range(n_here):
                lv_tok.append(list(rest[0]))
                lv_node.append([base if d + 1 <= depth else -1] * len(rest[0]))
            levels.append(lv_tok), child_ids.append(lv_node)
            for dd in range(1, depth - d + 1):
                toks = list(rest.get(dd, [])) if d + dd < depth else []
                levels.append([toks])
                child_ids.append([[base + dd] * len(toks)])
            break
        for toks in answers:
            ids = []
            for t in toks:
                if d + 1 <= depth:
                    ids.append(next_id)
                    next_id += 1
                else:
                    ids.append(-1)
            lv_tok.append(toks), lv_node.append(ids)
        for suffix, toks in zip(frontier, answers):
            nxt_frontier += [suffix + [t] for t in toks]
        levels.append(lv_tok), child_ids.append(lv_node)
        frontier = nxt_frontier
        if not frontier:
            break
    return _finalize(levels, child_ids, depth, "probed-positional" if collapsed else "probed")


_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()
_UNCONSTRAINED: dict = {}


def compile_constraint(fn: Optional[Callable], prompt: Sequence[int], depth: int,
                       other_prompt: Optional[Sequence[int]] = None, vocab_size: Optional[int] = None,
                       budget: int = 400_000, use_cache: bool = True) -> CSRTrie:
    """Compile `fn` for `depth` generated tokens. `fn=None` (no constraint) needs `vocab_size` and
    yields one node per depth allowing every token."""
    if fn is None:
        # cached per (V, depth): callers key device tries and sessions on the table's identity (beamSD.get_session), so a
        # fresh table per call would build a fresh multi-GB session per search
        if vocab_size is None:
            raise ValueError("compile_constraint(None, ...) needs vocab_size")
        key = (int(vocab_size), int(depth))
        if not use_cache or key not in _UNCONSTRAINED:
            csr = csr_from_positional({d: range(vocab_size) for d in range(depth)}, depth)
            if not use_cache:
                return csr
            _UNCONSTRAINED[key] = csr
        return _UNCONSTRAINED[key]
    if use_cache:
        try:
            hit = _CACHE.get(fn, {}).get(depth)
        except TypeError:
            hit = None
        if hit is not None:
            return hit
    rng = np.random.default_rng(0)
    csr = None
    trie = _reachable(fn, _is_trie)
    if trie is not None:
        root0 = _call(fn, prompt, [])
        candidates = [trie.trie_dict] + [v for v in trie.trie_dict.values() if isinstance(v, dict)]
        for root in candidates:
            if sorted(int(t) for t in root.keys()) == root0:
                cand = csr_from_trie_dict(root, depth)
                if _check(cand, fn, prompt, rng):
                    csr = cand
                    break
    if csr is None:
        pos = _reachable(fn, _is_positional)
        if pos is not None:
            cand = csr_from_positional(pos, depth)
            if _check(cand, fn, prompt, rng):
                csr = cand
    if csr is None:
        csr = _probe(fn, prompt, depth, budget)
    if other_prompt is not None and not _check(csr, fn, other_prompt, rng):
        raise ValueError("prefix_allowed_tokens_fn depends on the prompt; cannot be compiled to one table")
    if use_cache:
        try:
            _CACHE.setdefault(fn, {})[depth] = csr
        except TypeError:
            pass
    return csr
