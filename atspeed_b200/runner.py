"""User-sharded batch inference and result collection.

Each user is an independent speculative beam search (the reference's per-user loop,
/root/reference/code/inference.py:162-187, whose `--L/--R` range slice is the natural shard hook,
code/utils.py:146-147).  Rank r of W takes users r, r+W, r+2W, ... on its own GPU with a full
draft+target replica; the data path has no collective.  At the end ONE all-gather of fixed-stride
per-user records (items int32[K,L], scores fp32[K], {n_run, accept_steps, latency_us, valid}) makes the
ranked lists and metrics identical on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .metrics import computeTopNAccuracy


def shard_users(users: Sequence[int], rank: int, world: int) -> List[int]:
    """Interleaved partition: balances prompt length and round count better than contiguous blocks."""
    return [u for i, u in enumerate(users) if i % world == rank]


@dataclass
class UserRecords:
    users: np.ndarray        # [U] int32 user index
    items: np.ndarray        # [U, K, L] int32 generated code tokens, rank order
    scores: np.ndarray       # [U, K] float32
    meta: np.ndarray         # [U, 4] int32: n_run, total_accept_steps, latency_us, valid

    @staticmethod
    def empty(n: int, K: int, L: int) -> "UserRecords":
        return UserRecords(np.full(n, -1, np.int32), np.zeros((n, K, L), np.int32), np.zeros((n, K), np.float32),
                           np.zeros((n, 4), np.int32))

    def pack(self) -> np.ndarray:
        """Fixed-stride int32 record per user (scores bit-cast), ready for one all_gather."""
        U = len(self.users)
        return np.concatenate([self.users.reshape(U, 1), self.items.reshape(U, -1),
                               self.scores.view(np.int32).reshape(U, -1), self.meta], axis=1).astype(np.int32)

    @staticmethod
    def unpack(buf: np.ndarray, K: int, L: int) -> "UserRecords":
        U = buf.shape[0]
        o = 1
        items = buf[:, o:o + K * L].reshape(U, K, L); o += K * L
        scores = np.ascontiguousarray(buf[:, o:o + K]).view(np.float32); o += K
        return UserRecords(buf[:, 0].copy(), items.copy(), scores.copy(), buf[:, o:o + 4].copy())


def run_users(search: Callable[[List[int]], Dict], prompts: Callable[[int], List[int]], users: Sequence[int], K: int,
              L: int, sync: Optional[Callable[[], None]] = None) -> UserRecords:
    """Run `search(prompt_ids)` (e.g. Session.bssd bound to a gamma) for each user of this rank.
    `search` returns {"tokens": [k,L] int32, "scores": [k], "n_run", "total_accept_steps"}."""
    rec = UserRecords.empty(len(users), K, L)
    for i, u in enumerate(users):
        t0 = time.perf_counter()
        out = search(prompts(u))
        if sync is not None:
            sync()
        dt = time.perf_counter() - t0
        k = out["tokens"].shape[0]
        rec.users[i] = u
        rec.items[i, :k] = out["tokens"]
        rec.scores[i, :k] = out["scores"]
        rec.scores[i, k:] = -np.inf
        rec.meta[i] = (out.get("n_run", 0), out.get("total_accept_steps", 0), int(dt * 1e6), k)
    return rec


def common_prefix(prompts: Sequence[Sequence[int]], min_len: int = 8) -> List[int]:
    """The opening tokens every prompt shares (the instruction template of the reference's prompts, code/data.py:232-263),
    at most len(shortest prompt) - 1 of them; [] when fewer than `min_len` are shared."""
    if not prompts:
        return []
    n = min(len(p) for p in prompts) - 1
    first = prompts[0]
    for p in prompts[1:]:
        k = 0
        while k < n and p[k] == first[k]:
            k += 1
        n = k
    return list(first[:n]) if n >= min_len else []


def run_users_cohort(session, prompts: Callable[[int], List[int]], users: Sequence[int], gamma: int, K: int, L: int,
                     chunk: int = 256, share_prefix: bool = True) -> UserRecords:
    """The same loop through a cohort session (`Session(max_users > 1)`): this rank's users go through
    `session.bssd_batch` `chunk` at a time, their trees sharing forwards, and -- unless share_prefix is False -- the K/V of
    the tokens all their prompts open with computed once (`session.set_shared_prefix`).  A user's latency is its chunk's
    wall time divided by the chunk size (users of a cohort finish together)."""
    rec = UserRecords.empty(len(users), K, L)
    plist = [list(prompts(u)) for u in users]
    if hasattr(session, "set_shared_prefix"):
        session.set_shared_prefix(common_prefix(plist) if share_prefix else [])
    for c0 in range(0, len(users), chunk):
        t0 = time.perf_counter()
        outs = session.bssd_batch(plist[c0:c0 + chunk], gamma)
        dt = time.perf_counter() - t0
        for j, out in enumerate(outs):
            i = c0 + j
            k = out["tokens"].shape[0]
            rec.users[i] = users[i]
            rec.items[i, :k] = out["tokens"][:, :L]
            rec.scores[i, :k] = out["scores"]
            rec.scores[i, k:] = -np.inf
            rec.meta[i] = (out.get("n_run", 0), out.get("total_accept_steps", 0), int(dt * 1e6 / max(1, len(outs))), k)
    return rec


def gather_records(local: UserRecords, K: int, L: int, per_rank: int, device=None) -> UserRecords:
    """One all_gather of the padded per-rank record block; returns all users sorted by user index.
    Without an initialised process group (single process) this is the identity."""
    import torch.distributed as dist
    buf = local.pack()
    width = buf.shape[1]
    pad = np.zeros((per_rank, width), np.int32)
    pad[:, 0] = -1
    pad[: buf.shape[0]] = buf
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, t)
        allbuf = torch.cat(outs).cpu().numpy()
    else:
        allbuf = pad
    allbuf = allbuf[allbuf[:, 0] >= 0]
    allbuf = allbuf[np.argsort(allbuf[:, 0], kind="stable")]
    return UserRecords.unpack(allbuf, K, L)


def evaluate(dataset, rec: UserRecords, topN: Sequence[int]) -> Dict:
    """Recall/NDCG/... of the gathered ranked lists against the dataset's ground truth, plus the search
    statistics the reference logs per run (code/inference.py:152-156)."""
    gts = [dataset.ground_truth_strings(int(u)) for u in rec.users]
    preds = []
    for i in range(len(rec.users)):
        names = dataset.decode_items(rec.items[i])
        names += ["<none>"] * (max(topN) - len(names))
        preds.append(names)
    p, r, n, m = computeTopNAccuracy(gts, preds, list(topN))
    lat = rec.meta[:, 2].astype(np.float64) / 1e3
    n_run = np.maximum(rec.meta[:, 0], 1)
    K = rec.items.shape[1]
    return {"topN": list(topN), "precision": p, "recall": r, "ndcg": n, "mrr": m, "users": int(len(rec.users)),
            "latency_ms_p50": float(np.percentile(lat, 50)) if len(lat) else 0.0,
            "latency_ms_p95": float(np.percentile(lat, 95)) if len(lat) else 0.0,
            "ave_accept_tokens": float(np.mean(rec.meta[:, 1] * K / n_run)) if len(lat) else 0.0,
            "mean_n_run": float(np.mean(rec.meta[:, 0])) if len(lat) else 0.0}
