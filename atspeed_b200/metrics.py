"""Ranking metrics of the reference, vectorised (reference code/utils.py:215-271 `computeTopNAccuracy`).

Same definition and the same `round(x, 4)` on the averaged values: users with an empty ground-truth list
are skipped; Precision@N = hits/N, Recall@N = hits/|GT|, NDCG@N with binary gains and an ideal DCG over
min(N, |GT|) positions, MRR = 1/rank of the first hit.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple


def computeTopNAccuracy(GroundTruth: Sequence[Sequence], predictedIndices: Sequence[Sequence], topN: Sequence[int],
                        rank=None) -> Tuple[List[float], List[float], List[float], List[float]]:
    precision, recall, ndcg, mrr = [], [], [], []
    disc = [1.0 / math.log2(j + 2) for j in range(max(topN) if topN else 0)]
    for n in topN:
        sp = sr = sn = sm = 0.0
        users = 0
        for gt, pred in zip(GroundTruth, predictedIndices):
            if len(gt) == 0:
                continue
            users += 1
            gts = set(gt)
            hits = [j for j in range(n) if pred[j] in gts]
            dcg = sum(disc[j] for j in hits)
            idcg = sum(disc[: min(n, len(gt))])
            sp += len(hits) / n
            sr += len(hits) / len(gt)
            sn += dcg / idcg if idcg != 0 else 0.0
            sm += 1.0 / (hits[0] + 1.0) if hits else 0.0
        precision.append(round(sp / users, 4))
        recall.append(round(sr / users, 4))
        ndcg.append(round(sn / users, 4))
        mrr.append(round(sm / users, 4))
    return precision, recall, ndcg, mrr
