#!/usr/bin/env python
"""Micro-benchmark of the tree-attention kernel (C ABI atspeed_tree_attention) at cohort-forward sizes: T queries, each
seeing a causal prompt prefix plus a random subset of the tree slots, 32 heads x 128 (7B) -- microseconds per layer-launch.
K/V are cycled through buffers > L2.  ATSPEED_ATT_BQ / ATSPEED_ATT_PLO select the variants (read once per process).
usage: python tools/att_bench.py [--T 289,512] [--P 100] [--tree 240] [--iters 50]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atspeed_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", default="130,289,512")
ap.add_argument("--P", type=int, default=100)
ap.add_argument("--tree", type=int, default=240, help="tree/accepted slots after the prompt (<= 512)")
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--D", type=int, default=128)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
H, D, P, S = a.heads, a.D, a.P, a.P + a.tree
HD = H * D
NBUF = max(2, int(2 * (126 << 20) / (2 * S * HD * 2)) + 1)
rng = np.random.default_rng(0)
print(f"# BQ={os.environ.get('ATSPEED_ATT_BQ', '64')} PLO={os.environ.get('ATSPEED_ATT_PLO', '1')} H={H} D={D} P={P} S={S} kv buffers={NBUF}")
for T in [int(x) for x in a.T.split(",")]:
    q = (torch.randn(T, HD, device=dev) * 0.5).to(torch.bfloat16)
    kv = [((torch.randn(S, HD, device=dev) * 0.5).to(torch.bfloat16), (torch.randn(S, HD, device=dev) * 0.5).to(torch.bfloat16))
          for _ in range(NBUF)]
    pl = torch.full((T,), P, dtype=torch.int32, device=dev)
    vis = np.zeros((T, 16), dtype=np.uint32)
    for t in range(T):                                     # ~4 visible tree slots per token (its ancestor chain) + itself
        for j in rng.integers(0, a.tree, 4).tolist() + [t % a.tree]:
            vis[t, j >> 5] |= np.uint32(1) << np.uint32(j & 31)
    visd = torch.from_numpy(vis.view(np.int32)).to(dev)
    out = torch.empty(T, HD, device=dev, dtype=torch.bfloat16)

    def run(i):
        k, v = kv[i % NBUF]
        rc = lib.atspeed_tree_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), pl.data_ptr(), visd.data_ptr(), P, T, S, H, D,
                                        out.data_ptr(), st)
        assert rc == 0, lib.atspeed_last_error()

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e6))
    e0.record()
    for i in range(a.iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.iters
    kvb = 2.0 * S * HD * 2 + 2.0 * T * HD * 2
    print(f"T={T:4d}  {us:8.1f} us/launch   K/V+Q/O bytes {kvb / 1e6:6.2f} MB -> {kvb / us / 1e3:6.0f} GB/s", flush=True)
