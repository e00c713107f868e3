#!/usr/bin/env python
"""Phase timing of the tree-attention kernel (diagnostic build: ATSPEED_NVCC_DEFS=-DATT_TIMING python -m atspeed_b200.build
--force): clock64 stamps of thread 0 of CTA (0,0) at the phase boundaries, printed as microseconds at the SM clock the run saw.
usage: python tools/att_timing.py [--T 289] [--P 100] [--tree 240]"""
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
ap.add_argument("--T", type=int, default=289)
ap.add_argument("--P", type=int, default=100)
ap.add_argument("--tree", type=int, default=240)
ap.add_argument("--mhz", type=float, default=1900.0)
a = ap.parse_args()
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else lib
dev = torch.device("cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
H, D, P, S, T = 32, 128, a.P, a.P + a.tree, a.T
HD = H * D
NBUF = 40
rng = np.random.default_rng(0)
q = (torch.randn(T, HD, device=dev) * 0.5).to(torch.bfloat16)
kv = [((torch.randn(S, HD, device=dev) * 0.5).to(torch.bfloat16), (torch.randn(S, HD, device=dev) * 0.5).to(torch.bfloat16)) for _ in range(NBUF)]
pl = torch.full((T,), P, dtype=torch.int32, device=dev)
vis = np.zeros((T, 16), dtype=np.uint32)
for t in range(T):
    for j in rng.integers(0, a.tree, 4).tolist() + [t % a.tree]:
        vis[t, j >> 5] |= np.uint32(1) << np.uint32(j & 31)
visd = torch.from_numpy(vis.view(np.int32)).to(dev)
out = torch.empty(T, HD, device=dev, dtype=torch.bfloat16)
names = ["entry", "pdl_wait done", "Q/vis staged (sync)", "lists+prefetch issued", "tile0 start", "tile1 start", "tile2 start", "tile3 start",
         "dense done", "state handed off", "sparse done", "stored"]
for rep in range(6):
    k, v = kv[rep % NBUF]
    rc = lib.atspeed_tree_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), pl.data_ptr(), visd.data_ptr(), P, T, S, H, D, out.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    fn = lib.atspeed_debug_att_stamps if hasattr(lib, "atspeed_debug_att_stamps") else C.CDLL(os.path.join(ROOT, "atspeed_b200", "libatspeed_b200.so")).atspeed_debug_att_stamps
    assert fn(buf) == 0
    t = list(buf)[:12]
    n_tiles = (P + 63) // 64
    line = []
    prev = t[0]
    for i, nm in enumerate(names):
        if 4 <= i <= 7 and i - 4 >= n_tiles:
            continue
        line.append("%s +%.2f" % (nm, (t[i] - prev) / a.mhz))
        prev = t[i]
    print("rep %d total %.2f us | %s" % (rep, (t[11] - t[0]) / a.mhz, " | ".join(line)), flush=True)
