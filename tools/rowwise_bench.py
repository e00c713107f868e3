#!/usr/bin/env python
"""Micro-benchmark of the row-wise consumer kernels between the GEMMs (C ABI atspeed_debug_rowwise_us) at cohort-forward sizes:
microseconds per launch and GB/s of the bytes the kernel must move (slices x fp32 partial sums in, bf16 out) against the
measured copy peak.  usage: python tools/rowwise_bench.py [--T 289,400,480] [--iters 40]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atspeed_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", default="289,400,480")
ap.add_argument("--iters", type=int, default=40)
a = ap.parse_args()
lib = _lib.load()
torch.cuda.init()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
H, MLP, NH = 4096, 11008, 32
print(f"# hidden {H} mlp {MLP} heads {NH}; copy peak {peak:.0f} GB/s")
print(f"{'kernel':<18} {'T':>4} {'slices':>6} {'us':>7} {'MB':>7} {'GB/s':>7} {'of peak':>7}")
for T in [int(x) for x in a.T.split(",")]:
    for kind, name, slices_list in ((0, "qkv_rope_append", (1, 2, 3)), (1, "silu_mul", (1, 2)), (2, "residual_rmsnorm", (1, 4, 5))):
        for sl in slices_list:
            us = C.c_float(0)
            rc = lib.atspeed_debug_rowwise_us(kind, T, H, MLP, NH, sl, a.iters, C.byref(us), st)
            assert rc == 0, lib.atspeed_last_error()
            cols = (3 * H, 2 * MLP, H)[kind]
            out_b = (3 * H * 2, MLP * 2, H * 2 * 3)[kind]          # bf16 written (+ residual read/write for kind 2)
            mb = T * (cols * 4 * sl + out_b) / 1e6
            print(f"{name:<18} {T:>4} {sl:>6} {us.value:>7.1f} {mb:>7.1f} {mb / us.value * 1e3:>7.0f} {mb / us.value * 1e3 / peak:>7.2f}", flush=True)
