#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM (C ABI atspeed_gemm_bf16) on the 7B / 68M projection shapes: microseconds, algorithmic
GB/s (W + X + bf16 Y) and TFLOP/s per launch against MEASURED_PEAKS.json.  Weights are cycled through distinct buffers (> L2)
so every launch streams them from HBM.  Run on the GPU box; `--iters 1 --T 512` is the command ncu wraps for the --set full
capture of the cohort-forward GEMMs (T is then known, unlike inside a search).

usage: python tools/gemm_bench.py [--model 7b|68m] [--T 10,50,130,220,289,400,512] [--iters 30] [--json out.json]
       ATSPEED_GEMM_2CTA=0 keeps the single-CTA kernel for T > 256."""
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
ap.add_argument("--model", default="7b")
ap.add_argument("--T", default="10,50,130,220,289,400,512")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--json", default=None)
a = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
SHAPES = {"7b": {"qkv": (4096, (4096, 4096, 4096)), "o": (4096, (4096,)), "gate_up": (4096, (11008, 11008)),
                 "down": (11008, (4096,)), "lm_head": (4096, (32859,))},
          "68m": {"qkv": (768, (768, 768, 768)), "o": (768, (768,)), "gate_up": (768, (3072, 3072)), "down": (3072, (768,)),
                  "lm_head": (768, (32859,))}}[a.model]
NBUF = 6
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
res = []
print(f"# pair kernel (T > 256): {os.environ.get('ATSPEED_GEMM_2CTA', '1') != '0'}; peaks {peaks['hbm_gbs']:.0f} GB/s, "
      f"{peaks['bf16_tflops_sustained']:.0f} TFLOP/s sustained")
print(f"{'gemm':<8} {'T':>4} {'us':>8} {'GB/s':>8} {'hbm':>5} {'TFLOP/s':>8} {'tensor':>6}  slices")
for name, (K, rows) in SHAPES.items():
    ws = [[(torch.randn(r, K, device=dev) * 0.02).to(torch.bfloat16) for r in rows] for _ in range(NBUF)]
    for T in [int(x) for x in a.T.split(",")]:
        x = (torch.randn(T, K, device=dev) * 0.5).to(torch.bfloat16)
        ldo = sum(rows)
        r3 = list(rows) + [0] * (3 - len(rows))
        nb = C.c_size_t(0)
        assert lib.atspeed_gemm_scratch_bytes(T, K, r3[0], r3[1], r3[2], C.byref(nb)) == 0
        out = torch.empty(nb.value // 4, device=dev, dtype=torch.float32)

        def run(i):
            w = ws[i % NBUF]
            p = [t.data_ptr() for t in w] + [None] * (3 - len(w))
            rc = lib.atspeed_gemm_bf16(x.data_ptr(), T, K, p[0], r3[0], p[1], r3[1], p[2], r3[2], out.data_ptr(), None, ldo, st)
            assert rc == 0, lib.atspeed_last_error()

        for i in range(min(NBUF, max(1, a.iters))):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(6e6))          # ~3 ms of GPU spin: the launches below queue up, so GPU time is measured
        e0.record()
        for i in range(a.iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        byts = 2.0 * (sum(rows) * K + T * K + T * sum(rows))
        fl = 2.0 * sum(rows) * K * T
        gbs, tf = byts / us / 1e3, fl / us / 1e6
        res.append({"gemm": name, "T": T, "us": us, "GB/s": gbs, "hbm_frac": gbs / peaks["hbm_gbs"], "TFLOP/s": tf,
                    "tensor_frac": tf / peaks["bf16_tflops_sustained"], "slices": nb.value // 4 // (T * ldo)})
        print(f"{name:<8} {T:>4} {us:>8.1f} {gbs:>8.0f} {gbs / peaks['hbm_gbs']:>5.2f} {tf:>8.0f} {tf / peaks['bf16_tflops_sustained']:>6.2f}  "
              f"{nb.value // 4 // (T * ldo)}", flush=True)
    del ws
    torch.cuda.empty_cache()
if a.json:
    json.dump({"model": a.model, "pair_kernel": os.environ.get("ATSPEED_GEMM_2CTA", "1") != "0", "peaks": peaks, "results": res},
              open(a.json, "w"), indent=1)
