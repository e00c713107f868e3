"""Micro-benchmark of the tcgen05 GEMM (C ABI atspeed_gemm_bf16) on the 7B / 68M projection shapes.
Weights are cycled through distinct buffers (> L2) so every launch streams from HBM. Run on the GPU box."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atspeed_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6536.4
SHAPES = {"qkv": (4096, (4096, 4096, 4096)), "o": (4096, (4096,)), "gate_up": (4096, (11008, 11008)),
          "down": (11008, (4096,)), "lm_head": (4096, (32859,))}
if len(sys.argv) > 1 and sys.argv[1] == "68m":
    SHAPES = {"qkv": (768, (768, 768, 768)), "o": (768, (768,)), "gate_up": (768, (3072, 3072)), "down": (3072, (768,)),
              "lm_head": (768, (32859,))}
NBUF = 6


st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
print(f"{'gemm':<8} {'T':>4} {'us':>8} {'GB/s':>8} {'frac':>6}   (algorithmic bytes = W + X + bf16 Y)")
for name, (K, rows) in SHAPES.items():
    ws = [[(torch.randn(r, K, device=dev) * 0.02).to(torch.bfloat16) for r in rows] for _ in range(NBUF)]
    for T in (10, 50, 90, 130, 220, 289):
        x = (torch.randn(T, K, device=dev) * 0.5).to(torch.bfloat16)
        ldo = sum(rows)
        r3 = list(rows) + [0] * (3 - len(rows))
        nb = C.c_size_t(0)
        assert lib.atspeed_gemm_scratch_bytes(T, K, r3[0], r3[1], r3[2], C.byref(nb)) == 0
        out = torch.empty(nb.value // 4, device=dev, dtype=torch.float32)
        spl = nb.value // 4 // (T * ldo)

        def run(i):
            w = ws[i % NBUF]
            p = [t.data_ptr() for t in w] + [None] * (3 - len(w))
            r = list(rows) + [0] * (3 - len(rows))
            rc = lib.atspeed_gemm_bf16(x.data_ptr(), T, K, p[0], r[0], p[1], r[1], p[2], r[2], out.data_ptr(), None, ldo, st)
            assert rc == 0, lib.atspeed_last_error()

        for i in range(NBUF):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        torch.cuda._sleep(int(6e6))          # ~3 ms of GPU spin: the launches below queue up, so GPU time is measured
        e0.record()
        for i in range(n):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        byts = 2.0 * (sum(rows) * K + T * K + T * sum(rows))
        gbs = byts / us / 1e3
        print(f"{name:<8} {T:>4} {us:>8.1f} {gbs:>8.0f} {gbs / peak:>6.2f}  max_slices={spl}")
