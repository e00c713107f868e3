#!/usr/bin/env python
"""Stand-alone roofline measurements of kernel (a) (mask + log-softmax + top-B) and kernel (c) (KV row gather) through
the C ABI, at the sizes SURVEY.md section 8(d) names: achieved GB/s = algorithmic bytes / CUDA-event time, against the
measured copy peak in MEASURED_PEAKS.json.

  (a) algorithmic bytes = rows x V x sizeof(logit)              rows in {1, 81, 121, 1024, 8192}, fp32 and bf16 logits
  (c) algorithmic bytes = 2 x rows x planes x row_bytes         7B cache: 64 planes x 8 KiB rows; 68M cache: 4 x 1.5 KiB

Inputs are cycled through buffers whose total size exceeds L2 (126 MB), so every timed launch reads HBM.
usage:  python tools/kernel_abc_bench.py [--iters 20] [--json out.json]      (needs a B200; run under gpurun)
ncu:    ncu --set full --clock-control none -k regex:"mask_logsoftmax|kv_gather" -c 12 -o gpurun_out/prof_abc python tools/kernel_abc_bench.py --iters 1 --quick
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

L2_BYTES = 126 << 20


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def timed(fn, n_bufs, iters):
    """Mean microseconds per call of fn(i) (i = buffer index), CUDA events on the current stream, after warm-up."""
    for i in range(min(3, max(1, n_bufs))):
        fn(i % n_bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_bufs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def bench_a(lib, stream, rows_list, iters, peak):
    from atspeed_b200.constraint import compile_constraint
    from atspeed_b200.engine import DeviceTrie
    from atspeed_b200.prompts import load_dataset
    from bench import make_fn
    ds = load_dataset("beauty")
    csr = compile_constraint(make_fn(ds, "strict"), ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(3))
    trie = DeviceTrie(csr, torch.device("cuda"))
    V, B = ds.vocab_size, 40
    ldl = (V + 7) & ~7                                  # the engine's logits row stride (16-byte aligned rows)
    out = []
    rng = np.random.default_rng(0)
    for dtype, code, esz in (("fp32", 0, 4), ("bf16", 1, 2)):
        for rows in rows_list:
            nbytes = rows * V * esz
            n_bufs = int(min(64, max(1, -(-2 * L2_BYTES // nbytes))))
            tdt = torch.float32 if code == 0 else torch.bfloat16
            bufs = [(torch.randn(rows, ldl, device="cuda") * 2.0).to(tdt) for _ in range(n_bufs)]
            # realistic node mix: the root (91 children), depth-1 nodes (mean 72 children), deeper nodes (1-2 children)
            node = torch.from_numpy(rng.integers(0, csr.n_nodes, rows).astype(np.int32)).cuda()
            node[0] = 0
            ct = torch.zeros(rows * B, dtype=torch.int32, device="cuda")
            ce = torch.zeros_like(ct)
            cl = torch.zeros(rows * B, dtype=torch.float32, device="cuda")
            cc = torch.zeros(rows, dtype=torch.int32, device="cuda")
            lse = torch.zeros(rows, dtype=torch.float32, device="cuda")

            def call(i):
                rc = lib.atspeed_mask_logsoftmax_topk(bufs[i].data_ptr(), code, rows, V, ldl, node.data_ptr(), None,
                                                      C.byref(trie.desc), B, ct.data_ptr(), ce.data_ptr(), cl.data_ptr(),
                                                      cc.data_ptr(), lse.data_ptr(), stream)
                assert rc == 0, lib.atspeed_last_error().decode()

            us = timed(call, n_bufs, iters)
            gbs = nbytes / us / 1e3
            out.append({"kernel": "a:mask_logsoftmax_topk", "logits": dtype, "rows": rows, "V": V, "us": us,
                        "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_peak": gbs / peak, "buffers_cycled": n_bufs})
            print(f"(a) {dtype} rows={rows:5d}  {us:9.1f} us  {nbytes / 1e6:9.2f} MB  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of peak")
            del bufs
            torch.cuda.empty_cache()
    return out


def bench_c(lib, stream, iters, peak, quick):
    out = []
    cases = [("7b", 64, 8192, 700), ("68m", 4, 1536, 700)]
    for name, planes, row_bytes, S in cases:
        for rows in ((40,) if quick else (10, 40, 120)):
            nbytes = 2 * rows * planes * row_bytes
            n_bufs = int(min(16, max(2, -(-2 * L2_BYTES // (planes * S * row_bytes)))))
            bufs = [torch.randint(0, 255, (planes, S, row_bytes), dtype=torch.uint8, device="cuda") for _ in range(n_bufs)]
            perm = torch.randperm(S)
            src, dst = perm[:rows].to(torch.int32).cuda(), perm[rows:2 * rows].to(torch.int32).cuda()   # disjoint
            n_dev = torch.tensor([rows], dtype=torch.int32, device="cuda")

            def call(i):
                b = bufs[i]
                rc = lib.atspeed_kv_gather(b.data_ptr(), b.data_ptr(), S * row_bytes, S * row_bytes, planes, row_bytes,
                                           src.data_ptr(), dst.data_ptr(), n_dev.data_ptr(), rows, stream)
                assert rc == 0, lib.atspeed_last_error().decode()

            us = timed(call, n_bufs, iters)
            gbs = nbytes / us / 1e3
            out.append({"kernel": "c:kv_gather", "cache": name, "rows": rows, "planes": planes, "row_bytes": row_bytes, "us": us,
                        "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_peak": gbs / peak, "buffers_cycled": n_bufs})
            print(f"(c) {name:>3} rows={rows:4d} planes={planes:3d}  {us:9.1f} us  {nbytes / 1e6:9.2f} MB  {gbs:7.0f} GB/s  "
                  f"{gbs / peak:5.2f} of peak")
            del bufs
            torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="fewer sizes (for an ncu capture)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from atspeed_b200 import _lib
    lib = _lib.load()
    assert torch.cuda.is_available(), "needs a GPU"
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    peak, src = peak_gbs()
    print(f"# peak {peak:.0f} GB/s ({src}); inputs cycled through > 2 x L2 of buffers")
    res = bench_a(lib, stream, (121, 1024) if a.quick else (1, 81, 121, 1024, 8192), a.iters, peak)
    res += bench_c(lib, stream, a.iters, peak, a.quick)
    if a.json:
        json.dump({"peak_gbs": peak, "peak_source": src, "results": res}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
