"""GEMM experiment sweep (GPU box): plan overrides through env vars, a few shapes, one table."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atspeed_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
SHAPES = {"qkv": (4096, (4096, 4096, 4096)), "o": (4096, (4096,)), "gate_up": (4096, (11008, 11008)),
          "down": (11008, (4096,)), "lm_head": (4096, (32859,))}
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
NBUF = 6


def bench(name, T, env):
    for k in ("ATSPEED_GEMM_2CTA", "ATSPEED_GEMM_2CTA_MIN", "ATSPEED_GEMM_BM", "ATSPEED_GEMM_STAGES", "ATSPEED_GEMM_CTAS", "ATSPEED_PDL", "ATSPEED_GEMM_BUFS", "ATSPEED_GEMM_BBOX", "ATSPEED_GEMM_PACKED", "ATSPEED_GEMM_DBG", "ATSPEED_GEMM_CHAINS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    K, rows = SHAPES[name]
    ws = WS[name]
    x = (torch.randn(T, K, device=dev) * 0.5).to(torch.bfloat16)
    r3 = list(rows) + [0] * (3 - len(rows))
    nb = C.c_size_t(0)
    assert lib.atspeed_gemm_scratch_bytes(T, K, r3[0], r3[1], r3[2], C.byref(nb)) == 0
    out = torch.empty(nb.value // 4, device=dev, dtype=torch.float32)

    def run(i):
        w = ws[i % NBUF]
        p = [t.data_ptr() for t in w] + [None] * (3 - len(w))
        rc = lib.atspeed_gemm_bf16(x.data_ptr(), T, K, p[0], r3[0], p[1], r3[1], p[2], r3[2], out.data_ptr(), None, sum(rows), st)
        assert rc == 0, lib.atspeed_last_error()

    for i in range(NBUF):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    torch.cuda._sleep(int(6e6))
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    byts = 2.0 * (sum(rows) * K + T * K + T * sum(rows))
    return us, byts / us / 1e3


WS = {n: [[(torch.randn(r, K, device=dev) * 0.02).to(torch.bfloat16) for r in rows] for _ in range(NBUF)]
      for n, (K, rows) in SHAPES.items()}
TS = (10, 50, 130, 220)
CONFIGS = [("default", {}), ("1 tmem buf", {"ATSPEED_GEMM_BUFS": "1"}), ("bm128", {"ATSPEED_GEMM_BM": "128"}),
           ("bm256", {"ATSPEED_GEMM_BM": "256"}), ("bm128 x132", {"ATSPEED_GEMM_BM": "128", "ATSPEED_GEMM_CTAS": "132"}),
           ("bm128 2/SM", {"ATSPEED_GEMM_BM": "128", "ATSPEED_GEMM_STAGES": "2", "ATSPEED_GEMM_CTAS": "296"})]
if len(sys.argv) > 1 and sys.argv[1] == "bigT":
    CONFIGS = [("default", {}), ("no epi stores", {"ATSPEED_GEMM_DBG": "1"}), ("no mma", {"ATSPEED_GEMM_DBG": "2"}),
               ("neither", {"ATSPEED_GEMM_DBG": "3"})]
    TS = (256, 300, 400, 512)
elif len(sys.argv) > 1 and sys.argv[1] == "2cta":
    CONFIGS = [("2cta>256", {}), ("2cta>128", {"ATSPEED_GEMM_2CTA_MIN": "128"}), ("2cta>64", {"ATSPEED_GEMM_2CTA_MIN": "64"}),
               ("2cta>0", {"ATSPEED_GEMM_2CTA_MIN": "0"})]
    TS = (10, 50, 90, 130, 220, 256)
elif len(sys.argv) > 1 and sys.argv[1] == "chains":
    CONFIGS = [("default", {}), ("c1 b2", {"ATSPEED_GEMM_CHAINS": "1", "ATSPEED_GEMM_BUFS": "2"}),
               ("c2 b1", {"ATSPEED_GEMM_CHAINS": "2", "ATSPEED_GEMM_BUFS": "1"}),
               ("c2 b2", {"ATSPEED_GEMM_CHAINS": "2", "ATSPEED_GEMM_BUFS": "2"}),
               ("c4 b1", {"ATSPEED_GEMM_CHAINS": "4", "ATSPEED_GEMM_BUFS": "1"}),
               ("bm128 c2 b2", {"ATSPEED_GEMM_BM": "128", "ATSPEED_GEMM_CHAINS": "2", "ATSPEED_GEMM_BUFS": "2"}),
               ("bm128 c4 b1", {"ATSPEED_GEMM_BM": "128", "ATSPEED_GEMM_CHAINS": "4", "ATSPEED_GEMM_BUFS": "1"}),
               ("bm128 c4 b2", {"ATSPEED_GEMM_BM": "128", "ATSPEED_GEMM_CHAINS": "4", "ATSPEED_GEMM_BUFS": "2"}),
               ("bm256 c2 b1", {"ATSPEED_GEMM_BM": "256", "ATSPEED_GEMM_CHAINS": "2", "ATSPEED_GEMM_BUFS": "1"})]
    TS = (10, 50, 90, 130, 220)
elif len(sys.argv) > 1 and sys.argv[1] == "dbg":
    CONFIGS = [("default", {}), ("no epi stores", {"ATSPEED_GEMM_DBG": "1"}), ("no mma", {"ATSPEED_GEMM_DBG": "2"}),
               ("neither", {"ATSPEED_GEMM_DBG": "3"}), ("neither bm256", {"ATSPEED_GEMM_DBG": "3", "ATSPEED_GEMM_BM": "256"}),
               ("neither bm128", {"ATSPEED_GEMM_DBG": "3", "ATSPEED_GEMM_BM": "128"})]
elif len(sys.argv) > 1 and sys.argv[1] == "packed":
    CONFIGS = [("default", {}), ("packed", {"ATSPEED_GEMM_PACKED": "1"}), ("packed bm128", {"ATSPEED_GEMM_PACKED": "1", "ATSPEED_GEMM_BM": "128"}),
               ("packed bm256", {"ATSPEED_GEMM_PACKED": "1", "ATSPEED_GEMM_BM": "256"})]
elif len(sys.argv) > 1 and sys.argv[1] == "bbox":
    CONFIGS = [("default", {}), ("bbox112", {"ATSPEED_GEMM_BBOX": "112"}), ("bbox64", {"ATSPEED_GEMM_BBOX": "64"}),
               ("bbox16", {"ATSPEED_GEMM_BBOX": "16"})]
    TS = (130, 220)
print(f"{'config':<14}" + "".join(f"{n[:7] + ' ' + str(T):>13}" for n in SHAPES for T in TS))
for cname, env in CONFIGS:
    row = f"{cname:<14}"
    for n in SHAPES:
        for T in TS:
            try:
                us, gbs = bench(n, T, env)
                row += f"{us:>7.1f}/{gbs / 1e3:>4.2f} "
            except Exception as e:
                row += f"{'err':>13}"
    print(row, flush=True)
