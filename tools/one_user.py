"""One warm-up user + `--users` users through atspeed_bssd on the bench shapes: the short command ncu wraps
(launch list / --set full captures).  Never a bench value."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from atspeed_b200.constraint import compile_constraint  # noqa: E402
from atspeed_b200.engine import DeviceModel, DeviceTrie, ModelSpec, Session  # noqa: E402
from atspeed_b200.prompts import load_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=1)
ap.add_argument("--target", default="7b")
ap.add_argument("--draft", default="68m")
ap.add_argument("--K", type=int, default=10)
ap.add_argument("--N", type=int, default=40)
ap.add_argument("--gamma", type=int, default=3)
ap.add_argument("--dataset", default="beauty")
ap.add_argument("--constraint", default="strict")
ap.add_argument("--cohort", type=int, default=1, help="> 1: run --users users through atspeed_bssd_batch")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ds = load_dataset(a.dataset)
fn = bench.make_fn(ds, a.constraint)
specs = []
for name in (a.target, a.draft):
    s = bench.SHAPES[name]
    specs.append(ModelSpec(ds.vocab_size, s["hidden"], s["n_layers"], s["n_heads"], s["hidden"] // s["n_heads"], s["mlp"]))
tdm = DeviceModel(specs[0], bench.gpu_weights(specs[0], 1, dev), dev)
ddm = DeviceModel(specs[1], bench.gpu_weights(specs[1], 2, dev), dev)
csr = compile_constraint(fn, ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(1))
if a.cohort > 1:
    sess = Session(tdm, ddm, DeviceTrie(csr, dev), a.K, a.N, 4, max_users=a.cohort)
    for rep in range(2):       # first pass = warm-up
        outs = sess.bssd_batch([ds.prompt_ids(u) for u in range(a.users)], a.gamma)
    out = outs[-1]
else:
    sess = Session(tdm, ddm, DeviceTrie(csr, dev), a.K, a.N, 4)
    for u in range(a.users + 1):
        out = sess.bssd(ds.prompt_ids(u), a.gamma)
torch.cuda.synchronize()
print("ok", out["n_run"], out["accept_steps"], out["kernel_launches"])
