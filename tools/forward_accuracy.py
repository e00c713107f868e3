#!/usr/bin/env python
"""Accuracy of the bf16 CUDA forward at the benchmark LAYER shape against an fp32 yardstick: last-position prompt logits from
(1) this repo's kernels, (2) the CPU oracle in its bf16 contract, (3) the oracle in fp32, (4) HF LlamaForCausalLM bf16 on the
GPU, (5) the same HF module in fp32 -- all on the same random-init weights.  Prints pairwise mean |delta logit|.
usage: python tools/forward_accuracy.py [--layers 4] [--prompts 3]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from atspeed_b200.constraint import compile_constraint  # noqa: E402
from atspeed_b200.engine import DeviceModel, DeviceTrie, ModelSpec, Session  # noqa: E402
from atspeed_b200.prompts import load_dataset  # noqa: E402
from oracle import llama_ref as LR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--prompts", type=int, default=3)
ap.add_argument("--shape", default="7b")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ds = load_dataset("beauty")
V = ds.vocab_size
s = bench.SHAPES[a.shape]
spec = ModelSpec(V, s["hidden"], a.layers, s["n_heads"], s["hidden"] // s["n_heads"], s["mlp"])
dm = DeviceModel(spec, bench.gpu_weights(spec, 1, dev), dev)
W = {"embed": dm.embed, "norm": dm.norm, "lm_head": dm.lm_head, "layers": dm.layers}
fn = bench.make_fn(ds, "positional")
sess = Session(dm, None, DeviceTrie(compile_constraint(fn, ds.prompt_ids(0), 4), dev), K=10, N=10, max_new_tokens=4, max_prompt=400)
sh = LR.LlamaShape(V, spec.hidden, a.layers, spec.n_heads, spec.mlp, spec.head_dim, spec.rope_theta, spec.eps)
torch.set_num_threads(len(os.sched_getaffinity(0)))
o16, o32 = LR.RefLlama(sh, W, "bf16"), LR.RefLlama(sh, W, "fp32")
from transformers import LlamaConfig, LlamaForCausalLM
cfg = LlamaConfig(vocab_size=V, hidden_size=spec.hidden, intermediate_size=spec.mlp, num_hidden_layers=a.layers,
                  num_attention_heads=spec.n_heads, num_key_value_heads=spec.n_heads, tie_word_embeddings=False)
with torch.device(dev):
    hf = LlamaForCausalLM(cfg)
inv_freq = hf.model.rotary_emb.inv_freq.detach().clone().float()      # .to(bfloat16) must not round the rotary frequencies
hf = hf.to(torch.bfloat16).eval()
hf.model.rotary_emb.inv_freq = inv_freq
hf.load_state_dict(LR.weights_to_hf_state_dict(W), strict=False)
res = {}
rows = []
for u in range(a.prompts):
    prompt = ds.prompt_ids(u * 37)
    P = len(prompt)
    i32 = lambda x: torch.tensor(list(x), dtype=torch.int32, device=dev)
    ours = sess.forward_raw(0, i32(prompt), i32(range(P)), i32(range(P)), i32(range(1, P + 1)),
                            torch.zeros(P, 16, dtype=torch.int32, device=dev), P, P, i32([P - 1]))[0]
    vis = torch.tril(torch.ones(P, P, dtype=torch.bool))
    c16 = o16.forward(torch.tensor(prompt), torch.arange(P), vis, LR.RefCache(), torch.tensor([P - 1])).numpy()[0]
    c32 = o32.forward(torch.tensor(prompt), torch.arange(P), vis, LR.RefCache(), torch.tensor([P - 1])).numpy()[0]
    with torch.no_grad():
        h16 = hf(input_ids=torch.tensor([prompt], device=dev)).logits[0, -1].float().cpu().numpy()
    rows.append(dict(ours=ours, oracle_bf16=c16, oracle_fp32=c32, hf_bf16=h16, prompt=prompt))
hf = hf.float()
hf.model.rotary_emb.inv_freq = inv_freq
for r in rows:
    with torch.no_grad():
        r["hf_fp32"] = hf(input_ids=torch.tensor([r["prompt"]], device=dev)).logits[0, -1].float().cpu().numpy()
names = ["ours", "oracle_bf16", "hf_bf16", "oracle_fp32", "hf_fp32"]
print(f"# {a.shape} layer shape, {a.layers} layers, {a.prompts} prompts; logit std {np.mean([r['hf_fp32'].std() for r in rows]):.3f}; mean |delta logit|:")
print(" " * 12 + "".join(f"{n:>13}" for n in names))
for x in names:
    print(f"{x:>12}" + "".join(f"{np.mean([np.abs(r[x] - r[y]).mean() for r in rows]):13.5f}" for y in names))
