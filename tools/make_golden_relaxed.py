"""Generate tests/golden/bssd_relaxed.json: the UNMODIFIED reference's sampling branch (AtSpeed-R relaxed acceptance,
/root/reference/code/beamSD.py:65-74,293-369 with do_sample=True) under fixed torch seeds.

Model stack: oracle.llama_ref.RefLlama behind HFStyleProxy (the proxy restates transformers 4.41's
`_get_logits_processor` / `_get_logits_warper`, which transformers 5.5 no longer offers outside `generate`;
SURVEY 8c shim 4).  `top_k=50`, temperature 1.0 and 0.7.  The reference draws from torch's global generator, so a
case is (seed -> outputs); tests/test_oracle_relaxed.py replays each seed through oracle/bssd_ref.py in generator mode.
Cases where the reference raises (torch.multinomial over an all-zero / too-sparse distribution) are kept
with "raises": true -- they document where the reference is undefined."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as MG            # noqa: E402  (imports the reference, installs the shims)
from transformers import GenerationConfig   # noqa: E402
from oracle import llama_ref as LR  # noqa: E402
from atspeed_b200.prompts import load_dataset, RESPONSE_SEP, BOS_ID, EOS_ID  # noqa: E402
from atspeed_b200.generation_trie import suffix_prefix_allowed_tokens_fn      # noqa: E402
import re                            # noqa: E402


def run(pt, pd_, prompt, K, N, gamma, fn, seed, temperature):
    ids = torch.tensor([prompt])
    for m, nb in ((pt, K), (pd_, N)):
        m.generation_config.num_beams = nb
        m.generation_config.max_new_tokens = 4
        m.generation_config.do_sample = True
        m.generation_config.temperature = temperature
        m.generation_config.top_k = 50
    torch.manual_seed(seed)
    with MG.Recorder() as rec:
        out = MG.ref_beamSD.BSSD(pt, pd_, {"input_ids": ids}, gamma, 4, prefix_allowed_tokens_fn=fn)
    return out, rec.rounds


def main():
    t0 = time.time()
    cases = []
    tok_re = re.compile(r"<[abcd]_\d+>")
    for name in ("beauty", "games"):
        ds = load_dataset(name)
        rds = MG.reference_dataset(name)
        enc = lambda s: [BOS_ID] + [ds.token_id[t] for t in tok_re.findall(s)]
        ref_trie = MG.RefTrie([enc(e) + [EOS_ID] for e in rds.get_all_items()])
        strict_fn = suffix_prefix_allowed_tokens_fn(ref_trie, RESPONSE_SEP)
        pos_fn = rds.get_prefix_allowed_tokens_fn(MG.FakeTokenizer(ds))
        V = ds.vocab_size
        sh_t, sh_d = LR.shape_small_target(V), LR.shape_small_draft(V)
        Wt = LR.make_weights(sh_t, 10, std=1.28 / 16, dtype=torch.bfloat16)
        Wd_ind = LR.make_weights(sh_d, 11, std=1.28 / (128 ** 0.5), dtype=torch.bfloat16)
        sh_dc = LR.LlamaShape(V, sh_t.hidden, 1, sh_t.n_heads, sh_t.mlp)
        Wd_cor = LR.make_weights(sh_dc, 12, std=1.28 / 16, dtype=torch.bfloat16, like=Wt, noise=0.03)
        gc = lambda: GenerationConfig(num_beams=1, max_new_tokens=4, do_sample=True, top_k=50)
        for dname, shd, Wd in (("correlated", sh_dc, Wd_cor), ("independent", sh_d, Wd_ind)):
            pt = LR.HFStyleProxy(LR.RefLlama(sh_t, Wt, "bf16"), gc())
            pd_ = LR.HFStyleProxy(LR.RefLlama(shd, Wd, "bf16"), gc())
            for cname, fn in (("positional", pos_fn), ("strict", strict_fn)):
                grid = [(10, 40, 3, 1.0), (20, 40, 3, 1.0), (5, 10, 2, 1.0), (10, 40, 3, 0.7), (1, 40, 3, 1.0)]
                for (K, N, gamma, temp) in grid:
                    for u in (0, 1, 17)[: (3 if (K, N, gamma, temp) == (10, 40, 3, 1.0) else 1)]:
                        for seed in (2025, 7):
                            prompt = ds.prompt_ids(u)
                            base = {"stack": "ref_bf16", "dataset": name, "user": u, "draft": dname, "constraint": cname,
                                    "K": K, "N": N, "gamma": gamma, "temperature": temp, "top_k": 50, "seed": seed}
                            try:
                                out, rounds = run(pt, pd_, prompt, K, N, gamma, fn, seed, temp)
                            except Exception as e:   # the reference is undefined here
                                cases.append(dict(base, raises=True, error=repr(e)[:120]))
                                continue
                            P = len(prompt)
                            cases.append(dict(base, raises=False, bssd=MG.pack(out, P), n_run=out["n_run"],
                                              total_accept_steps=out["total_accept_steps"],
                                              ave_accept_tokens=out["ave_accept_tokens"], rounds=rounds))
        print(name, len(cases), f"{time.time() - t0:.0f}s", flush=True)
    json.dump({"cases": cases}, open(os.path.join(MG.OUT, "bssd_relaxed.json"), "w"))
    ok = [c for c in cases if not c["raises"]]
    print("cases", len(cases), "raised", len(cases) - len(ok), "accept histogram",
          torch.bincount(torch.tensor([c["total_accept_steps"] for c in ok])).tolist(), f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
