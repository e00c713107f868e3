"""More users for the oracle pin (SURVEY 8c asks for 64 Beauty + 64 Games test users on the tiny config): the UNMODIFIED
reference (`BSSD`, `target_generate`) on the hf_fp32 stack of tools/make_golden.py -- real transformers fp32 tiny target and
its correlated 1-layer draft -- at K=10, N=40, gamma=3, users spread over the whole test set, the strict and the positional
constraint alternating.  Compact records (ranked items, scores, accepted steps per round) -> tests/golden/bssd_strict_users.json,
replayed by tests/test_oracle_bssd.py on the CPU.  Run in the build container (needs /root/reference)."""
import json
import os
import re
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as MG   # noqa: E402  (imports the reference and the shims)

LR, load_dataset = MG.LR, MG.load_dataset


def main(n_users=64):
    t0 = time.time()
    cases = []
    for name in ("beauty", "games"):
        ds = load_dataset(name)
        rds = MG.reference_dataset(name)
        tok_re = re.compile(r"<[abcd]_\d+>")
        enc = lambda s: [MG.BOS_ID] + [ds.token_id[t] for t in tok_re.findall(s)]
        ref_trie = MG.RefTrie([enc(e) + [MG.EOS_ID] for e in rds.get_all_items()])
        fns = {"strict": MG.suffix_prefix_allowed_tokens_fn(ref_trie, MG.RESPONSE_SEP),
               "positional": rds.get_prefix_allowed_tokens_fn(MG.FakeTokenizer(ds))}
        V = ds.vocab_size
        sh_t, sh_d = LR.shape_tiny_target(V), LR.shape_tiny_draft(V)
        Wt = LR.make_weights(sh_t, 0, std=1.28 / 8)
        Wd = LR.make_weights(sh_d, 2, std=1.28 / 8, like=Wt, noise=0.05)
        hf_t, hf_d = MG.hf_model(sh_t, Wt), MG.hf_model(sh_d, Wd)
        users = sorted({int(x) for x in np.linspace(3, ds.n_users - 2, n_users)})
        for i, u in enumerate(users):
            cname = "strict" if i % 2 == 0 else "positional"
            prompt = ds.prompt_ids(u)
            out, rounds = MG.run_reference(MG.LegacyKV(hf_t), MG.LegacyKV(hf_d), prompt, 10, 40, 3, fns[cname])
            hf_t.generation_config.num_beams = 10
            tg = MG.ref_beamSD.target_generate(MG.LegacyKV(hf_t), {"input_ids": torch.tensor([prompt])}, 4,
                                               prefix_allowed_tokens_fn=fns[cname])
            assert torch.equal(out["beam_sequence"], tg["beam_sequence"]), "strict BSSD must be lossless"
            P = len(prompt)
            cases.append({"stack": "hf_fp32", "dataset": name, "user": u, "draft": "correlated", "constraint": cname,
                          "K": 10, "N": 40, "gamma": 3, "bssd": MG.pack(out, P), "n_run": out["n_run"],
                          "total_accept_steps": out["total_accept_steps"],
                          "accept_steps": [r["n_matches"] for r in rounds]})
        print(name, len(cases), f"{time.time() - t0:.0f}s", flush=True)
    # Recall / NDCG of ranked lists, by the reference's own computeTopNAccuracy (code/utils.py:215-271).  Random-init models
    # never hit the ground truth, so the lists are the recorded ones with the user's first ground-truth item planted at
    # rank (i mod 10) for two users out of three -- the recipe tests/test_runner_metrics.py repeats.
    from utils import computeTopNAccuracy
    metrics = {}
    for name in ("beauty", "games"):
        ds = load_dataset(name)
        rds = MG.reference_dataset(name)
        gts, preds = [], []
        for i, c in enumerate([c for c in cases if c["dataset"] == name]):
            gt = list(rds[c["user"]]["labels"])
            names = ds.decode_items(c["bssd"]["items"])
            if i % 3 != 2:
                names[i % 10] = gt[0]
            gts.append(gt)
            preds.append(names)
        metrics[name] = [list(x) for x in computeTopNAccuracy(gts, preds, [1, 5, 10])]
    json.dump({"cases": cases, "metrics": {"topN": [1, 5, 10], "values": metrics}},
              open(os.path.join(MG.OUT, "bssd_strict_users.json"), "w"))
    print("reference metrics", metrics)
    print("cases", len(cases), "accept histogram", np.bincount([c["total_accept_steps"] for c in cases]).tolist())


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
