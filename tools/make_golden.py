"""Generate tests/golden/*.json by running the UNMODIFIED reference (/root/reference/code/beamSD.py,
generation_trie.py, data.py) in this container.  The reference cannot travel to the GPU box, so its
outputs are committed as small fixtures together with this script.

Two model stacks drive the reference's BSSD / target_generate:
  hf_fp32   real transformers LlamaForCausalLM (fp32, CPU) behind the SURVEY 8c `LegacyKV` proxy --
            the reference's own stack; also checked against HF `generate(num_beams=K)`.
  ref_bf16  oracle.llama_ref.RefLlama in bf16-rounding mode behind HFStyleProxy -- the numerical
            contract of the CUDA forward; these cases are what the GPU end-to-end tests replay.
Weights are never stored: they are regenerated from (shape, seed, std) by oracle.llama_ref.make_weights.

Shims (none edits the reference): `ipdb` stub module, `torch.cuda.synchronize` no-op on this
CUDA-less host, legacy list-of-(k,v) caches wrapped into DynamicCache for transformers 5.5.
"""
import json, os, re, sys, types, time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
sys.path.insert(0, "/root/reference/code")
torch.cuda.synchronize = lambda *a, **k: None

import beamSD as ref_beamSD                       # noqa: E402  (the reference, unmodified)
from generation_trie import Trie as RefTrie        # noqa: E402
import data as ref_data                            # noqa: E402

from transformers import GenerationConfig, LlamaConfig, LlamaForCausalLM   # noqa: E402
from transformers.cache_utils import DynamicCache                            # noqa: E402

from atspeed_b200.prompts import load_dataset, RESPONSE_SEP, BOS_ID, EOS_ID  # noqa: E402
from atspeed_b200.generation_trie import suffix_prefix_allowed_tokens_fn      # noqa: E402
from oracle import llama_ref as LR                                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class FakeTokenizer:
    """Just enough of LlamaTokenizer for reference code/data.py:84-104."""
    eos_token_id = EOS_ID

    def __init__(self, ds):
        self.ds = ds

    def __call__(self, text):
        if text == "Response:":
            return {"input_ids": [BOS_ID] + list(RESPONSE_SEP)}
        return {"input_ids": [BOS_ID, self.ds.token_id[text]]}


class LegacyKV:  # SURVEY 8c proxy, verbatim idea
    def __init__(self, m):
        object.__setattr__(self, "_m", m)

    def __getattr__(self, k):
        return getattr(self._m, k)

    def __call__(self, **kw):
        pkv = kw.get("past_key_values")
        if isinstance(pkv, (list, tuple)):
            kw["past_key_values"] = DynamicCache(ddp_cache_data=[(k, v) for k, v, *_ in pkv])
        return self._m(**kw)


def hf_model(shape, W):
    cfg = LlamaConfig(vocab_size=shape.vocab, hidden_size=shape.hidden, intermediate_size=shape.mlp,
                      num_hidden_layers=shape.n_layers, num_attention_heads=shape.n_heads,
                      num_key_value_heads=shape.n_heads, rms_norm_eps=shape.eps, rope_theta=shape.rope_theta,
                      tie_word_embeddings=False, attention_bias=False, pad_token_id=0, bos_token_id=1,
                      eos_token_id=2)
    m = LlamaForCausalLM(cfg).eval()
    m.load_state_dict(LR.weights_to_hf_state_dict(W), strict=True)
    return m


def reference_dataset(name):
    args = SimpleNamespace(dataset=name, data_path="/root/reference/data", max_his_len=20, his_sep=", ",
                           index_file=".LCRec-1e-3lr.json", add_prefix=False, llama=True, subseq=False)
    return ref_data.SeqRecDataset(args, mode="test", sample_num=-1)


class Recorder:
    """Records what the reference's draft_beam_search / verify return, by wrapping the module-level
    names BSSD looks up (no source edit)."""

    def __init__(self):
        self.rounds = []
        self._d, self._v = ref_beamSD.draft_beam_search, ref_beamSD.verify

    def __enter__(self):
        def d(*a, **k):
            out = self._d(*a, **k)
            self.rounds.append({
                "step_len": [int(x) for x in out["step_len"]],
                "draft_tokens": [[int(t) for t in x] for x in out["step_beam_tokens"]],
                "draft_parents": [[int(t) for t in x] for x in out["step_beam_indices"]],
            })
            return out

        def v(*a, **k):
            out = self._v(*a, **k)
            self.rounds[-1]["n_matches"] = int(out["n_matches"])
            self.rounds[-1]["beam_scores"] = [float(x) for x in out["beam_scores"]]
            self.rounds[-1]["beam_last_tokens"] = [int(x) for x in out["beam_sequence"][:, -1]]
            return out

        ref_beamSD.draft_beam_search, ref_beamSD.verify = d, v
        return self

    def __exit__(self, *a):
        ref_beamSD.draft_beam_search, ref_beamSD.verify = self._d, self._v


def run_reference(target, draft, prompt, K, N, gamma, fn, do_sample=False, seed=None):
    ids = torch.tensor([prompt])
    for m, nb in ((target, K), (draft, N)):
        m.generation_config.num_beams = nb
        m.generation_config.max_new_tokens = 4
        m.generation_config.do_sample = do_sample
        m.generation_config.temperature = 1.0
        m.generation_config.top_k = 50 if do_sample else None
    if seed is not None:
        torch.manual_seed(seed)
    with Recorder() as rec:
        out = ref_beamSD.BSSD(target, draft, {"input_ids": ids}, gamma, 4, prefix_allowed_tokens_fn=fn)
    return out, rec.rounds


def pack(out, P):
    return {"items": out["beam_sequence"][:, P:].tolist(), "scores": [float(x) for x in out["beam_scores"]]}


def main():
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    facts = {}
    cases = []
    for name in ("beauty", "games"):
        ds = load_dataset(name)
        rds = reference_dataset(name)
        # ---- pin the prompt builder: same users, same history items, same ground truth -----------
        assert len(rds) == ds.n_users
        tok_re = re.compile(r"<[abcd]_\d+>")
        for u in list(range(0, ds.n_users, 97)) + [ds.n_users - 1]:
            d = rds[u]
            ref_codes = [ds.token_id[t] for t in tok_re.findall(d["input_ids"])]
            mine = [t for t in ds.prompt_ids(u) if t >= 32000]
            assert ref_codes == mine, (name, u)
            assert d["labels"] == ds.ground_truth_strings(u), (name, u)
        assert sorted(rds.get_new_tokens()) == sorted(ds.token_id, key=ds.token_id.get)
        # ---- constraint functions of the reference ---------------------------------------------------
        all_items = rds.get_all_items()
        enc = lambda s: [BOS_ID] + [ds.token_id[t] for t in tok_re.findall(s)]
        ref_trie = RefTrie([enc(e) + [EOS_ID] for e in all_items])          # code/inference.py:130
        strict_fn = suffix_prefix_allowed_tokens_fn(ref_trie, RESPONSE_SEP)   # generate_teacher_data.py:174-188
        pos_fn = rds.get_prefix_allowed_tokens_fn(FakeTokenizer(ds))          # code/data.py:84-104
        facts[name] = {"vocab": ds.vocab_size, "n_users": ds.n_users, "n_item_seqs": len(ref_trie),
                       "level_ranges": ds.level_ranges(),
                       "positional_sizes": [len(pos_fn(0, torch.tensor(ds.prompt_ids(0) + [0] * d))) for d in range(5)],
                       "root_children": len(strict_fn(0, torch.tensor(ds.prompt_ids(0))))}
        V = ds.vocab_size
        users = [0, 1, 2, 5, 17, 100, 1000, ds.n_users - 1]

        # ---- stack 1: the reference's own stack, HF fp32 tiny models ---------------------------------
        sh_t, sh_d = LR.shape_tiny_target(V), LR.shape_tiny_draft(V)
        Wt = LR.make_weights(sh_t, 0, std=1.28 / 8)
        Wd_ind = LR.make_weights(sh_d, 1, std=1.28 / 8)
        Wd_cor = LR.make_weights(sh_d, 2, std=1.28 / 8, like=Wt, noise=0.05)
        hf_t = hf_model(sh_t, Wt)
        for dname, Wd in (("independent", Wd_ind), ("correlated", Wd_cor)):
            hf_d = hf_model(sh_d, Wd)
            for cname, fn in (("strict", strict_fn), ("positional", pos_fn)):
                grid = [(10, 10, 3), (10, 40, 3), (20, 40, 3), (5, 40, 2), (1, 40, 4), (20, 20, 2)]
                for (K, N, gamma) in grid:
                    for u in users[: (8 if (K, N, gamma) == (10, 40, 3) else 3)]:
                        prompt = ds.prompt_ids(u)
                        out, rounds = run_reference(LegacyKV(hf_t), LegacyKV(hf_d), prompt, K, N, gamma, fn)
                        hf_t.generation_config.num_beams = K
                        tg = ref_beamSD.target_generate(LegacyKV(hf_t), {"input_ids": torch.tensor([prompt])}, 4,
                                                        prefix_allowed_tokens_fn=fn)
                        P = len(prompt)
                        assert torch.equal(out["beam_sequence"], tg["beam_sequence"]), "strict BSSD must be lossless"
                        cases.append({"stack": "hf_fp32", "dataset": name, "user": u, "draft": dname,
                                      "constraint": cname, "K": K, "N": N, "gamma": gamma,
                                      "bssd": pack(out, P), "target_generate": pack(tg, P),
                                      "n_run": out["n_run"], "total_accept_steps": out["total_accept_steps"],
                                      "ave_accept_tokens": out["ave_accept_tokens"], "rounds": rounds})
            # HF generate == target_generate, a handful of cases
            for u in users[:2]:
                prompt = ds.prompt_ids(u)
                g = hf_t.generate(input_ids=torch.tensor([prompt]), num_beams=10, num_return_sequences=10,
                                  max_new_tokens=4, do_sample=False, prefix_allowed_tokens_fn=strict_fn,
                                  return_dict_in_generate=True, output_scores=True, length_penalty=1.0,
                                  early_stopping=False, pad_token_id=0)
                hf_t.generation_config.num_beams = 10
                tg = ref_beamSD.target_generate(LegacyKV(hf_t), {"input_ids": torch.tensor([prompt])}, 4,
                                                prefix_allowed_tokens_fn=strict_fn)
                assert torch.equal(g.sequences, tg["beam_sequence"]), "HF generate != target_generate"
                assert torch.allclose(g.sequences_scores * 4, tg["beam_scores"], atol=1e-4)
        print("hf_fp32 done", len(cases), f"{time.time() - t0:.0f}s")

        # ---- stack 2: bf16-contract oracle model behind the reference's algorithm ----------------------
        sh_t, sh_d = LR.shape_small_target(V), LR.shape_small_draft(V)
        Wt = LR.make_weights(sh_t, 10, std=1.28 / 16, dtype=torch.bfloat16)
        Wd_ind = LR.make_weights(sh_d, 11, std=1.28 / (128 ** 0.5), dtype=torch.bfloat16)
        sh_dc = LR.LlamaShape(V, sh_t.hidden, 1, sh_t.n_heads, sh_t.mlp)
        Wd_cor = LR.make_weights(sh_dc, 12, std=1.28 / 16, dtype=torch.bfloat16, like=Wt, noise=0.03)
        gc = lambda: GenerationConfig(num_beams=1, max_new_tokens=4, do_sample=False)
        for dname, shd, Wd in (("independent", sh_d, Wd_ind), ("correlated", sh_dc, Wd_cor)):
            pt = LR.HFStyleProxy(LR.RefLlama(sh_t, Wt, "bf16"), gc())
            pd_ = LR.HFStyleProxy(LR.RefLlama(shd, Wd, "bf16"), gc())
            for cname, fn in (("strict", strict_fn), ("positional", pos_fn)):
                grid = [(10, 40, 3), (20, 40, 3), (5, 10, 2), (1, 40, 3), (10, 10, 4)]
                for (K, N, gamma) in grid:
                    for u in users[: (6 if (K, N, gamma) == (10, 40, 3) else 2)]:
                        prompt = ds.prompt_ids(u)
                        out, rounds = run_reference(pt, pd_, prompt, K, N, gamma, fn)
                        P = len(prompt)
                        cases.append({"stack": "ref_bf16", "dataset": name, "user": u, "draft": dname,
                                      "constraint": cname, "K": K, "N": N, "gamma": gamma,
                                      "bssd": pack(out, P), "n_run": out["n_run"],
                                      "total_accept_steps": out["total_accept_steps"],
                                      "ave_accept_tokens": out["ave_accept_tokens"], "rounds": rounds})
        print("ref_bf16 done", len(cases), f"{time.time() - t0:.0f}s")

    from utils import computeTopNAccuracy            # reference code/utils.py:215-271
    metric = computeTopNAccuracy([["x", "y"], ["z"], []], [["x", "q", "y"], ["a", "b", "z"], ["a", "b", "c"]], [1, 3])
    facts["metric_known_answer"] = [list(x) for x in metric]
    json.dump({"facts": facts, "weights": {
        "hf_fp32": {"target": ["tiny_target", 0], "independent": ["tiny_draft", 1], "correlated": ["tiny_draft", 2, 0.05],
                    "std": 1.28 / 8},
        "ref_bf16": {"target": ["small_target", 10], "independent": ["small_draft", 11],
                     "correlated": ["small_target_1layer", 12, 0.03]}}, "cases": cases},
              open(os.path.join(OUT, "bssd_strict.json"), "w"))
    print("cases", len(cases), "accept histogram",
          np.bincount([c["total_accept_steps"] for c in cases]).tolist(), f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
