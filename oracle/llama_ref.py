"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the LLaMA decoder forward that the reference reaches through
`model(**inputs)` (reference code/beamSD.py:52,221 -> transformers LlamaForCausalLM.forward,
transformers pinned 4.41.0 by /root/reference/README.md:13; the transformers source is NOT under
/root/reference, so its published algorithm is restated here and pinned against the installed
transformers 5.5.0 `LlamaForCausalLM` in tests/test_oracle_llama.py and tools/make_golden.py).

Algorithm (HF modeling_llama): h = embed[ids]; per layer: x = RMSNorm(h)*g1; q,k,v = x Wq^T, x Wk^T,
x Wv^T; RoPE(rotate_half layout) on q,k at `position_ids`; attention over the KV cache with an
arbitrary visibility matrix (the reference passes a 4-D additive mask, 0 = visible, finfo.min =
hidden; beams live on the sequence axis, batch is 1); h += attn Wo^T; x = RMSNorm(h)*g2;
h += (silu(x Wg^T) * x Wu^T) Wd^T; logits = RMSNorm(h)*gf lm_head^T.

precision = "fp32": plain fp32 everywhere (what HF does for an fp32 model).
precision = "bf16": weights are bf16-representable and activations are rounded to bf16 at the
    points where an HF bf16 module rounds (after every Linear, inside RMSNorm, inside
    apply_rotary_pos_emb, after SiLU, after the gate*up product, after each residual add); GEMMs and
    the softmax accumulate in fp32.  Logits are returned in fp32 WITHOUT a final bf16 rounding
    unless round_logits=True (transformers 4.41 rounds them to bf16 and upcasts, SURVEY hard part 3).
    This is the numerical contract the CUDA forward (atspeed_b200/csrc/engine.cu `forward`: gemm.cu, elementwise.cu,
    attention.cu) is built to.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


@dataclass
class LlamaShape:
    vocab: int
    hidden: int
    n_layers: int
    n_heads: int
    mlp: int
    head_dim: int = 0
    rope_theta: float = 10000.0
    eps: float = 1e-6

    def __post_init__(self):
        if not self.head_dim:
            self.head_dim = self.hidden // self.n_heads


def shape_tiny_target(vocab):  # BASELINE.json configs[0]: tiny target, 2 layers
    return LlamaShape(vocab, 64, 2, 4, 128)


def shape_tiny_draft(vocab):   # 1-layer draft
    return LlamaShape(vocab, 64, 1, 4, 128)


def shape_small_target(vocab):  # GPU parity config: big enough to exercise every tile path
    return LlamaShape(vocab, 256, 2, 4, 512)


def shape_small_draft(vocab):
    return LlamaShape(vocab, 128, 1, 2, 256)


def shape_7b(vocab):           # LLaMA-7B shape (SURVEY 8 notation)
    return LlamaShape(vocab, 4096, 32, 32, 11008)


def shape_68m(vocab):          # LLaMA-68M shape (reference code/model.py:1023)
    return LlamaShape(vocab, 768, 2, 12, 3072)


WEIGHT_ORDER = ("embed", "layers", "norm", "lm_head")


def make_weights(shape: LlamaShape, seed: int, std: float = 0.02, dtype=torch.float32,
                 device="cpu", like: Optional[Dict] = None, noise: float = 0.0) -> Dict:
    """Deterministic random-init weights (normal(0, std); norm gains = 1), generated tensor by tensor
    in a fixed order from torch's CPU generator so the same call gives the same bits everywhere.
    `like` + `noise`: a perturbed copy of another model's weights (same shape, or fewer layers) --
    the "correlated draft" used to get non-zero acceptance (SURVEY 7 hard part 5)."""
    g = torch.Generator(device="cpu").manual_seed(seed)

    def rnd(*sz, base=None):
        w = torch.randn(*sz, generator=g, dtype=torch.float32) * std
        if base is not None:
            w = base.float().cpu() + noise * w / std * base.float().std()
        return w.to(dtype).to(device)

    def b(name, i=None):
        if like is None:
            return None
        return like[name] if i is None else like["layers"][i][name]

    W = {"embed": rnd(shape.vocab, shape.hidden, base=b("embed")), "layers": []}
    for i in range(shape.n_layers):
        ly = {}
        for nm, sz in (("wq", (shape.n_heads * shape.head_dim, shape.hidden)),
                       ("wk", (shape.n_heads * shape.head_dim, shape.hidden)),
                       ("wv", (shape.n_heads * shape.head_dim, shape.hidden)),
                       ("wo", (shape.hidden, shape.n_heads * shape.head_dim)),
                       ("wg", (shape.mlp, shape.hidden)),
                       ("wu", (shape.mlp, shape.hidden)),
                       ("wd", (shape.hidden, shape.mlp))):
            ly[nm] = rnd(*sz, base=b(nm, i))
        ly["ln1"] = torch.ones(shape.hidden, dtype=dtype, device=device)
        ly["ln2"] = torch.ones(shape.hidden, dtype=dtype, device=device)
        W["layers"].append(ly)
    W["norm"] = torch.ones(shape.hidden, dtype=dtype, device=device)
    W["lm_head"] = rnd(shape.vocab, shape.hidden, base=b("lm_head"))
    return W


def weights_to_hf_state_dict(W: Dict) -> Dict[str, torch.Tensor]:
    sd = {"model.embed_tokens.weight": W["embed"], "model.norm.weight": W["norm"],
          "lm_head.weight": W["lm_head"]}
    names = {"wq": "self_attn.q_proj", "wk": "self_attn.k_proj", "wv": "self_attn.v_proj",
             "wo": "self_attn.o_proj", "wg": "mlp.gate_proj", "wu": "mlp.up_proj", "wd": "mlp.down_proj",
             "ln1": "input_layernorm", "ln2": "post_attention_layernorm"}
    for i, ly in enumerate(W["layers"]):
        for k, v in ly.items():
            sd[f"model.layers.{i}.{names[k]}.weight"] = v
    return sd


def weights_from_hf(model) -> Dict:
    """Inverse of the above for an HF LlamaForCausalLM (shares storage)."""
    sd = model.state_dict()
    n = model.config.num_hidden_layers
    inv = {"wq": "self_attn.q_proj", "wk": "self_attn.k_proj", "wv": "self_attn.v_proj",
           "wo": "self_attn.o_proj", "wg": "mlp.gate_proj", "wu": "mlp.up_proj", "wd": "mlp.down_proj",
           "ln1": "input_layernorm", "ln2": "post_attention_layernorm"}
    W = {"embed": sd["model.embed_tokens.weight"], "norm": sd["model.norm.weight"],
         "lm_head": sd["lm_head.weight"], "layers": []}
    for i in range(n):
        W["layers"].append({k: sd[f"model.layers.{i}.{v}.weight"] for k, v in inv.items()})
    return W


def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


@dataclass
class RefCache:
    """KV cache: per layer k, v of shape [S, H, D] (fp32 tensors holding fp32 or bf16-valued data)."""
    k: List[torch.Tensor] = field(default_factory=list)
    v: List[torch.Tensor] = field(default_factory=list)

    def __len__(self):
        return 0 if not self.k else self.k[0].shape[0]

    def truncated(self, n: int) -> "RefCache":
        return RefCache([t[:n] for t in self.k], [t[:n] for t in self.v])


class RefLlama:
    def __init__(self, shape: LlamaShape, W: Dict, precision: str = "fp32", round_logits: bool = False):
        assert precision in ("fp32", "bf16")
        self.shape, self.precision, self.round_logits = shape, precision, round_logits
        self.r = _bf if precision == "bf16" else (lambda x: x)
        f = lambda t: t.detach().to("cpu", torch.float32)
        if precision == "bf16":
            f = lambda t: _bf(t.detach().to("cpu", torch.float32))
        self.W = {"embed": f(W["embed"]), "norm": f(W["norm"]), "lm_head": f(W["lm_head"]),
                  "layers": [{k: f(v) for k, v in ly.items()} for ly in W["layers"]]}
        d = shape.head_dim
        self.inv_freq = 1.0 / (shape.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))

    def new_cache(self) -> RefCache:
        return RefCache()

    def _rms(self, x, g):
        r = self.r
        var = x.pow(2).mean(-1, keepdim=True)
        return r(g * r(x * torch.rsqrt(var + self.shape.eps)))

    def _rope(self, x, pos):
        # x [T, H, D]; HF rotate_half layout (modeling_llama.apply_rotary_pos_emb)
        r = self.r
        fr = pos.to(torch.float32)[:, None] * self.inv_freq[None, :]
        emb = torch.cat((fr, fr), -1)
        cos, sin = r(emb.cos())[:, None, :], r(emb.sin())[:, None, :]
        h = x.shape[-1] // 2
        rot = torch.cat((-x[..., h:], x[..., :h]), -1)
        return r(r(x * cos) + r(rot * sin))

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, pos: torch.Tensor, vis: torch.Tensor, cache: RefCache,
                logit_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tokens [T], pos [T], vis bool [T, S_old + T] (True = query may attend that slot); the T new
        tokens occupy slots S_old .. S_old+T-1 of `cache` (appended in place). Returns fp32 logits
        for `logit_rows` (default all T rows)."""
        s, r, W = self.shape, self.r, self.W
        T = tokens.shape[0]
        S_old = len(cache)
        assert vis.shape == (T, S_old + T), (vis.shape, T, S_old)
        h = W["embed"][tokens]
        scale = 1.0 / math.sqrt(s.head_dim)
        first = not cache.k
        for li, ly in enumerate(W["layers"]):
            x = self._rms(h, ly["ln1"])
            q = r(x @ ly["wq"].T).view(T, s.n_heads, s.head_dim)
            k = r(x @ ly["wk"].T).view(T, s.n_heads, s.head_dim)
            v = r(x @ ly["wv"].T).view(T, s.n_heads, s.head_dim)
            q, k = self._rope(q, pos), self._rope(k, pos)
            if first:
                cache.k.append(k), cache.v.append(v)
            else:
                cache.k[li] = torch.cat((cache.k[li], k), 0)
                cache.v[li] = torch.cat((cache.v[li], v), 0)
            K_, V_ = cache.k[li], cache.v[li]
            sc = torch.einsum("thd,shd->hts", q, K_) * scale
            sc = sc.masked_fill(~vis[None], float("-inf"))
            p = torch.softmax(sc, -1)
            a = r(torch.einsum("hts,shd->thd", p, V_).reshape(T, -1))
            h = r(h + r(a @ ly["wo"].T))
            x = self._rms(h, ly["ln2"])
            g = r(x @ ly["wg"].T)
            u = r(x @ ly["wu"].T)
            m = r(r(torch.nn.functional.silu(g)) * u)
            h = r(h + r(m @ ly["wd"].T))
        if logit_rows is not None:
            h = h[logit_rows]
        x = self._rms(h, W["norm"])
        logits = x @ W["lm_head"].T
        return _bf(logits) if (self.round_logits and self.precision == "bf16") else logits


class HFStyleProxy:
    """Lets the UNMODIFIED reference beamSD.py drive a RefLlama: same call signature and return
    attributes as the HF model object the reference expects (code/beamSD.py:52-58,102,221-230,
    418-429): `model(input_ids=[1,T], attention_mask=[1,1,T,S] additive, position_ids=[1,T],
    past_key_values=None | [(k,v)] with k [1,H,S,D])` -> `.logits [1,T,V]`, `.past_key_values`.
    Used only by tools/make_golden.py and oracle tests."""

    class _Out:
        def __init__(self, logits, pkv):
            self.logits, self.past_key_values = logits, pkv

    def __init__(self, ref: RefLlama, generation_config, dtype=torch.float32):
        self.ref, self.generation_config = ref, generation_config
        self.dtype, self.device = dtype, torch.device("cpu")
        self.n_forward = 0

    def __call__(self, input_ids, attention_mask, position_ids, past_key_values=None):
        ref = self.ref
        cache = RefCache()
        if past_key_values is not None:
            for k, v, *_ in past_key_values:
                cache.k.append(k[0].permute(1, 0, 2).contiguous().float())
                cache.v.append(v[0].permute(1, 0, 2).contiguous().float())
        vis = attention_mask[0, 0].float() > torch.finfo(torch.float32).min / 2
        T = input_ids.shape[1]
        S = len(cache) + T
        assert vis.shape[1] == S, (vis.shape, S)
        logits = ref.forward(input_ids[0], position_ids[0], vis, cache)
        self.n_forward += 1
        pkv = tuple((k.permute(1, 0, 2)[None], v.permute(1, 0, 2)[None]) for k, v in zip(cache.k, cache.v))
        return self._Out(logits[None], pkv)

    def _get_logits_processor(self, generation_config=None, input_ids_seq_length=None,
                              encoder_input_ids=None, prefix_allowed_tokens_fn=None,
                              logits_processor=None, device=None, **kw):
        from transformers.generation.logits_process import (LogitsProcessorList,
                                                            PrefixConstrainedLogitsProcessor)
        out = LogitsProcessorList()
        if prefix_allowed_tokens_fn is not None:
            out.append(PrefixConstrainedLogitsProcessor(prefix_allowed_tokens_fn,
                                                        generation_config.num_beams))
        return out

    def _get_logits_warper(self, generation_config):
        # transformers 4.41 GenerationMixin._get_logits_warper restated (SURVEY 8c shim 4):
        # temperature (if != 1) then top-k (if set) with min_tokens_to_keep = 2 when num_beams > 1.
        from transformers.generation.logits_process import (LogitsProcessorList, TemperatureLogitsWarper,
                                                            TopKLogitsWarper)
        out = LogitsProcessorList()
        gc = generation_config
        if gc.temperature is not None and gc.temperature != 1.0:
            out.append(TemperatureLogitsWarper(gc.temperature))
        if gc.top_k is not None and gc.top_k != 0:
            out.append(TopKLogitsWarper(top_k=gc.top_k, min_tokens_to_keep=2 if gc.num_beams > 1 else 1))
        return out
