"""Device-side model + search session on top of the C ABI (include/atspeed.h).

PyTorch is used for what it is good at here -- allocating device memory and owning the CUDA stream;
every computation on the path is a kernel of libatspeed_b200.so.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .constraint import CSRTrie


@dataclass
class ModelSpec:
    vocab: int
    hidden: int
    n_layers: int
    n_heads: int
    head_dim: int
    mlp: int
    eps: float = 1e-6
    rope_theta: float = 10000.0


_LAYER_KEYS = ("wq", "wk", "wv", "wo", "wg", "wu", "wd", "ln1", "ln2")
_HF_NAMES = {"wq": "self_attn.q_proj", "wk": "self_attn.k_proj", "wv": "self_attn.v_proj", "wo": "self_attn.o_proj",
             "wg": "mlp.gate_proj", "wu": "mlp.up_proj", "wd": "mlp.down_proj", "ln1": "input_layernorm",
             "ln2": "post_attention_layernorm"}


def rope_tables(head_dim: int, theta: float, max_pos: int, round_bf16: bool = True):
    """cos/sin exactly as HF LlamaRotaryEmbedding computes them (fp32 on the host), rounded to bf16 the way a
    bf16 model sees them (or left in fp32 for the fp32 parity mode), returned as fp32 [max_pos, head_dim/2]."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    fr = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
    if not round_bf16:
        return fr.cos().contiguous(), fr.sin().contiguous()
    return fr.cos().to(torch.bfloat16).float().contiguous(), fr.sin().to(torch.bfloat16).float().contiguous()


class DeviceModel:
    """A LLaMA decoder's weights as CUDA tensors plus the atspeed_model_desc pointing at them.
    dtype bfloat16 = the production path (tcgen05 GEMMs); float32 = the exact-parity mode (fp32 SIMT forward,
    csrc/forward_f32.cu), meant for the small parity configurations."""

    def __init__(self, spec: ModelSpec, weights: Dict, device, max_pos: int = 1024, dtype=torch.bfloat16):
        self.spec, self.device = spec, torch.device(device)
        if dtype not in (torch.bfloat16, torch.float32):
            raise _lib.AtSpeedError(f"unsupported weight dtype {dtype}: bfloat16 or float32")
        self.dtype = dtype

        def dev(t):
            t = t.detach()
            if t.dtype != dtype or t.device != self.device or not t.is_contiguous():
                t = t.to(self.device, dtype).contiguous()
            return t

        self.embed, self.norm, self.lm_head = dev(weights["embed"]), dev(weights["norm"]), dev(weights["lm_head"])
        self.layers = [{k: dev(ly[k]) for k in _LAYER_KEYS} for ly in weights["layers"]]
        assert len(self.layers) == spec.n_layers
        cos, sin = rope_tables(spec.head_dim, spec.rope_theta, max_pos, round_bf16=dtype == torch.bfloat16)
        self.rope_cos, self.rope_sin = cos.to(self.device), sin.to(self.device)
        self._ptrs = (C.c_void_p * (spec.n_layers * 9))()
        for i, ly in enumerate(self.layers):
            for j, k in enumerate(_LAYER_KEYS):
                self._ptrs[i * 9 + j] = ly[k].data_ptr()
        d = _lib.ModelDesc()
        d.vocab, d.hidden, d.n_layers, d.n_heads = spec.vocab, spec.hidden, spec.n_layers, spec.n_heads
        d.head_dim, d.mlp, d.rms_eps = spec.head_dim, spec.mlp, spec.eps
        d.embed, d.final_norm, d.lm_head = self.embed.data_ptr(), self.norm.data_ptr(), self.lm_head.data_ptr()
        d.layer_weights = C.cast(self._ptrs, C.POINTER(C.c_void_p))
        d.rope_cos, d.rope_sin, d.max_pos = self.rope_cos.data_ptr(), self.rope_sin.data_ptr(), max_pos
        d.weights_f32 = 1 if dtype == torch.float32 else 0
        self.desc = d

    @classmethod
    def from_hf(cls, model, device=None, max_pos: int = 1024, dtype=None) -> "DeviceModel":
        """From a transformers LlamaForCausalLM (the object the reference passes as target_model /
        draft_model, code/inference.py:76-100).  bf16 CUDA weights are used in place (zero copy), fp16 weights are
        converted to bf16 copies; an fp32 model keeps fp32 and runs the exact-parity fp32 forward unless `dtype`
        says otherwise."""
        cfg = model.config
        n_kv = getattr(cfg, "num_key_value_heads", None) or cfg.num_attention_heads
        if n_kv != cfg.num_attention_heads:
            raise _lib.AtSpeedError("grouped-query attention is not supported (LLaMA-7B/68M are MHA)")
        head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        theta = getattr(cfg, "rope_theta", None)
        if theta is None:
            theta = (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0)
        spec = ModelSpec(cfg.vocab_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, head_dim,
                         cfg.intermediate_size, cfg.rms_norm_eps, float(theta))
        sd = model.state_dict()
        for k, v in sd.items():
            if v.dtype in (torch.int8, torch.uint8):
                raise _lib.AtSpeedError(f"{k} is quantised; load the target in bf16/fp16")
        W = {"embed": sd["model.embed_tokens.weight"], "norm": sd["model.norm.weight"],
             "lm_head": sd.get("lm_head.weight", sd["model.embed_tokens.weight"]), "layers": []}
        if W["embed"].shape[0] != spec.vocab:
            spec.vocab = W["embed"].shape[0]
        for i in range(spec.n_layers):
            W["layers"].append({k: sd[f"model.layers.{i}.{n}.weight"] for k, n in _HF_NAMES.items()})
        device = device if device is not None else W["embed"].device
        if dtype is None:
            dtype = torch.float32 if W["embed"].dtype == torch.float32 else torch.bfloat16
        return cls(spec, W, device, max_pos, dtype)


def as_device_model(model, device=None) -> DeviceModel:
    if isinstance(model, DeviceModel):
        return model
    cached = getattr(model, "_atspeed_device_model", None)
    if cached is None:
        cached = DeviceModel.from_hf(getattr(model, "_m", model), device)
        try:
            object.__setattr__(model, "_atspeed_device_model", cached)
        except Exception:
            pass
    return cached


class DeviceTrie:
    def __init__(self, csr: CSRTrie, device):
        self.csr = csr
        self.child_off = torch.from_numpy(np.ascontiguousarray(csr.child_off)).to(device)
        self.child_tok = torch.from_numpy(np.ascontiguousarray(csr.child_tok)).to(device)
        self.child_node = torch.from_numpy(np.ascontiguousarray(csr.child_node)).to(device)
        d = _lib.TrieDesc()
        d.child_off, d.child_tok, d.child_node = (self.child_off.data_ptr(), self.child_tok.data_ptr(),
                                                  self.child_node.data_ptr())
        d.n_nodes, d.n_edges = csr.n_nodes, csr.n_edges
        self.desc = d


class Session:
    """One search configuration (model pair, K, N, max_new_tokens, constraint) with its workspace."""

    def __init__(self, target: DeviceModel, draft: Optional[DeviceModel], trie: DeviceTrie, K: int, N: int,
                 max_new_tokens: int = 4, max_prompt: Optional[int] = None, do_sample: bool = False,
                 top_k: Optional[int] = None, temperature: float = 1.0, seed: int = 0, max_users: int = 1,
                 cohort_tokens: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AtSpeedError("atspeed_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.target_model, self.draft_model, self.trie = target, draft, trie
        self.K, self.N, self.L = K, N, max_new_tokens
        if max_prompt is None:
            max_prompt = 512 - max(max_new_tokens - 1, 1) * N
        self.max_prompt = max_prompt
        self.device = target.device
        self.do_sample = bool(do_sample)
        cfg = _lib.Config(K, N, max_new_tokens, max_prompt,
                          torch.cuda.get_device_properties(self.device).multi_processor_count,
                          1 if do_sample else 0, int(top_k or 0), float(temperature or 1.0), int(seed) & (2 ** 64 - 1),
                          int(max_users), int(cohort_tokens))
        self.max_users = int(max_users)
        self.cfg = cfg
        nbytes = C.c_size_t(0)
        dptr = C.byref(draft.desc) if draft is not None else None
        _lib.check(self.lib.atspeed_session_workspace_bytes(C.byref(target.desc), dptr, C.byref(cfg), C.byref(nbytes)))
        self.workspace = torch.zeros(nbytes.value + 1024, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 1023) & ~1023
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_session_create(C.byref(target.desc), dptr, C.byref(cfg), C.byref(trie.desc),
                                                       C.c_void_p(base), nbytes.value, C.byref(handle)))
        self.handle = handle
        self._tok = (C.c_int32 * (K * max_new_tokens))()
        self._sc = (C.c_float * K)()
        self._cnt = C.c_int32(0)
        self._nm = C.c_int32(0)
        self._finalizer = weakref.finalize(self, self.lib.atspeed_session_destroy, handle)
        info = (C.c_int64 * 8)()
        _lib.check(self.lib.atspeed_session_info(self.handle, info))
        self.ldl, self.R_max, self.T_max, self.S_max, self.A_cap = (int(info[i]) for i in range(5))

    # -- helpers ---------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _prompt(prompt_ids: Sequence[int]):
        arr = np.ascontiguousarray(np.asarray(prompt_ids, dtype=np.int32))
        return arr, arr.ctypes.data_as(_lib.c_i32p), int(arr.shape[0])

    def _collect(self, stats: Optional[_lib.Stats] = None) -> Dict:
        n = self._cnt.value
        toks = np.ctypeslib.as_array(self._tok).reshape(self.K, self.L)[:n].copy()
        scores = np.ctypeslib.as_array(self._sc)[:n].copy()
        out = {"tokens": toks, "scores": scores}
        if stats is not None:
            out.update(n_run=stats.n_run, total_accept_steps=stats.total_accept_steps,
                       accept_steps=[stats.accept_steps[i] for i in range(min(stats.n_run, 8))],
                       target_forwards=stats.target_forwards, draft_forwards=stats.draft_forwards,
                       kernel_launches=stats.kernel_launches)
        return out

    # -- whole-loop entry points (one C call per user) ---------------------------------------------
    def bssd(self, prompt_ids: Sequence[int], gamma: int) -> Dict:
        arr, ptr, P = self._prompt(prompt_ids)
        st = _lib.Stats()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_bssd(self.handle, ptr, P, gamma, self._tok, self._sc, C.byref(self._cnt),
                                             C.byref(st), self._stream()))
        return self._collect(st)

    def set_shared_prefix(self, prefix_ids: Sequence[int]) -> int:
        """Cohort sessions: the tokens every prompt of the coming bssd_batch* calls starts with (the instruction template of
        the reference's prompts).  Their K/V rows are computed once here and copied into each user's caches on admission; a
        prompt that does not start with them makes the call fail.  Empty = off.  Returns the prefix length in use."""
        arr = np.ascontiguousarray(np.asarray(list(prefix_ids), dtype=np.int32))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_session_set_shared_prefix(self.handle, arr.ctypes.data_as(_lib.c_i32p), int(arr.shape[0]),
                                                                  self._stream()))
        self.shared_prefix = int(arr.shape[0])
        return self.shared_prefix

    def bssd_batch(self, prompts: Sequence[Sequence[int]], gamma: int) -> List[Dict]:
        """Cohort mode (Session(max_users > 1)): speculative beam search for all `prompts`, up to max_users of them in
        flight at a time with their trees packed into shared forwards.  Returns one dict per prompt, in order."""
        n = len(prompts)
        lens = np.asarray([len(p) for p in prompts], dtype=np.int32)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.int32) for p in prompts]))
        toks = np.zeros((n, self.K, self.L), dtype=np.int32)
        scores = np.zeros((n, self.K), dtype=np.float32)
        cnt = np.zeros(n, dtype=np.int32)
        stats = (_lib.Stats * n)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_bssd_batch(self.handle, n, flat.ctypes.data_as(_lib.c_i32p),
                                                   lens.ctypes.data_as(_lib.c_i32p), gamma, toks.ctypes.data_as(_lib.c_i32p),
                                                   scores.ctypes.data_as(_lib.c_f32p), cnt.ctypes.data_as(_lib.c_i32p),
                                                   C.cast(stats, C.c_void_p), self._stream()))
        out = []
        for i in range(n):
            st = stats[i]
            out.append({"tokens": toks[i, : cnt[i]].copy(), "scores": scores[i, : cnt[i]].copy(), "n_run": st.n_run,
                        "total_accept_steps": st.total_accept_steps,
                        "accept_steps": [st.accept_steps[r] for r in range(min(st.n_run, 8))],
                        "target_forwards": st.target_forwards, "draft_forwards": st.draft_forwards,
                        "kernel_launches": st.kernel_launches})
        return out

    def bssd_batch_device(self, prompts_dev: torch.Tensor, lens: Sequence[int], gamma: int, tokens_dev: torch.Tensor,
                          scores_dev: torch.Tensor) -> List[Dict]:
        """Cohort mode with the concatenated prompts (int32 CUDA tensor) and the results (int32 [n,K,6], fp32 [n,K] CUDA
        tensors) resident in HBM."""
        n = len(lens)
        la = np.asarray(lens, dtype=np.int32)
        stats = (_lib.Stats * n)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_bssd_batch_device(self.handle, n, prompts_dev.data_ptr(), la.ctypes.data_as(_lib.c_i32p),
                                                          gamma, tokens_dev.data_ptr(), scores_dev.data_ptr(),
                                                          C.cast(stats, C.c_void_p), self._stream()))
        return [{"n_run": st.n_run, "total_accept_steps": st.total_accept_steps, "target_forwards": st.target_forwards,
                 "draft_forwards": st.draft_forwards, "kernel_launches": st.kernel_launches} for st in stats]

    def bssd_device(self, prompt_dev: torch.Tensor, gamma: int, tokens_dev: torch.Tensor, scores_dev: torch.Tensor) -> Dict:
        """Prompt (int32 CUDA tensor) and results (int32 [K,6], fp32 [K] CUDA tensors) stay in HBM."""
        st = _lib.Stats()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_bssd_device(self.handle, prompt_dev.data_ptr(), int(prompt_dev.shape[0]), gamma,
                                                    tokens_dev.data_ptr(), scores_dev.data_ptr(), C.byref(st), self._stream()))
        return {"n_run": st.n_run, "total_accept_steps": st.total_accept_steps, "target_forwards": st.target_forwards,
                "draft_forwards": st.draft_forwards, "kernel_launches": st.kernel_launches}

    # -- sampling mode (AtSpeed-R) ---------------------------------------------------------------------
    def set_seed(self, seed: int, user_seq: int = 0):
        _lib.check(self.lib.atspeed_session_set_seed(self.handle, int(seed) & (2 ** 64 - 1), int(user_seq)))

    def sort_result(self):
        _lib.check(self.lib.atspeed_session_sort_result(self.handle, self._stream()))

    @property
    def sample_width(self) -> int:
        return int(self.lib.atspeed_session_sample_width(self.handle))

    def noise(self, seed: int, user_seq: int, rnd: int, level: int, site: int, kind: int, n: int) -> np.ndarray:
        """The exact noise the kernels consume at (user, round, level, site): kind 0 uint32 bits, 1 uniform, 2 Exp(1)."""
        stream = self.lib.atspeed_noise_stream(int(user_seq), int(rnd), int(level), int(site))
        out = torch.empty(n, dtype=torch.int32 if kind == 0 else torch.float32, device=self.device)
        _lib.check(self.lib.atspeed_noise_fill(int(seed) & (2 ** 64 - 1), stream, kind, n, out.data_ptr(), self._stream()))
        a = out.cpu().numpy()
        return a.view(np.uint32) if kind == 0 else a

    def profile(self, enable: bool):
        _lib.check(self.lib.atspeed_session_profile(self.handle, 1 if enable else 0))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        ms, cnt, by, fl = (C.c_double * 6)(), (C.c_int64 * 6)(), (C.c_double * 6)(), (C.c_double * 6)()
        _lib.check(self.lib.atspeed_session_profile_read(self.handle, ms, cnt, by, fl, self._stream()))
        names = ("gemm", "attention", "rowwise", "topk", "beam", "kvgather")
        return {n: {"ms": ms[i], "launches": int(cnt[i]), "bytes": by[i], "flops": fl[i]} for i, n in enumerate(names)}

    def target_generate(self, prompt_ids: Sequence[int]) -> Dict:
        arr, ptr, P = self._prompt(prompt_ids)
        st = _lib.Stats()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_target_generate(self.handle, ptr, P, self._tok, self._sc, C.byref(self._cnt),
                                                        C.byref(st), self._stream()))
        return self._collect(st)

    # -- stage-level entry points (mirror BSSD's loop body) ------------------------------------------
    def begin(self, prompt_ids: Sequence[int]):
        self._keep = self._prompt(prompt_ids)
        _lib.check(self.lib.atspeed_session_begin(self.handle, self._keep[1], self._keep[2], self._stream()))

    def draft(self, draft_len: int):
        _lib.check(self.lib.atspeed_session_draft(self.handle, draft_len, self._stream()))

    def target_forward(self, draft_len: int):
        _lib.check(self.lib.atspeed_session_target(self.handle, draft_len, self._stream()))

    def verify(self, draft_len: int) -> int:
        _lib.check(self.lib.atspeed_session_verify(self.handle, draft_len, C.byref(self._nm), self._stream()))
        return self._nm.value

    def step(self, model: int, width: int):
        _lib.check(self.lib.atspeed_session_step(self.handle, model, width, self._stream()))

    def result(self) -> Dict:
        _lib.check(self.lib.atspeed_session_result(self.handle, self._tok, self._sc, C.byref(self._cnt), self._stream()))
        return self._collect()

    # -- introspection -------------------------------------------------------------------------------
    def read(self, field: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        _lib.check(self.lib.atspeed_session_read(self.handle, field, out.ctypes.data_as(C.c_void_p), out.nbytes,
                                                 self._stream()))
        return out

    def levels(self) -> Dict[str, np.ndarray]:
        ML, MB = _lib.MAX_LEVELS, _lib.MAX_BEAMS
        return {"cnt": self.read(_lib.F_LEVEL_CNT, (ML,), np.int32),
                "tok": self.read(_lib.F_LEVEL_TOK, (ML, MB), np.int32),
                "parent": self.read(_lib.F_LEVEL_PARENT, (ML, MB), np.int32),
                "score": self.read(_lib.F_LEVEL_SCORE, (ML, MB), np.float32),
                "node": self.read(_lib.F_LEVEL_NODE, (ML, MB), np.int32)}

    def verify_trace(self) -> Dict[str, np.ndarray]:
        ML, MK = _lib.MAX_LEVELS, _lib.MAX_K
        return {"npick": self.read(_lib.F_NPICK, (ML,), np.int32),
                "pick_parent": self.read(_lib.F_PICK_PARENT, (ML, MK), np.int32),
                "pick_tok": self.read(_lib.F_PICK_TOK, (ML, MK), np.int32),
                "pick_score": self.read(_lib.F_PICK_SCORE, (ML, MK), np.float32),
                "hit_pos": self.read(_lib.F_HIT_POS, (ML, MK), np.int32),
                "acc": self.read(_lib.F_TR_ACC, (ML, _lib.MAX_BEAMS), np.int32),
                "lse_q": self.read(_lib.F_LSE_Q, (ML,), np.float32),
                "fallbacks": int(self.read(_lib.F_SCALARS, (16,), np.int32)[8])}

    def info(self):
        info = (C.c_int64 * 8)()
        _lib.check(self.lib.atspeed_session_info(self.handle, info))
        return [int(x) for x in info]

    def logits(self, model: int, rows: int) -> np.ndarray:
        full = self.read(_lib.F_LOGITS_TARGET if model == 0 else _lib.F_LOGITS_DRAFT, (rows, self.ldl), np.float32)
        return full[:, : self.target_model.spec.vocab]

    def forward_raw(self, model: int, tok, pos, slot, prefix_len, vis, vis_base: int, S: int, rows_idx):
        """tok/pos/slot/prefix_len int32 CUDA tensors [T], vis int32 CUDA [T,16], rows_idx int32 CUDA [R]."""
        T, R = tok.shape[0], rows_idx.shape[0]
        _lib.check(self.lib.atspeed_session_forward_raw(self.handle, model, tok.data_ptr(), pos.data_ptr(), slot.data_ptr(),
                                                        prefix_len.data_ptr(), vis.data_ptr(), vis_base, T, S,
                                                        rows_idx.data_ptr(), R, self._stream()))
        return self.logits(model, R)
