"""Drop-in host module for the reference's `beamSD.py` (file:line relative to /root/reference/code).

`from atspeed_b200.beamSD import *` gives the names `inference.py` binds through its star import
(code/inference.py:29): `BSSD` (:175), `target_generate` (:176), `Timer` (:177) and `set_seed`
(transformers', re-exported like beamSD.py:8), plus the stage functions `one_step_beam_search`,
`draft_beam_search`, `no_delay_draft_beam_search`, `target_beam_search`, `no_delay_target_beam_search`
and `verify`.  Signatures, the generation_config knobs that are read (num_beams on both models,
do_sample; beamSD.py:482-483,53,255) and the keys of the returned dicts are the reference's.

Behind the boundary nothing of the reference remains: the models' weights are handed (zero-copy when
they are bf16 CUDA tensors) to libatspeed_b200.so, the constraint callable is compiled once into a CSR
trie in HBM, and each stage is a handful of CUDA kernels enqueued by one C call.  The search state
(beams, tree masks, KV caches) stays on the device inside a `Session`; the dicts passed between the
stage functions carry that session instead of the reference's tensors.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Optional

import torch

from . import _lib
from .constraint import compile_constraint
from .engine import DeviceModel, DeviceTrie, Session, as_device_model

try:  # the reference re-exports transformers.set_seed through its star import (beamSD.py:8)
    from transformers import set_seed  # noqa: F401
except Exception:  # pragma: no cover
    def set_seed(seed: int):
        torch.manual_seed(seed)

__all__ = ["Timer", "set_seed", "one_step_beam_search", "draft_beam_search", "no_delay_draft_beam_search",
           "target_beam_search", "no_delay_target_beam_search", "verify", "BSSD", "target_generate"]


class Timer:
    """Context manager + decorator with the reference's semantics (beamSD.py:12-37): wall clock, a device
    synchronise on exit, the decorator injects result["time_cost"] and accepts `syn_device=`.  The only
    difference: on a CUDA-less host the synchronise is skipped instead of raising."""

    def __init__(self, func="", sync_cuda=True, syn_device=0):
        self.func, self.sync_cuda, self.syn_device = func, sync_cuda, syn_device

    def _sync(self):
        if self.sync_cuda and torch.cuda.is_available():
            torch.cuda.synchronize(self.syn_device)

    def __enter__(self):
        self.start = time.time()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self._sync()
        self.time_cost = time.time() - self.start

    def __call__(self, func):
        def wrapper(*args, syn_device=self.syn_device, **kwargs):
            self.syn_device = syn_device
            self.start = time.time()
            result = func(*args, **kwargs)
            self._sync()
            self.time_cost = time.time() - self.start
            result["time_cost"] = self.time_cost
            return result

        wrapper.__name__ = getattr(func, "__name__", "wrapped")
        wrapper.__doc__ = func.__doc__
        return wrapper


# ------------------------------------------------------------------------------------------------------
# session management
# ------------------------------------------------------------------------------------------------------
_SESSIONS: Dict[tuple, Session] = {}
_TRIES: Dict[tuple, DeviceTrie] = {}


def _device_of(model) -> torch.device:
    dev = getattr(model, "device", None)
    return torch.device(dev) if dev is not None else torch.device("cuda", torch.cuda.current_device())


def get_session(target_model, draft_model, prompt_ids: List[int], max_new_tokens: int,
                prefix_allowed_tokens_fn: Optional[Callable]) -> Session:
    """Session for (models, K, N, max_new_tokens, constraint); built on first use and cached."""
    gc = target_model.generation_config
    K = int(gc.num_beams)
    N = int(draft_model.generation_config.num_beams) if draft_model is not None else K
    # sampling knobs are read from the TARGET's config, like the reference's single warper (beamSD.py:479-481)
    do_sample = bool(getattr(gc, "do_sample", False))
    top_k = getattr(gc, "top_k", None) if do_sample else None
    temperature = float(getattr(gc, "temperature", None) or 1.0) if do_sample else 1.0
    if do_sample and not top_k:
        raise ValueError("do_sample=True needs generation_config.top_k (transformers 4.41's default was 50; "
                         "5.x defaults to None): the warped candidate set must be bounded (<= 64)")
    tdm = as_device_model(target_model)
    ddm = as_device_model(draft_model) if draft_model is not None else None
    csr = compile_constraint(prefix_allowed_tokens_fn, prompt_ids, max_new_tokens, vocab_size=tdm.spec.vocab)
    tkey = (id(csr), str(tdm.device))
    trie = _TRIES.get(tkey)
    if trie is None:
        trie = _TRIES[tkey] = DeviceTrie(csr, tdm.device)
    key = (id(tdm), id(ddm), K, N, max_new_tokens, id(trie), do_sample, top_k, temperature)
    sess = _SESSIONS.get(key)
    if sess is None:
        sess = _SESSIONS[key] = Session(tdm, ddm, trie, K, N, max_new_tokens, do_sample=do_sample, top_k=top_k,
                                        temperature=temperature)
    return sess


def _reseed(sess: Session, seed: Optional[int]):
    """Sampling mode: key the search's noise.  Like the reference, randomness follows torch's global generator
    (`set_seed(...)` makes runs reproducible): one 62-bit seed is drawn from it per search unless given."""
    if not sess.do_sample:
        return None
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    sess.set_seed(seed, 0)
    return seed


def clear_sessions():
    _SESSIONS.clear()
    _TRIES.clear()


def _prompt_list(inputs: Dict) -> List[int]:
    ids = inputs["input_ids"]
    if ids.dim() != 2 or ids.shape[0] != 1:
        raise ValueError("batch size 1 only (as the reference: beamSD.py:57,224 read batch index 0)")
    return [int(t) for t in ids[0].tolist()]


# ------------------------------------------------------------------------------------------------------
# stage functions (same names / argument order as the reference; `inputs` carries the Session)
# ------------------------------------------------------------------------------------------------------
def _levels_dict(sess: Session, draft_len: int) -> Dict:
    lv = sess.levels()
    cnt = lv["cnt"]
    out = {"step_len": [int(cnt[l]) for l in range(draft_len + 1)],
           "step_beam_tokens": tuple(torch.from_numpy(lv["tok"][l, :cnt[l]].astype("int64")) for l in range(1, draft_len + 1)),
           "step_beam_indices": tuple(torch.from_numpy(lv["parent"][l, :cnt[l]].astype("int64")) for l in range(1, draft_len + 1)),
           "step_beam_scores": tuple(torch.from_numpy(lv["score"][l, :cnt[l]].copy()) for l in range(1, draft_len + 1))}
    V = sess.target_model.spec.vocab
    out["step_seq_tokens"] = tuple(i * V + t for i, t in zip(out["step_beam_indices"], out["step_beam_tokens"]))
    out["beam_scores"] = out["step_beam_scores"][-1] if draft_len else None
    return out


@Timer()
def one_step_beam_search(model, inputs: Dict, beam_size: int, beam_scores=None, beam_sequence=None,
                         logits_processor=None, logits_warper=None) -> Dict:
    """One beam-search step on the session's current beams (beamSD.py:40-106). `model` selects the target
    (inputs["_target_model"] is model) or the draft."""
    sess: Session = inputs["_session"]
    which = 0 if model is inputs.get("_target_model") else 1
    sess.step(which, beam_size)
    out = dict(inputs)
    if inputs.get("_trace"):
        r = sess.result()
        out.update(beam_tokens=torch.from_numpy(r["tokens"][:, -1].astype("int64")), beam_scores=torch.from_numpy(r["scores"]))
    return out


def _draft_beam_search(model, inputs: Dict, draft_len: int, beam_size: int, beam_scores=None, beam_sequence=None,
                       logits_processor=None, logits_warper=None) -> Dict:
    sess: Session = inputs["_session"]
    if beam_size != sess.N:
        raise ValueError(f"draft beam_size {beam_size} != session N {sess.N}")
    sess.draft(draft_len)
    out = {"_session": sess, "draft_len": draft_len}
    if inputs.get("_trace"):
        out.update(_levels_dict(sess, draft_len))
    return out


@Timer()
def draft_beam_search(*args, **kwargs):
    return _draft_beam_search(*args, **kwargs)


@Timer(sync_cuda=False)
def no_delay_draft_beam_search(*args, **kwargs):
    return _draft_beam_search(*args, **kwargs)


def _target_beam_search(model, inputs: Dict, draft_outputs: Dict, beam_size: int, beam_scores=None, beam_sequence=None,
                        logits_processor=None) -> Dict:
    sess: Session = inputs["_session"]
    sess.target_forward(draft_outputs["draft_len"])
    out = {"_session": sess}
    if inputs.get("_trace"):
        rows = sess.info()[6]
        out["next_token_scores"] = torch.from_numpy(sess.logits(0, rows))
    return out


@Timer()
def target_beam_search(*args, **kwargs):
    return _target_beam_search(*args, **kwargs)


@Timer(sync_cuda=False)
def no_delay_target_beam_search(*args, **kwargs):
    return _target_beam_search(*args, **kwargs)


@Timer()
def verify(target_model, target_model_inputs: Dict, draft_outputs: Dict, target_outputs: Dict, draft_beam_size: int,
           beam_size: int, beam_scores=None, beam_sequence=None, logits_processor=None, logits_warper=None) -> Dict:
    """verify (beamSD.py:242-456) = kernel (b) + kernel (c): AtSpeed-S strict top-K (greedy branch) or, with
    generation_config.do_sample, AtSpeed-R relaxed acceptance (sampling branch)."""
    sess: Session = target_model_inputs["_session"]
    n_matches = sess.verify(draft_outputs["draft_len"])
    out = {"n_matches": n_matches, "target_model_inputs": target_model_inputs, "draft_model_inputs": target_model_inputs}
    if target_model_inputs.get("_trace"):
        out["trace"] = sess.verify_trace()
        r = sess.result()
        out["beam_scores"] = torch.from_numpy(r["scores"])
        out["beam_tokens"] = torch.from_numpy(r["tokens"].astype("int64"))
    return out


# ------------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------------
def _pack(sess: Session, prompt: List[int], res: Dict, device) -> Dict:
    toks = torch.from_numpy(res["tokens"].astype("int64")).to(device)
    p = torch.tensor(prompt, dtype=torch.long, device=device)
    return {"beam_sequence": torch.cat((p[None, :].expand(toks.shape[0], -1), toks), dim=1),
            "beam_scores": torch.from_numpy(res["scores"]).to(device)}


@Timer()
@torch.no_grad()
def BSSD(target_model, draft_model, inputs: Dict, gamma: int, max_new_tokens: int,
         logits_processor=None, prefix_allowed_tokens_fn=None, trace: bool = False, seed: Optional[int] = None) -> Dict:
    """Speculative beam search (beamSD.py:458-542). Same arguments and result keys as the reference.
    generation_config.do_sample selects AtSpeed-R (relaxed acceptance); `seed` pins its noise (default: drawn from
    torch's global generator, so `set_seed` reproduces a run)."""
    if logits_processor:
        raise NotImplementedError("extra logits processors are not supported; pass prefix_allowed_tokens_fn")
    prompt = _prompt_list(inputs)
    sess = get_session(target_model, draft_model, prompt, max_new_tokens, prefix_allowed_tokens_fn)
    K, N = sess.K, sess.N
    dev = sess.device
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []
    state = {"_session": sess, "_target_model": target_model, "_trace": trace}
    with torch.cuda.device(dev):
        used_seed = _reseed(sess, seed)
        sess.begin(prompt)
        done, accept_steps, rounds = 0, [], []
        while done < max_new_tokens:
            draft_len = min(gamma, max_new_tokens - done - 1)                  # beamSD.py:504
            if draft_len == 0:                                                   # beamSD.py:505-509
                sess.step(0, K)
                break
            e = [ev() for _ in range(4)]
            e[0].record()
            d_out = _draft_beam_search(draft_model, state, draft_len, N)
            e[1].record()
            t_out = _target_beam_search(target_model, state, d_out, N)
            e[2].record()
            n_matches = sess.verify(draft_len)
            e[3].record()
            marks.append(e)
            if trace:
                # "beams": the K beams the next round starts from (generated tokens so far, root order, and their scores)
                rounds.append({"draft": d_out, "target": t_out, "verify": sess.verify_trace(), "n_matches": n_matches,
                               "beams": sess.result()})
            done += n_matches + 1
            accept_steps.append(n_matches)
        sess.sort_result()                                                       # beamSD.py:529-531 (sampling only)
        res = sess.result()
        torch.cuda.current_stream(dev).synchronize()
    out = _pack(sess, prompt, res, dev)
    n_run = len(accept_steps)
    total = sum(accept_steps)
    tc = [sum(m[i].elapsed_time(m[i + 1]) for m in marks) / 1e3 for i in range(3)]
    out.update({"n_run": n_run, "total_accept_steps": total, "total_accept_tokens": total * K,
                "ave_accept_tokens": total * K / n_run if n_run else 0.0,
                "draft_time_cost": tc[0], "target_time_cost": tc[1], "verify_time_cost": tc[2],
                "accept_steps": accept_steps})
    if trace:
        out["rounds"] = rounds
    if used_seed is not None:
        out["seed"] = used_seed
    return out


@Timer()
@torch.no_grad()
def target_generate(model, inputs: Dict, max_new_tokens: int, logits_processor=None,
                    prefix_allowed_tokens_fn=None, seed: Optional[int] = None) -> Dict:
    """Plain tree-mask beam search on the target (beamSD.py:544-595)."""
    prompt = _prompt_list(inputs)
    sess = get_session(model, None, prompt, max_new_tokens, prefix_allowed_tokens_fn)
    _reseed(sess, seed)
    res = sess.target_generate(prompt)
    return _pack(sess, prompt, res, sess.device)
