"""ctypes binding of libatspeed_b200.so (include/atspeed.h). Fails loudly when the library is missing --
there is no CPU or eager fallback behind this module."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libatspeed_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)


class ModelDesc(C.Structure):
    _fields_ = [("vocab", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32),
                ("head_dim", C.c_int32), ("mlp", C.c_int32), ("rms_eps", C.c_float),
                ("embed", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p),
                ("layer_weights", C.POINTER(C.c_void_p)),
                ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("max_pos", C.c_int32), ("weights_f32", C.c_int32)]


class TrieDesc(C.Structure):
    _fields_ = [("child_off", C.c_void_p), ("child_tok", C.c_void_p), ("child_node", C.c_void_p),
                ("n_nodes", C.c_int32), ("n_edges", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("max_new_tokens", C.c_int32), ("max_prompt", C.c_int32),
                ("num_sms", C.c_int32), ("do_sample", C.c_int32), ("top_k", C.c_int32), ("temperature", C.c_float),
                ("seed", C.c_uint64), ("max_users", C.c_int32), ("cohort_tokens", C.c_int32)]


class PromptTemplate(C.Structure):
    _fields_ = [("bos", C.c_int32), ("n_prefix", C.c_int32), ("n_suffix", C.c_int32), ("n_resp", C.c_int32),
                ("n_sep_even", C.c_int32), ("n_sep_odd", C.c_int32), ("ids", C.c_int32 * 120)]


class Stats(C.Structure):
    _fields_ = [("n_run", C.c_int32), ("total_accept_steps", C.c_int32), ("accept_steps", C.c_int32 * 8),
                ("target_forwards", C.c_int32), ("draft_forwards", C.c_int32), ("kernel_launches", C.c_int32)]


# every symbol include/atspeed.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "atspeed_last_error": (C.c_char_p, []),
    "atspeed_abi_version": (C.c_int, []),
    "atspeed_debug_gemm_trace": (C.c_int, [C.POINTER(C.c_uint32), C.c_int32]),
    "atspeed_debug_rowwise_us": (C.c_int, [C.c_int32] * 7 + [c_f32p, C.c_void_p]),
    "atspeed_debug_plan_packs": (C.c_int, [c_i32p, c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i32p,
                                           C.POINTER(C.c_uint8), c_i32p]),
    "atspeed_session_workspace_bytes": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(ModelDesc), C.POINTER(Config),
                                                  C.POINTER(C.c_size_t)]),
    "atspeed_session_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(ModelDesc), C.POINTER(Config),
                                         C.POINTER(TrieDesc), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "atspeed_session_destroy": (C.c_int, [C.c_void_p]),
    "atspeed_session_begin": (C.c_int, [C.c_void_p, c_i32p, C.c_int32, C.c_void_p]),
    "atspeed_session_draft": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "atspeed_session_target": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "atspeed_session_verify": (C.c_int, [C.c_void_p, C.c_int32, c_i32p, C.c_void_p]),
    "atspeed_session_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "atspeed_session_sort_result": (C.c_int, [C.c_void_p, C.c_void_p]),
    "atspeed_session_set_seed": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "atspeed_session_set_shared_prefix": (C.c_int, [C.c_void_p, c_i32p, C.c_int32, C.c_void_p]),
    "atspeed_noise_stream": (C.c_uint64, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]),
    "atspeed_noise_host_u32": (C.c_uint32, [C.c_uint64, C.c_uint64, C.c_uint32]),
    "atspeed_noise_fill": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "atspeed_session_sample_width": (C.c_int, [C.c_void_p]),
    "atspeed_session_result": (C.c_int, [C.c_void_p, c_i32p, c_f32p, c_i32p, C.c_void_p]),
    "atspeed_bssd": (C.c_int, [C.c_void_p, c_i32p, C.c_int32, C.c_int32, c_i32p, c_f32p, c_i32p,
                               C.POINTER(Stats), C.c_void_p]),
    "atspeed_bssd_batch": (C.c_int, [C.c_void_p, C.c_int32, c_i32p, c_i32p, C.c_int32, c_i32p, c_f32p, c_i32p,
                                     C.c_void_p, C.c_void_p]),
    "atspeed_bssd_batch_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, c_i32p, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "atspeed_bssd_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.POINTER(Stats), C.c_void_p]),
    "atspeed_session_begin_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "atspeed_session_result_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "atspeed_session_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "atspeed_session_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
    "atspeed_target_generate": (C.c_int, [C.c_void_p, c_i32p, C.c_int32, c_i32p, c_f32p, c_i32p,
                                          C.POINTER(Stats), C.c_void_p]),
    "atspeed_session_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "atspeed_session_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "atspeed_session_forward_raw": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                              C.c_void_p]),
    "atspeed_mask_logsoftmax_topk": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                               C.c_void_p, C.POINTER(TrieDesc), C.c_int32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "atspeed_kv_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "atspeed_gemm_plan": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    c_i32p, c_i32p]),
    "atspeed_gemm_scratch_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "atspeed_gemm_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "atspeed_build_prompts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.POINTER(PromptTemplate), C.c_void_p, C.c_void_p]),
    "atspeed_tree_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
}

# enum atspeed_field
F_LEVEL_CNT, F_LEVEL_TOK, F_LEVEL_PARENT, F_LEVEL_SCORE, F_SCALARS = 0, 1, 2, 3, 4
F_PICK_PARENT, F_PICK_TOK, F_PICK_SCORE, F_HIT_POS, F_NPICK = 5, 6, 7, 8, 9
F_LOGITS_TARGET, F_LOGITS_DRAFT, F_ROW_NODE, F_LEVEL_NODE, F_TR_ACC, F_LSE_Q = 10, 11, 12, 13, 14, 15
SITE_DRAFT, SITE_ACCEPT, SITE_PERM, SITE_RESIDUAL, SITE_BONUS, SITE_STEP = 0, 1, 2, 3, 4, 5
ABI_VERSION = 2
MAX_LEVELS, MAX_BEAMS, MAX_K, MAX_NEW, VIS_WORDS = 5, 64, 32, 6, 16

_lib = None


class AtSpeedError(RuntimeError):
    pass


def load():
    """Load (once) and type the shared library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AtSpeedError(f"{LIB_PATH} is missing: build it with `python -m atspeed_b200.build` "
                           "(there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.atspeed_abi_version() != ABI_VERSION:
        raise AtSpeedError("libatspeed_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise AtSpeedError(f"libatspeed_b200 error {rc}: {load().atspeed_last_error().decode()}")
