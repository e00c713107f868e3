/*
 * atspeed.h -- C ABI of libatspeed_b200.so: the B200-native replacement for the hot path of
 * Linxyhaha/AtSpeed, i.e. the speculative beam-search draft/verify loop of code/beamSD.py.
 *
 * The reference is pure Python: there is no FFI in it to mirror.  Each entry point below therefore
 * cites the reference *function* (file:line relative to /root/reference) whose work it replaces; the
 * Python host module atspeed_b200/beamSD.py keeps the reference's names and signatures on top of it
 * (see INTEGRATION.md for the ctypes binding and the one-line change in inference.py).
 *
 * Conventions
 *   - every pointer is a plain device or host address as stated; no torch / C++ types cross the ABI;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); the only calls
 *     that synchronise are the ones that return host values (documented per function);
 *   - buffers are owned by the caller (PyTorch allocates them); the library owns only its handle and a
 *     few bytes of pinned host memory inside it;
 *   - return value 0 = success, negative = error; atspeed_last_error() gives the text.  Nothing throws,
 *     nothing calls exit();
 *   - thread-safety: calls on distinct sessions are independent; one session must be driven from one
 *     thread at a time.
 */
#ifndef ATSPEED_H
#define ATSPEED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATSPEED_ABI_VERSION 2

/* limits compiled into the library */
#define ATSPEED_MAX_K 32          /* target beams  (reference: run_beam_sizes 10/20, code/script/inference.sh:16) */
#define ATSPEED_MAX_N 64          /* draft beams   (reference: draft_beam_size 40, code/script/inference.sh:15)   */
#define ATSPEED_MAX_NEW_TOKENS 6  /* reference uses 4 (code/inference.py:147) */
#define ATSPEED_VIS_WORDS 16

const char* atspeed_last_error(void);
int atspeed_abi_version(void);
/* Diagnostics (nothing like it in the reference).  With ATSPEED_GEMM_TRACE=1 in the environment every GEMM CTA records its
 * progress in mapped host memory: 16 words per CTA {launch seq, kernel<<24|cluster rank<<16|SM id, CTA phase, TMEM phase,
 * producer unit, MMA unit, epilogue segment, grid<<16|T, ...}; this copies up to max_words of it, returns the words copied
 * (0 = tracing off).  Readable while a launch is stuck.  Every mbarrier wait of the library is bounded separately
 * (ATSPEED_SPIN_LIMIT_MS, default 4000): an expired wait traps and atspeed_last_error() names kernel/CTA/role/barrier. */
int atspeed_debug_gemm_trace(uint32_t* out, int32_t max_words);
/* Diagnostics / CPU tests: the cohort scheduler's packing of one step's target forwards.  Item i needs T[i] tokens and R[i]
 * logit rows and has been held back waited[i] steps; pack_of[i] receives its pack (packs numbered in the order they are
 * opened, best-fit decreasing), run_now[b] whether pack b runs in this step (>= 7/8 full, or deferral off, or no more work can
 * arrive, or an item has waited twice, or nothing else would run), *n_packs the pack count.  Pure host arithmetic. */
int atspeed_debug_plan_packs(const int32_t* T, const int32_t* R, const int32_t* waited, int32_t n, int32_t T_max, int32_t R_max,
                             int32_t defer, int32_t no_more_work, int32_t* pack_of, uint8_t* run_now, int32_t* n_packs);
/* Diagnostics: microseconds per launch of one row-wise consumer kernel of the forward (kind 0 RoPE + KV append, 1 SiLU * up,
 * 2 residual + RMSNorm) at T tokens with every output column held in `slices` fp32 partial-sum slices; inputs are cycled
 * through buffers larger than L2.  tools/rowwise_bench.py prints GB/s against MEASURED_PEAKS.json. */
int atspeed_debug_rowwise_us(int32_t kind, int32_t T, int32_t hidden, int32_t mlp, int32_t n_heads, int32_t slices, int32_t iters,
                             float* us_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model + session
 * ---------------------------------------------------------------------------------------------- */

/* A LLaMA decoder as laid out by HF `LlamaForCausalLM.state_dict()` (what `model(**inputs)` runs at
 * code/beamSD.py:52,221): row-major [out_features, in_features] tensors on the device, all bf16 (production: tcgen05
 * GEMMs, bf16 activations) or -- weights_f32 = 1 -- all fp32: the exact-parity mode (plain fp32 SIMT kernels,
 * csrc/forward_f32.cu) that reproduces the reference's ranked lists bit for bit on the small parity configurations. */
typedef struct atspeed_model_desc {
    int32_t vocab, hidden, n_layers, n_heads, head_dim, mlp;
    float rms_eps;
    const void* embed;        /* [vocab, hidden]  model.embed_tokens.weight */
    const void* final_norm;   /* [hidden]         model.norm.weight         */
    const void* lm_head;      /* [vocab, hidden]  lm_head.weight            */
    /* HOST array of n_layers*9 DEVICE pointers, per layer in this order:
     * q_proj, k_proj, v_proj, o_proj, gate_proj, up_proj, down_proj, input_layernorm, post_attention_layernorm */
    const void* const* layer_weights;
    const float* rope_cos;    /* [max_pos, head_dim/2] fp32, bf16-rounded values (LlamaRotaryEmbedding) */
    const float* rope_sin;
    int32_t max_pos;
    int32_t weights_f32;      /* 0: every weight pointer is bf16; 1: every weight pointer is fp32 (rope tables unrounded) */
} atspeed_model_desc;

/* The compiled constraint (code/generation_trie.py Trie / code/data.py:84-104 positional fn): CSR child table. */
typedef struct atspeed_trie_desc {
    const int32_t* child_off;   /* [n_nodes + 1] device */
    const int32_t* child_tok;   /* [n_edges]     device, ascending within a node */
    const int32_t* child_node;  /* [n_edges]     device, -1 beyond the compiled depth */
    int32_t n_nodes, n_edges;
} atspeed_trie_desc;

typedef struct atspeed_config {
    int32_t K;                /* target.generation_config.num_beams  (code/beamSD.py:482) */
    int32_t N;                /* draft.generation_config.num_beams   (code/beamSD.py:483) */
    int32_t max_new_tokens;   /* code/inference.py:147 */
    int32_t max_prompt;       /* longest prompt the session must hold */
    int32_t num_sms;          /* 0 = query the device */
    /* AtSpeed-R, relaxed (sampling) acceptance -- generation_config.do_sample / top_k / temperature as read at
     * code/beamSD.py:53,255,479-481 (set by code/inference.py:149-150).  do_sample = 0: AtSpeed-S strict top-K. */
    int32_t do_sample;
    int32_t top_k;            /* TopKLogitsWarper k (transformers 4.41 default 50); 1..64 required when do_sample */
    float temperature;        /* TemperatureLogitsWarper; 1.0 = off */
    uint64_t seed;            /* key of the counter-based noise (see atspeed_session_set_seed) */
    /* Cohort mode (atspeed_bssd_batch): up to max_users (1..16) independent searches advance together, their trees packed
     * into the same forwards (<= 512 tokens each) so the weights are streamed once per step for all of them.  0 or 1 =
     * single-user session.  Each user keeps its own KV caches and beam tree; results per user are those of a
     * single-user session.  bf16 models only. */
    int32_t max_users;
    int32_t cohort_tokens;    /* most tokens a cohort forward may pack (256..512 = the GEMM's limit; 0 = 512) */
} atspeed_config;

typedef struct atspeed_session atspeed_session;

/* Bytes of device workspace a session needs (activations, KV caches, beam-tree state, logits). */
int atspeed_session_workspace_bytes(const atspeed_model_desc* target, const atspeed_model_desc* draft,
                                    const atspeed_config* cfg, size_t* bytes);
/* `draft` may be NULL (target_generate only).  `workspace` is a device buffer of at least the size above,
 * 1024-byte aligned, that stays alive until atspeed_session_destroy. */
int atspeed_session_create(const atspeed_model_desc* target, const atspeed_model_desc* draft,
                           const atspeed_config* cfg, const atspeed_trie_desc* trie, void* workspace,
                           size_t workspace_bytes, atspeed_session** out);
int atspeed_session_destroy(atspeed_session* s);

/* ------------------------------------------------------------------------------------------------
 * The hot path, stage by stage (mirrors BSSD's loop body, code/beamSD.py:503-526)
 * ---------------------------------------------------------------------------------------------- */

/* Start a user: copies the prompt ids (HOST int32[P], the `inputs["input_ids"]` of code/beamSD.py:486) to the
 * device and resets the beam tree (beam_scores = 0, one root; code/beamSD.py:487-499). Asynchronous. */
int atspeed_session_begin(atspeed_session* s, const int32_t* prompt_host, int32_t P, void* stream);

/* draft_beam_search (code/beamSD.py:108-179): `draft_len` beam-search steps of width N on the draft model,
 * each = forward + kernel (a) + merge.  Asynchronous. */
int atspeed_session_draft(atspeed_session* s, int32_t draft_len, void* stream);

/* target_beam_search (code/beamSD.py:190-232): ONE target forward over roots + all draft levels with the
 * tree mask, then kernel (a) with B = K on every row of interest.  Asynchronous. */
int atspeed_session_target(atspeed_session* s, int32_t draft_len, void* stream);

/* verify (code/beamSD.py:242-456): kernel (b) -- the greedy branch = AtSpeed-S strict top-K, or with
 * cfg.do_sample the sampling branch = AtSpeed-R relaxed acceptance (:293-321,332-369) -- then kernel (c) on both
 * caches.  Writes the accepted length to *n_matches_host after synchronising the stream. */
int atspeed_session_verify(atspeed_session* s, int32_t draft_len, int32_t* n_matches_host, void* stream);

/* one_step_beam_search (code/beamSD.py:40-106) on the current beams: model 0 = target, 1 = draft; the new
 * beams become level `+1` of the tree.  Used for the final step (code/beamSD.py:505-509) and by
 * atspeed_target_generate.  Asynchronous. */
int atspeed_session_step(atspeed_session* s, int32_t model, int32_t width, void* stream);

/* Sampling mode only (no-op otherwise): order the current beams by score, descending -- the final
 * `beam_scores.sort(descending=True)` of code/beamSD.py:529-531.  Call once, after the last step.  Asynchronous. */
int atspeed_session_sort_result(atspeed_session* s, void* stream);

/* Final beams: tokens_host int32[K * max_new_tokens] (generated suffix per beam, score-descending),
 * scores_host float[K], *count = beams returned.  Synchronises the stream. */
int atspeed_session_result(atspeed_session* s, int32_t* tokens_host, float* scores_host, int32_t* count,
                           void* stream);

typedef struct atspeed_stats {
    int32_t n_run;                /* rounds                      (code/beamSD.py:527) */
    int32_t total_accept_steps;   /* sum of n_matches            (code/beamSD.py:528) */
    int32_t accept_steps[8];      /* n_matches per round */
    int32_t target_forwards, draft_forwards;
    int32_t kernel_launches;      /* kernels of this library launched for the call */
} atspeed_stats;

/* BSSD (code/beamSD.py:458-542), whole loop for one user (prompt on the HOST).  Synchronises once per round
 * (the 4-byte n_matches) and once for the result. */
int atspeed_bssd(atspeed_session* s, const int32_t* prompt_host, int32_t P, int32_t gamma, int32_t* tokens_host,
                 float* scores_host, int32_t* count, atspeed_stats* stats, void* stream);

/* BSSD for n_users prompts at once (cohort mode, cfg.max_users > 1): a host-side scheduler keeps up to max_users
 * searches in flight, batches every draft step / target verify forward / final step of the users that are ready into one
 * forward of at most cfg.cohort_tokens tokens, and runs kernels (a), (b), (c) for all of them in single launches.  Users finish in
 * 1..4 rounds independently; a finished user's slot is refilled from the remaining prompts.
 *   prompts_host : int32, all prompts concatenated;  prompt_lens int32[n_users]
 *   tokens_host  : int32[n_users][K * max_new_tokens], scores_host float[n_users][K], counts int32[n_users]
 *   stats        : atspeed_stats[n_users] (may be NULL)
 * One stream synchronisation per scheduler step (the accepted lengths of the users verified in it). */
int atspeed_bssd_batch(atspeed_session* s, int32_t n_users, const int32_t* prompts_host, const int32_t* prompt_lens,
                       int32_t gamma, int32_t* tokens_host, float* scores_host, int32_t* counts, atspeed_stats* stats,
                       void* stream);

/* The same with the concatenated prompts resident in HBM (DEVICE int32; lengths stay on the host: they size the
 * forwards) and the results left there: tokens_dev int32[n_users][K][6], scores_dev float[n_users][K], record i = prompt
 * i.  This is what bench.py times as `value`. */
int atspeed_bssd_batch_device(atspeed_session* s, int32_t n_users, const int32_t* prompts_dev,
                              const int32_t* prompt_lens_host, int32_t gamma, int32_t* tokens_dev, float* scores_dev,
                              atspeed_stats* stats, void* stream);

/* Same loop with the prompt already resident in HBM (DEVICE int32[P]) and the result left on the device:
 * tokens_dev int32[K][6] (generated suffix per beam, row stride 6 = ATSPEED_MAX_NEW_TOKENS), scores_dev float[K].
 * The only host synchronisation is the per-round n_matches.  This is what bench.py times as `value`. */
int atspeed_bssd_device(atspeed_session* s, const int32_t* prompt_dev, int32_t P, int32_t gamma, int32_t* tokens_dev,
                        float* scores_dev, atspeed_stats* stats, void* stream);
int atspeed_session_begin_device(atspeed_session* s, const int32_t* prompt_dev, int32_t P, void* stream);
int atspeed_session_result_device(atspeed_session* s, int32_t* tokens_dev, float* scores_dev, void* stream);

/* target_generate (code/beamSD.py:544-595): plain tree-mask beam search on the target. */
int atspeed_target_generate(atspeed_session* s, const int32_t* prompt_host, int32_t P, int32_t* tokens_host,
                            float* scores_host, int32_t* count, atspeed_stats* stats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Randomness of the sampling mode.  The reference draws from torch generators at five sites per round
 * (code/beamSD.py:72-74,304,336,343,363).  Here every draw is Philox4x32-10(key = seed, counter = (index, stream)):
 * `stream` = atspeed_noise_stream(user_seq, round, level, site) with site 0 draft multinomial, 1 acceptance
 * uniforms, 2 random-K-of-accepted keys, 3 residual multinomial, 4 bonus multinomial, 5 plain target step;
 * `index` = position in the reference's flat [n_prev * V] space (sites 0,3,4,5) or the draft pick position (1,2).
 * A multinomial without replacement over p is the top-n of p / Exp(1) noise, as in ATen.
 * ---------------------------------------------------------------------------------------------- */
/* Cohort sessions only (no reference counterpart): the n tokens every prompt of the coming atspeed_bssd_batch* calls starts
 * with (the instruction template).  Their K/V rows are computed once, here, and copied into each user's caches on admission;
 * the users' forwards then skip them.  A prompt that does not start with them makes atspeed_bssd_batch* fail.  n = 0: off. */
int atspeed_session_set_shared_prefix(atspeed_session* s, const int32_t* prefix_host, int32_t n, void* stream);
/* Re-key the session: the next atspeed_session_begin* uses (seed, user_seq), the one after (seed, user_seq + 1), ... */
int atspeed_session_set_seed(atspeed_session* s, uint64_t seed, uint64_t user_seq);
uint64_t atspeed_noise_stream(uint64_t user_seq, uint32_t round, uint32_t level, uint32_t site);
/* The raw 32 random bits of (seed, stream, index), computed on the HOST by the same function the kernels inline
 * (csrc/noise.cuh): lets CPU-only tests pin the generator against the published Philox4x32-10 test vectors. */
uint32_t atspeed_noise_host_u32(uint64_t seed, uint64_t stream, uint32_t index);
/* out_dev[i] = noise(seed, stream, i), i < n: kind 0 raw uint32, 1 uniform (0,1) float, 2 Exp(1) float. Asynchronous.
 * Tests replay these exact numbers into the CPU oracle. */
int atspeed_noise_fill(uint64_t seed, uint64_t stream, int32_t kind, int32_t n, void* out_dev, void* stream_handle);

/* ------------------------------------------------------------------------------------------------
 * Introspection for parity tests (reads device state back; synchronises)
 * ---------------------------------------------------------------------------------------------- */
enum atspeed_field {
    ATSPEED_F_LEVEL_CNT = 0,     /* int32[5]                                   */
    ATSPEED_F_LEVEL_TOK = 1,     /* int32[5][64]   step_beam_tokens            */
    ATSPEED_F_LEVEL_PARENT = 2,  /* int32[5][64]   step_beam_indices           */
    ATSPEED_F_LEVEL_SCORE = 3,   /* float[5][64]   beam_scores of each level   */
    ATSPEED_F_SCALARS = 4,       /* int32[16]                                  */
    ATSPEED_F_PICK_PARENT = 5,   /* int32[5][32]   verify: target picks per level */
    ATSPEED_F_PICK_TOK = 6,
    ATSPEED_F_PICK_SCORE = 7,    /* float[5][32] */
    ATSPEED_F_HIT_POS = 8,       /* int32[5][32]   draft position of each pick, -1 = miss */
    ATSPEED_F_NPICK = 9,         /* int32[5] */
    ATSPEED_F_LOGITS_TARGET = 10,/* float[rows][ld]: last target logits (rows, ld via atspeed_session_info) */
    ATSPEED_F_LOGITS_DRAFT = 11,
    ATSPEED_F_ROW_NODE = 12,     /* int32[R_max]   trie node of each logits row of the last batch */
    ATSPEED_F_LEVEL_NODE = 13,   /* int32[5][64] */
    ATSPEED_F_TR_ACC = 14,       /* int32[5][64]   relaxed verify: acceptance flag of each draft pick per level */
    ATSPEED_F_LSE_Q = 15         /* float[5]       log-normaliser of the draft's flat softmax per step */
};
int atspeed_session_read(atspeed_session* s, int32_t field, void* host_dst, size_t bytes, void* stream);
/* info[0]=logits ld, [1]=R_max, [2]=T_max, [3]=S_max(target), [4]=A_cap, [5]=kernel launches so far,
 * [6]=rows of the last target batch, [7]=rows of the last draft batch.  The sampling mode's warped candidate
 * count per row B = max(top_k, 2 if K > 1 else 1) is returned by atspeed_session_sample_width. */
int atspeed_session_info(atspeed_session* s, int64_t* info8);
int atspeed_session_sample_width(atspeed_session* s);

/* Per-launch CUDA-event timing (the reference's `Timer` blocks, code/beamSD.py:12-37,51,60,220,276, without the
 * forced device syncs): when enabled every kernel launch is bracketed by two events on the caller's stream.
 * profile_read synchronises and returns, per category {0 gemm, 1 attention, 2 row-wise, 3 kernel (a), 4 beam/verify
 * kernel (b), 5 kernel (c)}: total milliseconds, launch count, algorithmic bytes and (GEMM only; may be NULL) floating-point
 * operations, then resets the counters. */
int atspeed_session_profile(atspeed_session* s, int32_t enable);
int atspeed_session_profile_read(atspeed_session* s, double* ms6, int64_t* count6, double* bytes6, double* flops6,
                                 void* stream);

/* Run one forward of model `model` (0 target, 1 draft) on an explicit batch (all DEVICE arrays): used by the
 * forward parity tests.  tok/pos/slot/prefix_len int32[T], vis uint32[T][16] relative to slot `vis_base`,
 * rows_idx int32[R].  Logits land in the session's logits buffer (ATSPEED_F_LOGITS_*). Asynchronous. */
int atspeed_session_forward_raw(atspeed_session* s, int32_t model, const int32_t* tok, const int32_t* pos,
                                const int32_t* slot, const int32_t* prefix_len, const uint32_t* vis, int32_t vis_base,
                                int32_t T, int32_t S, const int32_t* rows_idx, int32_t R, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone kernels (unit parity tests, roofline benches, drop-in use)
 * ---------------------------------------------------------------------------------------------- */

/* Kernel (a): F.log_softmax + PrefixConstrainedLogitsProcessor + per-row top-B (code/beamSD.py:58-78,285-325).
 * logits: DEVICE [rows, ld], fp32 (logits_bf16 = 0) or bf16 (1); row_node int32[rows] (trie node per row, <0 =
 * skip); n_rows_dev optional DEVICE scalar limiting the active rows.  Outputs (DEVICE): cand_tok/cand_edge
 * int32[rows*B], cand_logp float[rows*B] sorted by (logp desc, token asc), cand_cnt int32[rows], lse float[rows]. */
int atspeed_mask_logsoftmax_topk(const void* logits, int32_t logits_bf16, int32_t rows, int32_t V, int64_t ld,
                                 const int32_t* row_node, const int32_t* n_rows_dev, const atspeed_trie_desc* trie,
                                 int32_t B, int32_t* cand_tok, int32_t* cand_edge, float* cand_logp, int32_t* cand_cnt,
                                 float* lse, void* stream);

/* Kernel (c): row gather with TMA bulk copies.  For p < n_planes, i < rows: copy row_bytes from
 * src_base + p*src_plane_stride + src_rows[i]*row_bytes to dst_base + p*dst_plane_stride + dst_rows[i]*row_bytes.
 * Replaces the KV slicing/re-copy of code/beamSD.py:418-429 and HF's index_select cache reorder. */
int atspeed_kv_gather(const void* src_base, void* dst_base, int64_t src_plane_stride, int64_t dst_plane_stride,
                      int32_t n_planes, int32_t row_bytes, const int32_t* src_rows, const int32_t* dst_rows,
                      const int32_t* n_rows_dev, int32_t rows, void* stream);

/* The tcgen05 GEMM of the forward (csrc/gemm.cu): y[t][colbase_i + n] = sum_k x[t][k] * w_i[n][k].
 * x bf16 [T, K]; up to three weights w_i bf16 [rows_i, K] sharing x.  The kernel leaves fp32 partial-sum slices in
 * `scratch` (atspeed_gemm_scratch_bytes; tiles cut along K across its persistent CTAs); when `out` is non-NULL the
 * slices are then reduced in fixed order into out fp32 [T][ldo] (in the forward that reduction is fused into the
 * consuming row-wise kernel). */
/* Host-side work decomposition of one GEMM launch (no GPU needed; tests/test_gemm_plan.py): info16 = {BM, KB,
 * total_tiles, units_per_cta U, grid, max_slices, stages, tmem_cols, n_bufs, T_pad, tiles0, tiles1, tiles2, 0, 0, 0};
 * slices_of_col (optional, int32[rows0+rows1+rows2]) receives the number of partial-sum slices the consumers add for
 * each output column (the device-side SplitMap arithmetic evaluated on the host). */
int atspeed_gemm_plan(int32_t T, int32_t K, int32_t rows0, int32_t rows1, int32_t rows2, int32_t num_sms, int32_t allow_cut,
                      int32_t* info16, int32_t* slices_of_col);
int atspeed_gemm_scratch_bytes(int32_t T, int32_t K, int32_t rows0, int32_t rows1, int32_t rows2, size_t* bytes);
int atspeed_gemm_bf16(const void* x, int32_t T, int32_t K, const void* w0, int32_t rows0, const void* w1, int32_t rows1,
                      const void* w2, int32_t rows2, float* scratch, float* out, int32_t ldo, void* stream);

/* Tree attention of the forward (see csrc/attention.cu). q/out bf16 [T, n_heads*head_dim]; caches bf16
 * [S, n_heads*head_dim]; prefix_len int32[T]; vis uint32[T][16] relative to vis_base. */
int atspeed_tree_attention(const void* q, const void* kcache, const void* vcache, const int32_t* prefix_len,
                           const uint32_t* vis, int32_t vis_base, int32_t T, int32_t S, int32_t n_heads,
                           int32_t head_dim, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side prompt builder (the step in front of the path; reference code/data.py:232-263 _process_test_data +
 * code/collator.py:50-75 TestCollator, which tokenise "<template> item, item, ... <template> Response:" on the host).
 * A prompt's token ids are a pure function of the user's history item ids: [BOS] prefix, then 4 code-token ids per item with
 * the separator run between items (alternating between its one- and two-piece tokenisation), suffix, the "Response:" run.
 *   ids = prefix | suffix | resp | sep_even | sep_odd  (concatenated, in this order)
 * For user i the kernel reads hist_len[i] item ids at hist_items + hist_begin[i] and writes its prompt at
 * prompts_dev + out_off[i] (the concatenated layout atspeed_bssd_batch_device consumes).  The prompt length is
 * 1 + n_prefix + 4h + ceil((h-1)/2)*n_sep_even + floor((h-1)/2)*n_sep_odd + n_suffix + n_resp, computed by the caller.
 * Asynchronous on `stream`. */
#define ATSPEED_PROMPT_TEMPLATE_IDS 120
typedef struct atspeed_prompt_template {
    int32_t bos, n_prefix, n_suffix, n_resp, n_sep_even, n_sep_odd;
    int32_t ids[ATSPEED_PROMPT_TEMPLATE_IDS];
} atspeed_prompt_template;
int atspeed_build_prompts(const int32_t* item_tok_dev, const int32_t* hist_items_dev, const int64_t* hist_begin_dev,
                          const int32_t* hist_len_dev, const int64_t* out_off_dev, int32_t n_users,
                          const atspeed_prompt_template* tmpl, int32_t* prompts_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATSPEED_H */
