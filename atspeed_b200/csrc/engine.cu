// Session orchestration + the C ABI of libatspeed_b200 (include/atspeed.h).
//
// Host-side runtime of the draft/verify loop (reference code/beamSD.py:458-542 `BSSD`, :544-595
// `target_generate`): it only sequences kernel launches on the caller's stream.  All search state lives
// on the device (beam.cuh); the single device->host read per round is the accepted length.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/atspeed.h"
#include "beam.cuh"
#include "kernels.h"

namespace atspeed {

static thread_local char g_err[768] = "";

// ---- bounded mbarrier waits: the mapped diagnostic record (common.cuh) ----
static HangDiag* g_diag_host = nullptr;
static HangDiag* g_diag_dev = nullptr;
static unsigned long long g_spin_limit_ns = 0;

SpinGuard spin_guard() {
    static std::once_flag once;
    std::call_once(once, []() {
        const char* e = getenv("ATSPEED_SPIN_LIMIT_MS");
        long long ms = e ? atoll(e) : 4000;
        if (ms < 0) ms = 0;
        g_spin_limit_ns = static_cast<unsigned long long>(ms) * 1000000ull;
        void* h = nullptr;
        void* d = nullptr;
        if (cudaHostAlloc(&h, sizeof(HangDiag), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
            cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
            memset(h, 0, sizeof(HangDiag));
            g_diag_host = static_cast<HangDiag*>(h);
            g_diag_dev = static_cast<HangDiag*>(d);
        } else {
            cudaGetLastError();      // no record: an expired wait still traps
        }
    });
    return SpinGuard{g_diag_dev, g_spin_limit_ns};
}

static unsigned int* g_trace_host = nullptr;
static unsigned int* g_trace_dev = nullptr;
static std::atomic<unsigned int> g_trace_seq{0};

GemmTrace gemm_trace() {
    static std::once_flag once;
    std::call_once(once, []() {
        const char* e = getenv("ATSPEED_GEMM_TRACE");
        if (!(e && atoi(e) == 1)) return;
        void* h = nullptr;
        void* d = nullptr;
        const size_t bytes = sizeof(unsigned int) * TRACE_WORDS * TRACE_CTAS;
        if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
            cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
            memset(h, 0, bytes);
            g_trace_host = static_cast<unsigned int*>(h);
            g_trace_dev = static_cast<unsigned int*>(d);
        } else {
            cudaGetLastError();
        }
    });
    if (!g_trace_dev) return GemmTrace{nullptr, 0};
    return GemmTrace{g_trace_dev, ++g_trace_seq};
}

void hang_diag_describe(char* buf, size_t n) {
    if (n) buf[0] = 0;
    const volatile HangDiag* d = g_diag_host;
    if (!d || d->flag == 0) return;
    static const char* const kern[] = {"?", "gemm_wx_tcgen05", "gemm_wx_tcgen05_2cta", "kv_gather", "row-wise consumer"};
    static const char* const role[] = {"?", "TMA producer", "MMA issuer", "epilogue", "copy thread"};
    static const char* const bar[] = {"?", "empty", "full", "accum_full", "accum_empty", "row", "partial-sum flag of worker"};
    const unsigned k = d->kernel < 5 ? d->kernel : 0, r = d->role < 5 ? d->role : 0, b = d->barrier < 7 ? d->barrier : 0;
    snprintf(buf, n,
             " [device stall: %s block %u thread %u (%s) waited %.0f ms for %s[%u] parity %u, unit %u of [%u,%u), T=%u%s]",
             kern[k], d->block, d->thread, role[r], d->waited_ns * 1e-6, bar[b], d->index, d->parity, d->unit, d->u_begin,
             d->u_end, d->T, d->flag == 2 ? "" : ", record incomplete");
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    const size_t len = strlen(g_err);
    if (len + 1 < sizeof(g_err)) hang_diag_describe(g_err + len, sizeof(g_err) - len);
}
const char* last_error() { return g_err; }

}  // namespace atspeed

#include "session.h"

using namespace atspeed;

namespace atspeed {

static double gemm_bytes(const GemmWeights& g, int T) {
    double rows = 0;
    for (int i = 0; i < g.n; ++i) rows += g.rows[i];
    // algorithmic traffic: every weight once, the bf16 activations once, a bf16 result once
    return 2.0 * (rows * g.K + static_cast<double>(T) * g.K + static_cast<double>(T) * rows);
}

static double gemm_flops(const GemmWeights& g, int T) {
    double rows = 0;
    for (int i = 0; i < g.n; ++i) rows += g.rows[i];
    return 2.0 * rows * g.K * T;
}

static GemmWeights shape_only(int K, std::initializer_list<int> rows) {
    GemmWeights g;
    memset(&g, 0, sizeof(g));
    g.K = K;
    int col = 0;
    for (int r : rows) { g.rows[g.n] = r; g.colbase[g.n] = col; col += r; ++g.n; }
    return g;
}

// fp32 partial-sum workspace: the widest (slices x T_max x columns) any projection of the model can need
static size_t part_elems(const atspeed_model_desc& d, int T_max, int num_sms) {
    const int HD = d.n_heads * d.head_dim;
    size_t e = 0;
    auto upd = [&](const GemmWeights& g, int cols) {
        const size_t v = static_cast<size_t>(gemm_max_slices(g, T_max, num_sms)) * T_max * cols;
        if (v > e) e = v;
    };
    upd(shape_only(d.hidden, {HD, HD, HD}), 3 * HD);
    upd(shape_only(HD, {d.hidden}), d.hidden);
    upd(shape_only(d.hidden, {d.mlp, d.mlp}), 2 * d.mlp);
    upd(shape_only(d.mlp, {d.hidden}), d.hidden);
    return e;
}

static void carve_model(Carver& c, ModelRT& m, const atspeed_model_desc& d, int T_max, int R_max, int S_max, int num_sms,
                        int n_users) {
    m.d = d;
    m.HD = d.n_heads * d.head_dim;
    m.ldl = (d.vocab + 7) & ~7;
    m.f32 = d.weights_f32 != 0;
    m.elem_bytes = m.f32 ? 4 : 2;
    m.last_rows = 0;
    m.forwards = 0;
    m.kv_plane = static_cast<long long>(S_max) * m.HD;
    if (m.f32) {
        m.fh = c.take<float>(static_cast<size_t>(T_max) * d.hidden);
        m.fx = c.take<float>(static_cast<size_t>(T_max) * d.hidden);
        m.fq = c.take<float>(static_cast<size_t>(T_max) * m.HD);
        m.fa = c.take<float>(static_cast<size_t>(T_max) * m.HD);
        m.fm = c.take<float>(static_cast<size_t>(T_max) * d.mlp);
        m.fxsel = c.take<float>(static_cast<size_t>(R_max) * d.hidden);
        m.fqkv = c.take<float>(static_cast<size_t>(T_max) * 3 * m.HD);
        m.fgu = c.take<float>(static_cast<size_t>(T_max) * 2 * d.mlp);
        m.logits = c.take<float>(static_cast<size_t>(R_max) * m.ldl);
        m.fkv = c.take<float>(static_cast<size_t>(n_users) * d.n_layers * 2 * m.kv_plane);
        return;
    }
    m.h = c.take<__nv_bfloat16>(static_cast<size_t>(T_max) * d.hidden);
    m.x = c.take<__nv_bfloat16>(static_cast<size_t>(T_max) * d.hidden);
    m.q = c.take<__nv_bfloat16>(static_cast<size_t>(T_max) * m.HD);
    m.a = c.take<__nv_bfloat16>(static_cast<size_t>(T_max) * m.HD);
    m.m = c.take<__nv_bfloat16>(static_cast<size_t>(T_max) * d.mlp);
    m.xsel = c.take<__nv_bfloat16>(static_cast<size_t>(R_max) * d.hidden);
    m.part = c.take<float>(part_elems(d, T_max, num_sms));
    m.logits = c.take<float>(static_cast<size_t>(R_max) * m.ldl);
    m.kv = c.take<__nv_bfloat16>(static_cast<size_t>(n_users) * d.n_layers * 2 * m.kv_plane);
}

static void carve_tree(Carver& c, TreeDev& t) {
    t.cnt = c.take<int>(MAX_LEVELS);
    t.tok = c.take<int>(MAX_LEVELS * MAX_BEAMS);
    t.parent = c.take<int>(MAX_LEVELS * MAX_BEAMS);
    t.node = c.take<int>(MAX_LEVELS * MAX_BEAMS);
    t.slot = c.take<int>(MAX_LEVELS * MAX_BEAMS);
    t.score = c.take<float>(MAX_LEVELS * MAX_BEAMS);
    t.gen = c.take<int>(MAX_LEVELS * MAX_BEAMS * MAX_NEW);
    t.vis = c.take<uint32_t>(MAX_LEVELS * MAX_BEAMS * VIS_WORDS);
    t.scal = c.take<int>(SC_COUNT);
    t.tr_pick_parent = c.take<int>(MAX_LEVELS * MAX_K);
    t.tr_pick_tok = c.take<int>(MAX_LEVELS * MAX_K);
    t.tr_pick_score = c.take<float>(MAX_LEVELS * MAX_K);
    t.tr_hit_pos = c.take<int>(MAX_LEVELS * MAX_K);
    t.tr_npick = c.take<int>(MAX_LEVELS);
    t.gather_src = c.take<int>(MAX_LEVELS * MAX_K);
    t.gather_dst = c.take<int>(MAX_LEVELS * MAX_K);
    t.miss_tok = c.take<int>(MAX_K);
    t.miss_pos = c.take<int>(MAX_K);
    t.miss_slot = c.take<int>(MAX_K);
    t.miss_vis = c.take<uint32_t>(MAX_K * VIS_WORDS);
    t.dcand_tok = c.take<int>(MAX_LEVELS * MAX_BEAMS * MAX_BEAMS);
    t.dcand_logp = c.take<float>(MAX_LEVELS * MAX_BEAMS * MAX_BEAMS);
    t.dcand_edge = c.take<int>(MAX_LEVELS * MAX_BEAMS * MAX_BEAMS);
    t.dcand_cnt = c.take<int>(MAX_LEVELS * MAX_BEAMS);
    t.lse_q = c.take<float>(MAX_LEVELS);
    t.tr_acc = c.take<int>(MAX_LEVELS * MAX_BEAMS);
}

static void carve_session(Carver& c, atspeed_session* s, const atspeed_model_desc* target, const atspeed_model_desc* draft) {
    const int U = s->max_users;
    const int UX = U > 1 ? U + 1 : U;               // cohort sessions: one more slot (trees, prompts, KV) for the shared prompt prefix
    carve_model(c, s->tgt, *target, s->T_max, s->R_max, s->S_max, s->num_sms, UX);
    if (draft) carve_model(c, s->dft, *draft, s->T_max, s->R_max, s->S_max, s->num_sms, UX);
    s->kv_user_elems_tgt = static_cast<long long>(target->n_layers) * 2 * s->tgt.kv_plane;
    s->kv_user_elems_dft = draft ? static_cast<long long>(draft->n_layers) * 2 * s->dft.kv_plane : 0;
    s->trees_host.resize(UX);
    for (int u = 0; u < UX; ++u) carve_tree(c, s->trees_host[u]);
    s->tree = s->trees_host[0];                     // the single-user stages drive user slot 0
    s->trees_dev = c.take<TreeDev>(UX);
    BatchDev& b = s->batch;
    b.tok = c.take<int>(s->T_max);
    b.pos = c.take<int>(s->T_max);
    b.slot = c.take<int>(s->T_max);
    b.prefix_len = c.take<int>(s->T_max);
    b.vis = c.take<uint32_t>(static_cast<size_t>(s->T_max) * VIS_WORDS);
    b.tok_user = c.take<int>(s->T_max);
    b.rows_idx = c.take<int>(s->R_max);
    b.row_node = c.take<int>(s->R_max);
    s->cand_tok = c.take<int>(static_cast<size_t>(s->R_max) * MAX_BEAMS);
    s->cand_edge = c.take<int>(static_cast<size_t>(s->R_max) * MAX_BEAMS);
    s->cand_logp = c.take<float>(static_cast<size_t>(s->R_max) * MAX_BEAMS);
    s->cand_cnt = c.take<int>(s->R_max);
    s->lse = c.take<float>(s->R_max);
    s->prompts_dev = c.take<int>(static_cast<size_t>(UX) * s->cfg.max_prompt);
    s->prompt_dev = s->prompts_dev;
    s->prefix_iota_dev = c.take<int>(s->cfg.max_prompt);
    s->prefix_bad_dev = c.take<int>(1);
    s->fused_part = c.take<float>(gemm_fused_part_elems(s->T_max, s->num_sms));
    s->fused_flags = c.take<unsigned int>(1024);
    s->collect_dev = c.take<int>(U * 4);
    s->res_tok_dev = c.take<int>(static_cast<size_t>(U) * MAX_K * MAX_NEW);
    s->res_score_dev = c.take<float>(static_cast<size_t>(U) * MAX_K);
    s->res_cnt_dev = c.take<int>(U);
}

static int check_cfg(const atspeed_model_desc* target, const atspeed_model_desc* draft, const atspeed_config* cfg) {
    ATS_CHECK_ARG(target && cfg, "null target/config");
    ATS_CHECK_ARG(cfg->K >= 1 && cfg->K <= MAX_K, "K=%d outside [1,%d]", cfg->K, MAX_K);
    ATS_CHECK_ARG(cfg->N >= cfg->K && cfg->N <= MAX_BEAMS, "N=%d outside [K,%d]", cfg->N, MAX_BEAMS);
    ATS_CHECK_ARG(cfg->max_new_tokens >= 1 && cfg->max_new_tokens <= MAX_NEW && cfg->max_new_tokens < MAX_LEVELS + 1,
                  "max_new_tokens=%d outside [1,%d]", cfg->max_new_tokens, MAX_LEVELS);
    ATS_CHECK_ARG(cfg->max_prompt >= 1, "max_prompt=%d", cfg->max_prompt);
    ATS_CHECK_ARG(cfg->max_users >= 0 && cfg->max_users <= MAX_USERS, "max_users=%d outside [0,%d]", cfg->max_users, MAX_USERS);
    ATS_CHECK_ARG(cfg->cohort_tokens == 0 || (cfg->cohort_tokens >= 256 && cfg->cohort_tokens <= 512),
                  "cohort_tokens=%d outside [256,512]", cfg->cohort_tokens);
    if (cfg->do_sample) {
        ATS_CHECK_ARG(cfg->top_k >= 1 && cfg->top_k <= MAX_BEAMS,
                      "do_sample needs generation_config.top_k in [1,%d] (transformers 4.41 default 50), got %d", MAX_BEAMS,
                      cfg->top_k);
        ATS_CHECK_ARG(cfg->temperature > 0.f, "temperature=%f", cfg->temperature);
    }
    const int bits = cfg->max_new_tokens * cfg->K + cfg->K + (MAX_LEVELS - 1) * cfg->N + cfg->K;
    ATS_CHECK_ARG(bits <= MAX_TREE_SLOTS, "K=%d N=%d need %d tree slots > %d", cfg->K, cfg->N, bits, MAX_TREE_SLOTS);
    for (const atspeed_model_desc* d : {target, draft}) {
        if (!d) continue;
        if (!d->weights_f32) {
            ATS_CHECK_ARG(d->hidden % 8 == 0 && d->mlp % 8 == 0 && (d->n_heads * d->head_dim) % 8 == 0,
                          "hidden/mlp/heads*head_dim must be multiples of 8");
            ATS_CHECK_ARG(d->head_dim == 16 || d->head_dim == 32 || d->head_dim == 64 || d->head_dim == 128,
                          "head_dim=%d unsupported", d->head_dim);
        } else {
            ATS_CHECK_ARG(d->head_dim % 2 == 0, "head_dim=%d must be even", d->head_dim);
        }
        ATS_CHECK_ARG(d->vocab == target->vocab, "draft and target vocabularies differ");
    }
    return ATS_OK;
}

static void session_dims(atspeed_session* s) {
    const atspeed_config& c = s->cfg;
    const int dl_max = c.max_new_tokens - 1 < MAX_LEVELS - 1 ? c.max_new_tokens - 1 : MAX_LEVELS - 1;
    int T = c.max_prompt + dl_max * c.N;
    if (T < c.max_prompt + c.K) T = c.max_prompt + c.K;   // target_generate-style steps of width K
    if (T > 512) T = 512;
    s->T_max = T;
    s->geom.K = c.K; s->geom.N = c.N; s->geom.A_cap = c.max_new_tokens * c.K; s->geom.V = 0;
    s->R_max = c.K + (MAX_LEVELS - 1) * c.N + 1;
    s->max_users = c.max_users > 1 ? c.max_users : 1;
    if (s->max_users > 1) { s->T_max = c.cohort_tokens > 0 ? c.cohort_tokens : 512; s->R_max = s->T_max; }   // tokens packed per forward
    s->S_max = c.max_prompt + s->geom.A_cap + s->geom.lvl_off(MAX_LEVELS) + c.K;
}

static int build_gemm(GemmWeights& g, int K, std::initializer_list<std::pair<const void*, int>> ws) {
    g.n = 0; g.K = K;
    int col = 0;
    for (auto& w : ws) {
        g.rows[g.n] = w.second;
        g.colbase[g.n] = col;
        ATS_TRY(make_tmap_bf16_kmajor(&g.tmap[g.n], w.first, w.second, K, 128));
        ATS_TRY(make_tmap_bf16_kmajor(&g.tmap256[g.n], w.first, w.second, K, 256));
        ATS_TRY(make_tmap_bf16_kmajor(&g.tmap64[g.n], w.first, w.second, K, 64));
        col += w.second;
        ++g.n;
    }
    return ATS_OK;
}

static int build_model(ModelRT& m, int num_sms) {
    const atspeed_model_desc& d = m.d;
    if (m.f32) return ATS_OK;      // the fp32 parity forward reads d.layer_weights directly (no tensor maps)
    (void)num_sms;
    m.layers.resize(d.n_layers);
    for (int l = 0; l < d.n_layers; ++l) {
        const void* const* w = d.layer_weights + static_cast<size_t>(l) * 9;
        LayerRT& L = m.layers[l];
        ATS_TRY(build_gemm(L.qkv, d.hidden, {{w[0], m.HD}, {w[1], m.HD}, {w[2], m.HD}}));
        ATS_TRY(build_gemm(L.o, m.HD, {{w[3], d.hidden}}));
        ATS_TRY(build_gemm(L.gu, d.hidden, {{w[4], d.mlp}, {w[5], d.mlp}}));
        ATS_TRY(build_gemm(L.down, d.mlp, {{w[6], d.hidden}}));
        L.ln1 = static_cast<const __nv_bfloat16*>(w[7]);
        L.ln2 = static_cast<const __nv_bfloat16*>(w[8]);
    }
    ATS_TRY(build_gemm(m.lm, d.hidden, {{d.lm_head, d.vocab}}));
    return ATS_OK;
}

// fp32 exact-parity forward (csrc/forward_f32.cu): same batch descriptors and KV-slot layout as the bf16 forward
static int forward_f32(atspeed_session* s, ModelRT& m, const BatchDesc& b, int T, int S, const int* rows_idx, int R,
                       cudaStream_t st) {
    const atspeed_model_desc& d = m.d;
    const int HD = m.HD;
    auto W = [&](int l, int j) { return static_cast<const float*>(d.layer_weights[static_cast<size_t>(l) * 9 + j]); };
    PROF(s, CAT_ELEM, 0, f32_embed(static_cast<const float*>(d.embed), b.tok, T, d.hidden, d.vocab, m.fh, st));
    s->launches += 1;
    for (int l = 0; l < d.n_layers; ++l) {
        float* kc = m.fkv + static_cast<long long>(l) * 2 * m.kv_plane;
        float* vc = kc + m.kv_plane;
        PROF(s, CAT_ELEM, 0, f32_rmsnorm(m.fh, W(l, 7), T, d.hidden, d.rms_eps, m.fx, nullptr, st));
        for (int j = 0; j < 3; ++j)
            PROF(s, CAT_GEMM, 0, f32_gemm(m.fx, W(l, j), T, HD, d.hidden, m.fqkv + j * HD, 3 * HD, false, st));
        PROF(s, CAT_ELEM, 0,
             f32_rope_append(m.fqkv, b, T, d.n_heads, d.head_dim, d.rope_cos, d.rope_sin, d.max_pos, m.fq, kc, vc, st));
        PROF(s, CAT_ATTN, 0, f32_tree_attention(m.fq, kc, vc, b, T, S, d.n_heads, d.head_dim, m.fa, st));
        PROF(s, CAT_GEMM, 0, f32_gemm(m.fa, W(l, 3), T, d.hidden, HD, m.fh, d.hidden, true, st));        // h += o_proj(a)
        PROF(s, CAT_ELEM, 0, f32_rmsnorm(m.fh, W(l, 8), T, d.hidden, d.rms_eps, m.fx, nullptr, st));
        PROF(s, CAT_GEMM, 0, f32_gemm(m.fx, W(l, 4), T, d.mlp, d.hidden, m.fgu, 2 * d.mlp, false, st));
        PROF(s, CAT_GEMM, 0, f32_gemm(m.fx, W(l, 5), T, d.mlp, d.hidden, m.fgu + d.mlp, 2 * d.mlp, false, st));
        PROF(s, CAT_ELEM, 0, f32_silu_mul(m.fgu, T, d.mlp, m.fm, st));
        PROF(s, CAT_GEMM, 0, f32_gemm(m.fm, W(l, 6), T, d.hidden, d.mlp, m.fh, d.hidden, true, st));     // h += down(m)
        s->launches += 12;
    }
    PROF(s, CAT_ELEM, 0, f32_rmsnorm(m.fh, static_cast<const float*>(d.final_norm), R, d.hidden, d.rms_eps, m.fxsel, rows_idx, st));
    PROF(s, CAT_GEMM, 0, f32_gemm(m.fxsel, static_cast<const float*>(d.lm_head), R, d.vocab, d.hidden, m.logits, m.ldl, false, st));
    s->launches += 2;
    m.last_rows = R;
    m.forwards++;
    return ATS_OK;
}

// One forward of `m` over the batch in `b` (T tokens; attention scans KV slots [0, S)), logits for R rows.
int forward(atspeed_session* s, ModelRT& m, const BatchDesc& b, int T, int S, const int* rows_idx, int R,
            cudaStream_t st) {
    const atspeed_model_desc& d = m.d;
    ATS_CHECK_ARG(T >= 1 && T <= s->T_max, "forward: T=%d exceeds T_max=%d", T, s->T_max);
    ATS_CHECK_ARG(R >= 1 && R <= s->R_max, "forward: R=%d exceeds R_max=%d", R, s->R_max);
    ATS_CHECK_ARG(S <= s->S_max, "forward: S=%d exceeds S_max=%d", S, s->S_max);
    if (m.f32) return forward_f32(s, m, b, T, S, rows_idx, R, st);
    const auto* embed = static_cast<const __nv_bfloat16*>(d.embed);
    // work decomposition of the four projection shapes at this T (identical for every layer) + activation operand maps
    const LayerRT& L0 = m.layers[0];
    // fused epilogues (opt-in, ATSPEED_FUSED_EPI=1): the q|k|v GEMM applies RoPE and appends to the KV cache, the gate|up GEMM
    // applies SiLU * up; default: fp32 partial-sum slices + row-wise consumer kernels for both
    const bool fused = s->fused;
    static const bool dup_rowwise = []() { const char* e = getenv("ATSPEED_DEBUG_DUP_ROWWISE"); return e && atoi(e) == 1; }();
    GemmPlan p_qkv, p_o, p_gu, p_down, p_lm;
    if (fused) {
        ATS_TRY(gemm_make_plan_fused(L0.qkv, T, s->num_sms, EPI_QKV_ROPE, &p_qkv));
        ATS_TRY(gemm_make_plan_fused(L0.gu, T, s->num_sms, EPI_SILU_MUL, &p_gu));
    } else {
        ATS_TRY(gemm_make_plan(L0.qkv, T, s->num_sms, true, &p_qkv));
        ATS_TRY(gemm_make_plan(L0.gu, T, s->num_sms, true, &p_gu));
    }
    ATS_TRY(gemm_make_plan(L0.o, T, s->num_sms, true, &p_o));
    ATS_TRY(gemm_make_plan(L0.down, T, s->num_sms, true, &p_down));
    ATS_TRY(gemm_make_plan(m.lm, R, s->num_sms, false, &p_lm));
    const SplitMap sm_o = gemm_split_map(L0.o, p_o), sm_down = gemm_split_map(L0.down, p_down);
    SplitMap sm_qkv, sm_gu;
    memset(&sm_qkv, 0, sizeof(sm_qkv));
    memset(&sm_gu, 0, sizeof(sm_gu));
    if (!fused) { sm_qkv = gemm_split_map(L0.qkv, p_qkv); sm_gu = gemm_split_map(L0.gu, p_gu); }
    XMap xm_x, xm_xg, xm_a, xm_m, xm_sel;       // one per consumer: the box of a map depends on the consumer's cluster size
    ATS_TRY(gemm_make_xmap(&xm_x, m.x, T, d.hidden, p_qkv.cluster));
    ATS_TRY(gemm_make_xmap(&xm_xg, m.x, T, d.hidden, p_gu.cluster));
    ATS_TRY(gemm_make_xmap(&xm_a, m.a, T, m.HD, p_o.cluster));
    ATS_TRY(gemm_make_xmap(&xm_m, m.m, T, d.mlp, p_down.cluster));
    ATS_TRY(gemm_make_xmap(&xm_sel, m.xsel, R, d.hidden, p_lm.cluster));
    const int c_qkv = 3 * m.HD, c_gu = 2 * d.mlp;
    OMap om_qkv, om_o, om_gu, om_down, om_lm;
    if (!fused) {
        ATS_TRY(gemm_make_omap(&om_qkv, L0.qkv, m.part, c_qkv, static_cast<long long>(T) * c_qkv, T, p_qkv.max_slices));
        ATS_TRY(gemm_make_omap(&om_gu, L0.gu, m.part, c_gu, static_cast<long long>(T) * c_gu, T, p_gu.max_slices));
    }
    ATS_TRY(gemm_make_omap(&om_o, L0.o, m.part, d.hidden, static_cast<long long>(T) * d.hidden, T, p_o.max_slices));
    FusedEpi e_qkv, e_gu;
    memset(&e_qkv, 0, sizeof(e_qkv));
    memset(&e_gu, 0, sizeof(e_gu));
    if (fused) {
        e_qkv.kind = EPI_QKV_ROPE; e_qkv.part = s->fused_part; e_qkv.flags = s->fused_flags;
        e_qkv.pos = b.pos; e_qkv.slot = b.slot; e_qkv.tok_user = b.tok_user;
        e_qkv.rope_cos = d.rope_cos; e_qkv.rope_sin = d.rope_sin; e_qkv.max_pos = d.max_pos; e_qkv.head_dim = d.head_dim; e_qkv.HD = m.HD;
        e_qkv.qbuf = m.q;
        for (int i = 0; i < MAX_USERS; ++i) e_qkv.kv_off[i] = b.ckv.kv_off[i];
        e_gu.kind = EPI_SILU_MUL; e_gu.part = s->fused_part; e_gu.flags = s->fused_flags; e_gu.m = m.m; e_gu.mlp = d.mlp;
    }
    auto next_epoch = [&]() { if (++s->fused_epoch == 0) ++s->fused_epoch; return s->fused_epoch; };
    ATS_TRY(gemm_make_omap(&om_down, L0.down, m.part, d.hidden, static_cast<long long>(T) * d.hidden, T, p_down.max_slices));
    ATS_TRY(gemm_make_omap(&om_lm, m.lm, m.logits, m.ldl, 0, R, 1));
    PROF(s, CAT_ELEM, 0, embed_rows(embed, b.tok, T, d.hidden, d.vocab, m.h, st));
    PROF(s, CAT_ELEM, 0, rmsnorm_rows(m.h, m.layers[0].ln1, T, d.hidden, d.rms_eps, m.x, nullptr, st));
    s->launches += 2;
    for (int l = 0; l < d.n_layers; ++l) {
        LayerRT& L = m.layers[l];
        __nv_bfloat16* kc = m.kv + static_cast<long long>(l) * 2 * m.kv_plane;
        __nv_bfloat16* vc = kc + m.kv_plane;
        if (fused) {
            e_qkv.kcache = kc; e_qkv.vcache = vc; e_qkv.epoch = next_epoch();
            PROF_GEMM(s, L.qkv, T, gemm_wx_fused(L.qkv, xm_x, p_qkv, e_qkv, st));
        } else {
            PROF_GEMM(s, L.qkv, T, gemm_wx(L.qkv, xm_x, p_qkv, om_qkv, st));
            PROF(s, CAT_ELEM, 0,
                 qkv_rope_append(m.part, sm_qkv, static_cast<long long>(T) * c_qkv, c_qkv, b, T, d.n_heads, d.head_dim,
                                 d.rope_cos, d.rope_sin, d.max_pos, m.q, kc, vc, st));
            if (dup_rowwise)      // timing experiment: both consumers are idempotent, so running them twice only costs time
                PROF(s, CAT_ELEM, 0,
                     qkv_rope_append(m.part, sm_qkv, static_cast<long long>(T) * c_qkv, c_qkv, b, T, d.n_heads, d.head_dim,
                                     d.rope_cos, d.rope_sin, d.max_pos, m.q, kc, vc, st));
        }
        PROF(s, CAT_ATTN, 0, tree_attention(m.q, kc, vc, b, T, S, d.n_heads, d.head_dim, m.a, st));
        PROF_GEMM(s, L.o, T, gemm_wx(L.o, xm_a, p_o, om_o, st));
        PROF(s, CAT_ELEM, 0,
             residual_rmsnorm(m.h, m.part, sm_o, static_cast<long long>(T) * d.hidden, d.hidden, L.ln2, T, d.hidden,
                              d.rms_eps, m.x, st));
        if (fused) {
            e_gu.epoch = next_epoch();
            PROF_GEMM(s, L.gu, T, gemm_wx_fused(L.gu, xm_xg, p_gu, e_gu, st));
        } else {
            PROF_GEMM(s, L.gu, T, gemm_wx(L.gu, xm_xg, p_gu, om_gu, st));
            PROF(s, CAT_ELEM, 0, silu_mul(m.part, sm_gu, static_cast<long long>(T) * c_gu, c_gu, T, d.mlp, m.m, st));
            if (dup_rowwise) PROF(s, CAT_ELEM, 0, silu_mul(m.part, sm_gu, static_cast<long long>(T) * c_gu, c_gu, T, d.mlp, m.m, st));
        }
        PROF_GEMM(s, L.down, T, gemm_wx(L.down, xm_m, p_down, om_down, st));
        const __nv_bfloat16* next_ln = l + 1 < d.n_layers ? m.layers[l + 1].ln1 : nullptr;
        PROF(s, CAT_ELEM, 0,
             residual_rmsnorm(m.h, m.part, sm_down, static_cast<long long>(T) * d.hidden, d.hidden, next_ln, T,
                              d.hidden, d.rms_eps, m.x, st));
        s->launches += fused ? 7 : 9;
    }
    PROF(s, CAT_ELEM, 0,
         rmsnorm_rows(m.h, static_cast<const __nv_bfloat16*>(d.final_norm), R, d.hidden, d.rms_eps, m.xsel, rows_idx, st));
    PROF_GEMM(s, m.lm, R, gemm_wx(m.lm, xm_sel, p_lm, om_lm, st));
    s->launches += 2;
    m.last_rows = R;
    m.forwards++;
    return ATS_OK;
}

static BatchDesc batch_desc(const atspeed_session* s) {
    BatchDesc b;
    memset(&b, 0, sizeof(b));
    b.tok = s->batch.tok; b.pos = s->batch.pos; b.slot = s->batch.slot; b.prefix_len = s->batch.prefix_len;
    b.vis = s->batch.vis; b.vis_base = s->P; b.n_valid = nullptr;
    return b;
}

CandOut shared_cand(atspeed_session* s) { return CandOut{s->cand_tok, s->cand_edge, s->cand_logp, s->cand_cnt}; }
// the draft's step-`level` candidates are kept per level in sampling mode: verify needs q on them
static CandOut draft_level_cand(atspeed_session* s, int level) {
    const size_t o = static_cast<size_t>(level) * MAX_BEAMS * MAX_BEAMS;
    return CandOut{s->tree.dcand_tok + o, s->tree.dcand_edge + o, s->tree.dcand_logp + o, s->tree.dcand_cnt + level * MAX_BEAMS};
}
SampleCfg sample_cfg(const atspeed_session* s) {
    SampleCfg sc;
    sc.B = s->sample_B;
    sc.inv_temp = 1.0f / s->cfg.temperature;
    sc.seed = s->seed;
    sc.stream_base = noise_stream(s->user_seq, static_cast<uint32_t>(s->round), 0, 0);
    return sc;
}

int run_topk(atspeed_session* s, ModelRT& m, int R, int B, const CandOut& o, cudaStream_t st) {
    PROF(s, CAT_TOPK, static_cast<double>(R) * m.d.vocab * 4.0,
         mask_logsoftmax_topk(m.logits, 0, R, m.d.vocab, m.ldl, s->batch.row_node, nullptr, s->trie, B, o.tok, o.edge,
                              o.logp, o.cnt, s->lse, st));
    s->launches += 1;
    return ATS_OK;
}

// one beam-search step of `width` on model m from tree level `level` -> level + 1
// (one_step_beam_search, beamSD.py:40-106).  `first`: the roots are the prompt itself.
static int search_step(atspeed_session* s, ModelRT& m, int level, int width, bool first, bool with_missing,
                       cudaStream_t st) {
    const TreeGeom& g = s->geom;
    const int P = s->P;
    BatchPlan plan;
    memset(&plan, 0, sizeof(plan));
    int T, R, S;
    if (level == 0 && first) {
        plan.with_prompt = 1; plan.l_from = 1; plan.l_to = 0; plan.rows_from = 1; plan.root_row = 1; plan.width = width;
        T = P; R = 1; S = P;
    } else {
        plan.with_missing = with_missing ? 1 : 0;
        plan.l_from = plan.l_to = plan.rows_from = level; plan.width = width;
        const int cap = level == 0 ? g.K : width;
        T = (with_missing ? g.K : 0) + cap; R = cap;
        S = g.tree_slot(P, level, 0) + cap;
    }
    PROF(s, CAT_BEAM, 0, tree_build_batch(s->tree, s->batch, g, plan, s->prompt_dev, P, T, R, st));
    s->launches += 1;
    ATS_TRY(forward(s, m, batch_desc(s), T, S, s->batch.rows_idx, R, st));
    if (s->cfg.do_sample) {
        // beamSD.py:65-74: warp (temperature, top-k) the masked log-probs, then `width` samples without replacement
        const bool is_draft = &m == &s->dft;
        const CandOut o = is_draft ? draft_level_cand(s, level) : shared_cand(s);
        ATS_TRY(run_topk(s, m, R, s->sample_B, o, st));
        PROF(s, CAT_BEAM, 0,
             tree_select_sample(s->tree, g, s->trie, level, 0, o.tok, o.edge, o.logp, o.cnt, width, P, sample_cfg(s),
                                is_draft ? SITE_DRAFT : SITE_STEP, st));
    } else {
        ATS_TRY(run_topk(s, m, R, width, shared_cand(s), st));
        PROF(s, CAT_BEAM, 0,
             tree_select(s->tree, g, s->trie, level, 0, width, s->cand_tok, s->cand_edge, s->cand_logp, s->cand_cnt, width, P, st));
    }
    s->launches += 1;
    return ATS_OK;
}

}  // namespace atspeed

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* atspeed_last_error(void) { return atspeed::last_error(); }

int atspeed_debug_gemm_trace(uint32_t* out, int32_t max_words) {
    if (!atspeed::g_trace_host || !out) return 0;
    const int n = max_words < TRACE_WORDS * TRACE_CTAS ? max_words : TRACE_WORDS * TRACE_CTAS;
    for (int i = 0; i < n; ++i) out[i] = static_cast<volatile unsigned int*>(atspeed::g_trace_host)[i];
    return n;
}
int atspeed_abi_version(void) { return ATSPEED_ABI_VERSION; }

int atspeed_session_workspace_bytes(const atspeed_model_desc* target, const atspeed_model_desc* draft,
                                    const atspeed_config* cfg, size_t* bytes) {
    ATS_TRY(check_cfg(target, draft, cfg));
    ATS_CHECK_ARG(bytes, "null bytes");
    atspeed_session tmp;
    tmp.cfg = *cfg;
    tmp.num_sms = cfg->num_sms > 0 ? cfg->num_sms : 148;
    session_dims(&tmp);
    Carver c{nullptr, 0};
    carve_session(c, &tmp, target, draft);
    *bytes = c.off + 4096;
    return ATS_OK;
}

int atspeed_session_create(const atspeed_model_desc* target, const atspeed_model_desc* draft, const atspeed_config* cfg,
                           const atspeed_trie_desc* trie, void* workspace, size_t workspace_bytes, atspeed_session** out) {
    ATS_TRY(check_cfg(target, draft, cfg));
    ATS_CHECK_ARG(trie && workspace && out, "null trie/workspace/out");
    ATS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    int dev = 0, sms = cfg->num_sms;
    if (sms <= 0) {
        ATS_CUDA(cudaGetDevice(&dev));
        ATS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    atspeed_session* s = new atspeed_session();
    s->cfg = *cfg;
    s->num_sms = sms;
    s->has_draft = draft != nullptr;
    session_dims(s);
    s->geom.V = target->vocab;
    s->trie.child_off = trie->child_off; s->trie.child_tok = trie->child_tok; s->trie.child_node = trie->child_node;
    s->trie.n_nodes = trie->n_nodes; s->trie.n_edges = trie->n_edges;
    Carver c{static_cast<uint8_t*>(workspace), 0};
    carve_session(c, s, target, draft);
    if (c.off > workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", c.off, workspace_bytes);
        delete s;
        return ATS_ERR_ARG;
    }
    int r = build_model(s->tgt, sms);
    if (r == ATS_OK && draft) r = build_model(s->dft, sms);
    if (r != ATS_OK) { delete s; return r; }
    if (cudaMemcpy(s->trees_dev, s->trees_host.data(), sizeof(TreeDev) * s->trees_host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("uploading the per-user tree table failed");
        delete s;
        return ATS_ERR_CUDA;
    }
    if (cudaMallocHost(&s->pinned, 64 * sizeof(int)) != cudaSuccess) {
        set_error("cudaMallocHost failed");
        delete s;
        return ATS_ERR_CUDA;
    }
    memset(s->pinned, 0, 64 * sizeof(int));
    s->cohort_pinned = nullptr;
    if (s->max_users > 1 &&
        cudaMallocHost(&s->cohort_pinned, sizeof(int) * (MAX_USERS * 4 + MAX_USERS * MAX_K * MAX_NEW + MAX_USERS + 4) +
                                              sizeof(float) * MAX_USERS * MAX_K) != cudaSuccess) {   // + the shared-prefix flag
        set_error("cudaMallocHost failed");
        cudaFreeHost(s->pinned);
        delete s;
        return ATS_ERR_CUDA;
    }
    s->P = 0;
    s->seed = cfg->seed; s->user_seq = 0; s->next_user_seq = 0; s->round = 0;
    {
        const int min_keep = cfg->K > 1 ? 2 : 1;     // transformers 4.41 _get_logits_warper: min_tokens_to_keep
        s->sample_B = cfg->top_k > min_keep ? cfg->top_k : min_keep;
        if (s->sample_B > MAX_BEAMS) s->sample_B = MAX_BEAMS;
    }
    s->launches = 0;
    s->prefix_len = 0;
    s->fused_epoch = 0;
    // opt-in (ATSPEED_FUSED_EPI=1): correct on every tile path (tests/test_gpu_fused_epilogue.py) but measured SLOWER than the
    // row-wise consumer kernels (124 vs 139 users/s, profiles/r02_fused_epilogue_ab.txt): with one TMEM accumulator at T > 256
    // the epilogue is exposed, and 128 epilogue threads per SM do math that the row-wise kernels spread over 2048
    { const char* e = getenv("ATSPEED_FUSED_EPI"); s->fused = e && atoi(e) == 1; }
    s->prof_on = false;
    s->prof_n = 0;
    for (double& b : s->prof_bytes) b = 0;
    for (double& b : s->prof_flops) b = 0;
    *out = s;
    return ATS_OK;
}

int atspeed_session_destroy(atspeed_session* s) {
    if (!s) return ATS_OK;
    if (s->pinned) cudaFreeHost(s->pinned);
    if (s->cohort_pinned) cudaFreeHost(s->cohort_pinned);
    for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
    delete s;
    return ATS_OK;
}

int atspeed_session_begin(atspeed_session* s, const int32_t* prompt_host, int32_t P, void* stream) {
    ATS_CHECK_ARG(s && prompt_host, "null session/prompt");
    ATS_CHECK_ARG(P >= 1 && P <= s->cfg.max_prompt, "prompt length %d outside [1,%d]", P, s->cfg.max_prompt);
    const int dl_max = s->cfg.max_new_tokens - 1;
    ATS_CHECK_ARG(P + dl_max * s->cfg.N <= s->T_max && P + s->cfg.K <= s->T_max,
                  "prompt length %d + %d tree tokens exceeds the %d-token forward limit", P, dl_max * s->cfg.N, s->T_max);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    s->P = P;
    ATS_CUDA(cudaMemcpyAsync(s->prompt_dev, prompt_host, sizeof(int) * P, cudaMemcpyHostToDevice, st));
    PROF(s, CAT_BEAM, 0, tree_begin(s->tree, s->batch, s->prompt_dev, P, st));
    s->launches += 1;
    s->user_seq = s->next_user_seq++;
    s->round = 0;
    s->pinned[H_FIRST] = 1;    // first
    s->pinned[H_MISS] = 0;    // missing ancestors pending for the draft
    s->pinned[H_LEVEL] = 0;   // result level
    return ATS_OK;
}

int atspeed_session_begin_device(atspeed_session* s, const int32_t* prompt_dev, int32_t P, void* stream) {
    ATS_CHECK_ARG(s && prompt_dev, "null session/prompt");
    ATS_CHECK_ARG(P >= 1 && P <= s->cfg.max_prompt, "prompt length %d outside [1,%d]", P, s->cfg.max_prompt);
    const int dl_max = s->cfg.max_new_tokens - 1;
    ATS_CHECK_ARG(P + dl_max * s->cfg.N <= s->T_max && P + s->cfg.K <= s->T_max,
                  "prompt length %d + %d tree tokens exceeds the %d-token forward limit", P, dl_max * s->cfg.N, s->T_max);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    s->P = P;
    ATS_CUDA(cudaMemcpyAsync(s->prompt_dev, prompt_dev, sizeof(int) * P, cudaMemcpyDeviceToDevice, st));
    PROF(s, CAT_BEAM, 0, tree_begin(s->tree, s->batch, s->prompt_dev, P, st));
    s->launches += 1;
    s->user_seq = s->next_user_seq++;
    s->round = 0;
    s->pinned[H_FIRST] = 1;
    s->pinned[H_MISS] = 0;
    s->pinned[H_LEVEL] = 0;
    return ATS_OK;
}

int atspeed_session_result_device(atspeed_session* s, int32_t* tokens_dev, float* scores_dev, void* stream) {
    ATS_CHECK_ARG(s && tokens_dev && scores_dev, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int level = s->pinned[H_LEVEL];
    // [K][MAX_NEW] generated tokens and [K] scores of the current beams, device to device, no synchronisation
    ATS_CUDA(cudaMemcpyAsync(tokens_dev, s->tree.gen + static_cast<size_t>(level) * MAX_BEAMS * MAX_NEW,
                             sizeof(int) * s->cfg.K * MAX_NEW, cudaMemcpyDeviceToDevice, st));
    ATS_CUDA(cudaMemcpyAsync(scores_dev, s->tree.score + static_cast<size_t>(level) * MAX_BEAMS, sizeof(float) * s->cfg.K,
                             cudaMemcpyDeviceToDevice, st));
    return ATS_OK;
}

int atspeed_session_profile(atspeed_session* s, int32_t enable) {
    ATS_CHECK_ARG(s, "null session");
    if (enable && s->prof_ev.empty()) {
        s->prof_ev.resize(2 * PROF_PAIRS);
        s->prof_cat.resize(PROF_PAIRS);
        for (cudaEvent_t& e : s->prof_ev) ATS_CUDA(cudaEventCreate(&e));
    }
    s->prof_on = enable != 0;
    s->prof_n = 0;
    for (double& b : s->prof_bytes) b = 0;
    for (double& b : s->prof_flops) b = 0;
    return ATS_OK;
}

int atspeed_session_profile_read(atspeed_session* s, double* ms6, int64_t* count6, double* bytes6, double* flops6,
                                 void* stream) {
    ATS_CHECK_ARG(s && ms6 && count6 && bytes6, "null argument");
    ATS_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    for (int c = 0; c < CAT_COUNT; ++c) {
        ms6[c] = 0; count6[c] = 0; bytes6[c] = s->prof_bytes[c];
        if (flops6) flops6[c] = s->prof_flops[c];
    }
    for (int i = 0; i < s->prof_n; ++i) {
        float ms = 0.f;
        ATS_CUDA(cudaEventElapsedTime(&ms, s->prof_ev[2 * i], s->prof_ev[2 * i + 1]));
        ms6[s->prof_cat[i]] += ms;
        count6[s->prof_cat[i]] += 1;
    }
    s->prof_n = 0;
    for (double& b : s->prof_bytes) b = 0;
    for (double& b : s->prof_flops) b = 0;
    return ATS_OK;
}

int atspeed_session_draft(atspeed_session* s, int32_t draft_len, void* stream) {
    ATS_CHECK_ARG(s && s->has_draft, "session has no draft model");
    ATS_CHECK_ARG(draft_len >= 1 && draft_len < MAX_LEVELS, "draft_len=%d", draft_len);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int j = 0; j < draft_len; ++j)
        ATS_TRY(search_step(s, s->dft, j, s->cfg.N, s->pinned[H_FIRST] != 0, j == 0 && s->pinned[H_MISS] != 0, st));
    return ATS_OK;
}

int atspeed_session_target(atspeed_session* s, int32_t draft_len, void* stream) {
    ATS_CHECK_ARG(s, "null session");
    ATS_CHECK_ARG(draft_len >= 1 && draft_len < MAX_LEVELS, "draft_len=%d", draft_len);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TreeGeom& g = s->geom;
    const int P = s->P;
    const bool first = s->pinned[H_FIRST] != 0;
    BatchPlan plan;
    memset(&plan, 0, sizeof(plan));
    plan.width = g.N;
    int T, R;
    if (first) {
        plan.with_prompt = 1; plan.l_from = 1; plan.l_to = draft_len; plan.rows_from = 1; plan.root_row = 1;
        T = P + draft_len * g.N; R = 1 + draft_len * g.N;
    } else {
        plan.l_from = 0; plan.l_to = draft_len; plan.rows_from = 0;
        T = g.K + draft_len * g.N; R = T;
    }
    const int S = g.tree_slot(P, draft_len, 0) + g.N;
    PROF(s, CAT_BEAM, 0, tree_build_batch(s->tree, s->batch, g, plan, s->prompt_dev, P, T, R, st));
    s->launches += 1;
    ATS_TRY(forward(s, s->tgt, batch_desc(s), T, S, s->batch.rows_idx, R, st));
    ATS_TRY(run_topk(s, s->tgt, R, s->cfg.do_sample ? s->sample_B : g.K, shared_cand(s), st));
    return ATS_OK;
}

int atspeed_session_verify(atspeed_session* s, int32_t draft_len, int32_t* n_matches_host, void* stream) {
    ATS_CHECK_ARG(s && n_matches_host, "null session/out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TreeGeom& g = s->geom;
    const bool first = s->pinned[H_FIRST] != 0;
    if (s->cfg.do_sample)
        PROF(s, CAT_BEAM, 0,
             tree_verify_relaxed(s->tree, g, s->trie, draft_len, first ? 1 : g.K, s->cand_tok, s->cand_edge, s->cand_logp,
                                 s->cand_cnt, s->P, sample_cfg(s), st));
    else
        PROF(s, CAT_BEAM, 0,
             tree_verify_strict(s->tree, g, s->trie, draft_len, first ? 1 : g.K, s->cand_tok, s->cand_edge, s->cand_logp,
                                s->cand_cnt, s->P, st));
    // kernel (c): move the survivors' ancestor rows into the accepted region, both caches, all layers
    const int max_rows = (draft_len + 1) * g.K;
    for (ModelRT* m : {&s->tgt, s->has_draft ? &s->dft : nullptr}) {
        if (!m) continue;
        PROF(s, CAT_GATHER, 0,
             kv_gather_rows(m->f32 ? static_cast<void*>(m->fkv) : static_cast<void*>(m->kv), m->kv_plane * m->elem_bytes,
                            m->d.n_layers * 2, m->HD * m->elem_bytes, s->tree.gather_src, s->tree.gather_dst,
                            s->tree.scal + SC_GATHER, max_rows, st));
        s->launches += 1;
    }
    s->launches += 1;
    ATS_CUDA(cudaMemcpyAsync(s->pinned, s->tree.scal, sizeof(int) * SC_COUNT, cudaMemcpyDeviceToHost, st));
    ATS_CUDA(cudaStreamSynchronize(st));
    *n_matches_host = s->pinned[SC_NMATCH];
    s->pinned[H_FIRST] = 0;
    s->pinned[H_MISS] = s->pinned[SC_MISS] > 0 ? 1 : 0;
    s->pinned[H_LEVEL] = 0;
    s->round += 1;
    return ATS_OK;
}

int atspeed_session_step(atspeed_session* s, int32_t model, int32_t width, void* stream) {
    ATS_CHECK_ARG(s, "null session");
    ATS_CHECK_ARG(model == 0 || (model == 1 && s->has_draft), "model=%d", model);
    ATS_CHECK_ARG(width >= 1 && width <= MAX_BEAMS, "width=%d", width);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int level = s->pinned[H_LEVEL];
    ATS_CHECK_ARG(level + 1 < MAX_LEVELS, "too many consecutive steps (%d)", level);
    ATS_TRY(search_step(s, model == 0 ? s->tgt : s->dft, level, width, s->pinned[H_FIRST] != 0 && level == 0,
                        model == 1 && level == 0 && s->pinned[H_MISS] != 0, st));
    s->pinned[H_LEVEL] = level + 1;
    return ATS_OK;
}

int atspeed_session_sort_result(atspeed_session* s, void* stream) {
    ATS_CHECK_ARG(s, "null session");
    if (!s->cfg.do_sample) return ATS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PROF(s, CAT_BEAM, 0, tree_sort_level(s->tree, s->pinned[H_LEVEL], st));
    s->launches += 1;
    return ATS_OK;
}

int atspeed_session_set_seed(atspeed_session* s, uint64_t seed, uint64_t user_seq) {
    ATS_CHECK_ARG(s, "null session");
    s->seed = seed;
    s->next_user_seq = user_seq;
    return ATS_OK;
}

uint64_t atspeed_noise_stream(uint64_t user_seq, uint32_t round, uint32_t level, uint32_t site) {
    return noise_stream(user_seq, round, level, site);
}

uint32_t atspeed_noise_host_u32(uint64_t seed, uint64_t stream, uint32_t index) { return philox_u32(seed, stream, index); }

int atspeed_noise_fill(uint64_t seed, uint64_t stream, int32_t kind, int32_t n, void* out_dev, void* stream_handle) {
    return noise_fill(seed, stream, kind, n, out_dev, static_cast<cudaStream_t>(stream_handle));
}

int atspeed_session_sample_width(atspeed_session* s) { return s ? s->sample_B : 0; }

int atspeed_session_result(atspeed_session* s, int32_t* tokens_host, float* scores_host, int32_t* count, void* stream) {
    ATS_CHECK_ARG(s && tokens_host && scores_host && count, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int level = s->pinned[H_LEVEL];
    int cnt[MAX_LEVELS];
    int gen[MAX_BEAMS * MAX_NEW];
    float sc[MAX_BEAMS];
    ATS_CUDA(cudaMemcpyAsync(cnt, s->tree.cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    ATS_CUDA(cudaMemcpyAsync(gen, s->tree.gen + static_cast<size_t>(level) * MAX_BEAMS * MAX_NEW, sizeof(gen),
                             cudaMemcpyDeviceToHost, st));
    ATS_CUDA(cudaMemcpyAsync(sc, s->tree.score + static_cast<size_t>(level) * MAX_BEAMS, sizeof(sc), cudaMemcpyDeviceToHost, st));
    ATS_CUDA(cudaStreamSynchronize(st));
    const int n = cnt[level] < s->cfg.K ? cnt[level] : s->cfg.K;
    const int L = s->cfg.max_new_tokens;
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < L; ++k) tokens_host[i * L + k] = gen[i * MAX_NEW + k];
        scores_host[i] = sc[i];
    }
    *count = n;
    return ATS_OK;
}

static int bssd_loop(atspeed_session* s, int32_t gamma, atspeed_stats* stats, long long l0, int tf0, int df0, void* stream) {
    const int L = s->cfg.max_new_tokens;
    int done = 0, n_run = 0, total = 0;
    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (done < L) {
        int dl = gamma < L - done - 1 ? gamma : L - done - 1;                 // beamSD.py:504
        if (dl > MAX_LEVELS - 1) dl = MAX_LEVELS - 1;
        if (dl == 0) {                                                         // beamSD.py:505-509
            ATS_TRY(atspeed_session_step(s, 0, s->cfg.K, stream));
            break;
        }
        ATS_TRY(atspeed_session_draft(s, dl, stream));
        ATS_TRY(atspeed_session_target(s, dl, stream));
        int m = 0;
        ATS_TRY(atspeed_session_verify(s, dl, &m, stream));
        done += m + 1;
        if (n_run < 8) acc[n_run] = m;
        ++n_run;
        total += m;
    }
    ATS_TRY(atspeed_session_sort_result(s, stream));
    if (stats) {
        stats->n_run = n_run;
        stats->total_accept_steps = total;
        for (int i = 0; i < 8; ++i) stats->accept_steps[i] = acc[i];
        stats->target_forwards = s->tgt.forwards - tf0;
        stats->draft_forwards = s->dft.forwards - df0;
        stats->kernel_launches = static_cast<int>(s->launches - l0);
    }
    return ATS_OK;
}

int atspeed_bssd(atspeed_session* s, const int32_t* prompt_host, int32_t P, int32_t gamma, int32_t* tokens_host,
                 float* scores_host, int32_t* count, atspeed_stats* stats, void* stream) {
    ATS_CHECK_ARG(s && s->has_draft, "session has no draft model");
    ATS_CHECK_ARG(gamma >= 1, "gamma=%d", gamma);
    const long long l0 = s->launches;
    const int tf0 = s->tgt.forwards, df0 = s->dft.forwards;
    ATS_TRY(atspeed_session_begin(s, prompt_host, P, stream));
    ATS_TRY(bssd_loop(s, gamma, stats, l0, tf0, df0, stream));
    return atspeed_session_result(s, tokens_host, scores_host, count, stream);
}

int atspeed_bssd_device(atspeed_session* s, const int32_t* prompt_dev, int32_t P, int32_t gamma, int32_t* tokens_dev,
                        float* scores_dev, atspeed_stats* stats, void* stream) {
    ATS_CHECK_ARG(s && s->has_draft, "session has no draft model");
    ATS_CHECK_ARG(gamma >= 1, "gamma=%d", gamma);
    const long long l0 = s->launches;
    const int tf0 = s->tgt.forwards, df0 = s->dft.forwards;
    ATS_TRY(atspeed_session_begin_device(s, prompt_dev, P, stream));
    ATS_TRY(bssd_loop(s, gamma, stats, l0, tf0, df0, stream));
    return atspeed_session_result_device(s, tokens_dev, scores_dev, stream);
}

int atspeed_target_generate(atspeed_session* s, const int32_t* prompt_host, int32_t P, int32_t* tokens_host,
                            float* scores_host, int32_t* count, atspeed_stats* stats, void* stream) {
    ATS_CHECK_ARG(s, "null session");
    ATS_CHECK_ARG(s->cfg.max_new_tokens < MAX_LEVELS, "max_new_tokens=%d too large for target_generate", s->cfg.max_new_tokens);
    const long long l0 = s->launches;
    const int tf0 = s->tgt.forwards;
    ATS_TRY(atspeed_session_begin(s, prompt_host, P, stream));
    for (int i = 0; i < s->cfg.max_new_tokens; ++i) ATS_TRY(atspeed_session_step(s, 0, s->cfg.K, stream));
    ATS_TRY(atspeed_session_sort_result(s, stream));
    ATS_TRY(atspeed_session_result(s, tokens_host, scores_host, count, stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->target_forwards = s->tgt.forwards - tf0;
        stats->kernel_launches = static_cast<int>(s->launches - l0);
    }
    return ATS_OK;
}

int atspeed_session_read(atspeed_session* s, int32_t field, void* host_dst, size_t bytes, void* stream) {
    ATS_CHECK_ARG(s && host_dst, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const void* src = nullptr;
    size_t avail = 0;
    const TreeDev& t = s->tree;
    switch (field) {
        case ATSPEED_F_LEVEL_CNT: src = t.cnt; avail = sizeof(int) * MAX_LEVELS; break;
        case ATSPEED_F_LEVEL_TOK: src = t.tok; avail = sizeof(int) * MAX_LEVELS * MAX_BEAMS; break;
        case ATSPEED_F_LEVEL_PARENT: src = t.parent; avail = sizeof(int) * MAX_LEVELS * MAX_BEAMS; break;
        case ATSPEED_F_LEVEL_SCORE: src = t.score; avail = sizeof(float) * MAX_LEVELS * MAX_BEAMS; break;
        case ATSPEED_F_LEVEL_NODE: src = t.node; avail = sizeof(int) * MAX_LEVELS * MAX_BEAMS; break;
        case ATSPEED_F_SCALARS: src = t.scal; avail = sizeof(int) * SC_COUNT; break;
        case ATSPEED_F_PICK_PARENT: src = t.tr_pick_parent; avail = sizeof(int) * MAX_LEVELS * MAX_K; break;
        case ATSPEED_F_PICK_TOK: src = t.tr_pick_tok; avail = sizeof(int) * MAX_LEVELS * MAX_K; break;
        case ATSPEED_F_PICK_SCORE: src = t.tr_pick_score; avail = sizeof(float) * MAX_LEVELS * MAX_K; break;
        case ATSPEED_F_HIT_POS: src = t.tr_hit_pos; avail = sizeof(int) * MAX_LEVELS * MAX_K; break;
        case ATSPEED_F_NPICK: src = t.tr_npick; avail = sizeof(int) * MAX_LEVELS; break;
        case ATSPEED_F_LOGITS_TARGET: src = s->tgt.logits; avail = sizeof(float) * s->R_max * s->tgt.ldl; break;
        case ATSPEED_F_LOGITS_DRAFT:
            ATS_CHECK_ARG(s->has_draft, "no draft model");
            src = s->dft.logits; avail = sizeof(float) * s->R_max * s->dft.ldl; break;
        case ATSPEED_F_ROW_NODE: src = s->batch.row_node; avail = sizeof(int) * s->R_max; break;
        case ATSPEED_F_TR_ACC: src = t.tr_acc; avail = sizeof(int) * MAX_LEVELS * MAX_BEAMS; break;
        case ATSPEED_F_LSE_Q: src = t.lse_q; avail = sizeof(float) * MAX_LEVELS; break;
        default: set_error("unknown field %d", field); return ATS_ERR_ARG;
    }
    ATS_CHECK_ARG(bytes <= avail, "field %d holds %zu bytes, %zu requested", field, avail, bytes);
    ATS_CUDA(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, st));
    ATS_CUDA(cudaStreamSynchronize(st));
    return ATS_OK;
}

int atspeed_session_info(atspeed_session* s, int64_t* info8) {
    ATS_CHECK_ARG(s && info8, "null argument");
    info8[0] = s->tgt.ldl; info8[1] = s->R_max; info8[2] = s->T_max; info8[3] = s->S_max; info8[4] = s->geom.A_cap;
    info8[5] = s->launches; info8[6] = s->tgt.last_rows; info8[7] = s->has_draft ? s->dft.last_rows : 0;
    return ATS_OK;
}

int atspeed_session_forward_raw(atspeed_session* s, int32_t model, const int32_t* tok, const int32_t* pos,
                                const int32_t* slot, const int32_t* prefix_len, const uint32_t* vis, int32_t vis_base,
                                int32_t T, int32_t S, const int32_t* rows_idx, int32_t R, void* stream) {
    ATS_CHECK_ARG(s && tok && pos && slot && prefix_len && vis && rows_idx, "null argument");
    ATS_CHECK_ARG(model == 0 || (model == 1 && s->has_draft), "model=%d", model);
    BatchDesc b;
    memset(&b, 0, sizeof(b));
    b.tok = tok; b.pos = pos; b.slot = slot; b.prefix_len = prefix_len; b.vis = vis; b.vis_base = vis_base; b.n_valid = nullptr;
    return forward(s, model == 0 ? s->tgt : s->dft, b, T, S, rows_idx, R, static_cast<cudaStream_t>(stream));
}

int atspeed_mask_logsoftmax_topk(const void* logits, int32_t logits_bf16, int32_t rows, int32_t V, int64_t ld,
                                 const int32_t* row_node, const int32_t* n_rows_dev, const atspeed_trie_desc* trie,
                                 int32_t B, int32_t* cand_tok, int32_t* cand_edge, float* cand_logp, int32_t* cand_cnt,
                                 float* lse, void* stream) {
    ATS_CHECK_ARG(logits && row_node && trie && cand_tok && cand_edge && cand_logp && cand_cnt, "null argument");
    TrieCSR t;
    t.child_off = trie->child_off; t.child_tok = trie->child_tok; t.child_node = trie->child_node;
    t.n_nodes = trie->n_nodes; t.n_edges = trie->n_edges;
    return mask_logsoftmax_topk(logits, logits_bf16, rows, V, ld, row_node, n_rows_dev, t, B, cand_tok, cand_edge,
                                cand_logp, cand_cnt, lse, static_cast<cudaStream_t>(stream));
}

int atspeed_kv_gather(const void* src_base, void* dst_base, int64_t src_plane_stride, int64_t dst_plane_stride,
                      int32_t n_planes, int32_t row_bytes, const int32_t* src_rows, const int32_t* dst_rows,
                      const int32_t* n_rows_dev, int32_t rows, void* stream) {
    ATS_CHECK_ARG(src_base && dst_base && src_rows && dst_rows, "null argument");
    return kv_gather_rows_oop(src_base, dst_base, src_plane_stride, dst_plane_stride, n_planes, row_bytes, src_rows,
                              dst_rows, n_rows_dev, rows, static_cast<cudaStream_t>(stream));
}

static int standalone_gemm(const void* x, int32_t T, int32_t K, const void* const* ws, const int* rs, GemmWeights* g,
                           GemmPlan* pl, XMap* xm) {
    memset(g, 0, sizeof(*g));
    g->K = K;
    int col = 0;
    for (int i = 0; i < 3 && ws[i]; ++i) {
        g->rows[i] = rs[i]; g->colbase[i] = col;
        ATS_TRY(make_tmap_bf16_kmajor(&g->tmap[i], ws[i], rs[i], K, 128));
        ATS_TRY(make_tmap_bf16_kmajor(&g->tmap256[i], ws[i], rs[i], K, 256));
        col += rs[i];
        g->n = i + 1;
    }
    int dev = 0, sms = 148;
    ATS_CUDA(cudaGetDevice(&dev));
    ATS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ATS_TRY(gemm_make_plan(*g, T, sms, true, pl));
    return x ? gemm_make_xmap(xm, x, T, K, pl->cluster) : ATS_OK;
}

int atspeed_gemm_plan(int32_t T, int32_t K, int32_t rows0, int32_t rows1, int32_t rows2, int32_t num_sms, int32_t allow_cut,
                      int32_t* info16, int32_t* slices_of_col) {
    ATS_CHECK_ARG(info16 && rows0 > 0, "null info / rows0=%d", rows0);
    GemmWeights g = shape_only(K, {rows0});
    if (rows1 > 0) { g.rows[1] = rows1; g.colbase[1] = rows0; g.n = 2; }
    if (rows2 > 0) { g.rows[2] = rows2; g.colbase[2] = rows0 + rows1; g.n = 3; }
    GemmPlan pl;
    ATS_TRY(gemm_make_plan(g, T, num_sms, allow_cut != 0, &pl));
    const int v[16] = {pl.BM, pl.KB, pl.total_tiles, pl.U, pl.grid, pl.max_slices, pl.stages, pl.tmem_cols, pl.n_bufs, pl.T_pad,
                       pl.tiles[0], pl.tiles[1], pl.tiles[2], pl.two_cta, pl.n_mma, pl.N_mma};   // [13]: 0 single CTA, 1 CTA pair
    for (int i = 0; i < 16; ++i) info16[i] = v[i];
    if (slices_of_col) {
        const SplitMap sm = gemm_split_map(g, pl);
        const int cols = rows0 + (rows1 > 0 ? rows1 : 0) + (rows2 > 0 ? rows2 : 0);
        for (int c = 0; c < cols; ++c) slices_of_col[c] = sm.slices(c);
    }
    return ATS_OK;
}

int atspeed_gemm_scratch_bytes(int32_t T, int32_t K, int32_t rows0, int32_t rows1, int32_t rows2, size_t* bytes) {
    ATS_CHECK_ARG(bytes, "null bytes");
    GemmWeights g = shape_only(K, {rows0});
    if (rows1 > 0) { g.rows[1] = rows1; g.colbase[1] = rows0; g.n = 2; }
    if (rows2 > 0) { g.rows[2] = rows2; g.colbase[2] = rows0 + rows1; g.n = 3; }
    int dev = 0, sms = 148;
    ATS_CUDA(cudaGetDevice(&dev));
    ATS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GemmPlan pl;
    ATS_TRY(gemm_make_plan(g, T, sms, true, &pl));
    *bytes = static_cast<size_t>(pl.max_slices) * T * (rows0 + rows1 + rows2) * sizeof(float);
    return ATS_OK;
}

int atspeed_gemm_bf16(const void* x, int32_t T, int32_t K, const void* w0, int32_t rows0, const void* w1, int32_t rows1,
                      const void* w2, int32_t rows2, float* scratch, float* out, int32_t ldo, void* stream) {
    ATS_CHECK_ARG(x && w0 && scratch, "null argument");
    const void* ws[3] = {w0, w1, w2};
    const int rs[3] = {rows0, rows1, rows2};
    GemmWeights g;
    GemmPlan pl;
    XMap xm;
    ATS_TRY(standalone_gemm(x, T, K, ws, rs, &g, &pl, &xm));
    const int cols = g.colbase[g.n - 1] + g.rows[g.n - 1];
    ATS_CHECK_ARG(!out || ldo >= cols, "ldo=%d < total columns %d", ldo, cols);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OMap om;
    ATS_TRY(gemm_make_omap(&om, g, scratch, cols, static_cast<long long>(T) * cols, T, pl.max_slices));
    ATS_TRY(gemm_wx(g, xm, pl, om, st));
    if (out) ATS_TRY(reduce_slices(scratch, gemm_split_map(g, pl), static_cast<long long>(T) * cols, cols, T, cols, out, ldo, st));
    return ATS_OK;
}

int atspeed_tree_attention(const void* q, const void* kcache, const void* vcache, const int32_t* prefix_len,
                           const uint32_t* vis, int32_t vis_base, int32_t T, int32_t S, int32_t n_heads, int32_t head_dim,
                           void* out, void* stream) {
    ATS_CHECK_ARG(q && kcache && vcache && prefix_len && vis && out, "null argument");
    BatchDesc b;
    memset(&b, 0, sizeof(b));
    b.prefix_len = prefix_len; b.vis = vis; b.vis_base = vis_base;
    return tree_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kcache),
                          static_cast<const __nv_bfloat16*>(vcache), b, T, S, n_heads, head_dim,
                          static_cast<__nv_bfloat16*>(out), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
