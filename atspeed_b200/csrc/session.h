// Session state shared by engine.cu (single-user stages + C ABI) and cohort.cu (multi-user batched scheduler).
#pragma once
#include <vector>

#include "../../include/atspeed.h"
#include "beam.cuh"
#include "kernels.h"

namespace atspeed {

struct LayerRT {
    GemmWeights qkv, o, gu, down;
    const __nv_bfloat16 *ln1, *ln2;
};

struct ModelRT {
    atspeed_model_desc d;
    std::vector<LayerRT> layers;
    GemmWeights lm;
    int HD;
    // activations (device, carved from the workspace)
    __nv_bfloat16 *h, *x, *q, *a, *m, *xsel, *kv;
    float *part, *logits;
    // fp32 exact-parity mode (d.weights_f32): fp32 activations and KV cache instead of the bf16 ones above
    bool f32;
    float *fh, *fx, *fq, *fa, *fm, *fxsel, *fqkv, *fgu, *fkv;
    int elem_bytes;       // bytes per KV / activation element (2 or 4)
    long long kv_plane;   // elements per K (or V) plane of one layer
    int ldl;              // logits row stride
    int last_rows;
    int forwards;
};

// host-side flags kept next to the pinned scalar mirror: is the root still the prompt, does the draft owe KV for
// accepted tokens, which tree level holds the current beams
enum { H_FIRST = 32, H_MISS = 33, H_LEVEL = 34 };

struct Carver {
    uint8_t* base;
    size_t off;
    template <typename T> T* take(size_t n) {
        off = (off + 1023) & ~size_t(1023);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

}  // namespace atspeed

struct atspeed_session {
    // ---- cohort mode (cfg.max_users > 1): several users' searches share every forward (cohort.cu) ----
    int max_users;
    std::vector<atspeed::TreeDev> trees_host;    // per-user tree state (device pointers)
    atspeed::TreeDev* trees_dev;                 // the same array in HBM (kernels index it by user slot)
    int* prompts_dev;                            // [max_users][max_prompt]
    int* collect_dev;                            // [max_users][4] verify outcomes
    int *res_tok_dev, *res_cnt_dev;              // [max_users][K][MAX_NEW], [max_users]
    float* res_score_dev;                        // [max_users][K]
    long long kv_user_elems_tgt, kv_user_elems_dft;   // elements between two users' KV caches
    int* cohort_pinned;                          // pinned host staging of the cohort scheduler
    // shared prompt prefix (atspeed_session_set_shared_prefix): user slot `max_users` of trees / prompts / KV caches holds the
    // prefix; its K/V rows are copied into a user's caches when the user is admitted and the user's forwards skip those tokens
    int prefix_len;
    int* prefix_iota_dev;                        // 0, 1, ..., max_prompt - 1
    int* prefix_bad_dev;                         // != 0: a device-resident prompt did not start with the prefix
    // fused GEMM epilogues (gemm.cu FusedEpi): in-kernel partial-sum workspace, per-CTA flags, launch epoch
    float* fused_part;
    unsigned int* fused_flags;
    unsigned int fused_epoch;
    bool fused;                                  // false: ATSPEED_FUSED_EPI=0, the row-wise consumer kernels run instead

    atspeed_config cfg;
    atspeed::TreeGeom geom;
    atspeed::TrieCSR trie;
    atspeed::ModelRT tgt, dft;
    bool has_draft;
    atspeed::TreeDev tree;
    atspeed::BatchDev batch;
    int *cand_tok, *cand_edge, *cand_cnt;
    float *cand_logp, *lse;
    int* prompt_dev;
    int T_max, R_max, S_max;
    int P;
    int num_sms;
    int* pinned;          // pinned host scratch
    long long launches;
    // sampling mode (AtSpeed-R): noise key, user sequence number of the current search, round within it
    unsigned long long seed, user_seq, next_user_seq;
    int round;
    int sample_B;
    // optional per-launch CUDA-event timing (bench.py roofline / share-of-step); off by default
    bool prof_on;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_cat;
    int prof_n;
    double prof_bytes[6];
    double prof_flops[6];
};

namespace atspeed {

enum { CAT_GEMM = 0, CAT_ATTN = 1, CAT_ELEM = 2, CAT_TOPK = 3, CAT_BEAM = 4, CAT_GATHER = 5, CAT_COUNT = 6 };
static constexpr int PROF_PAIRS = 16384;

static inline void prof_begin(atspeed_session* s, int cat, double bytes, cudaStream_t st, double flops = 0.0) {
    if (!s->prof_on || s->prof_n >= PROF_PAIRS) return;
    s->prof_cat[s->prof_n] = cat;
    s->prof_bytes[cat] += bytes;
    s->prof_flops[cat] += flops;
    cudaEventRecord(s->prof_ev[2 * s->prof_n], st);
}
static inline void prof_end(atspeed_session* s, cudaStream_t st) {
    if (!s->prof_on || s->prof_n >= PROF_PAIRS) return;
    cudaEventRecord(s->prof_ev[2 * s->prof_n + 1], st);
    s->prof_n++;
}
#define PROF(s, cat, bytes, call)              \
    do {                                       \
        prof_begin(s, cat, bytes, st);         \
        ATS_TRY(call);                         \
        prof_end(s, st);                       \
    } while (0)
#define PROF_GEMM(s, w, T, call)                                          \
    do {                                                                  \
        prof_begin(s, CAT_GEMM, gemm_bytes(w, T), st, gemm_flops(w, T));  \
        ATS_TRY(call);                                                    \
        prof_end(s, st);                                                  \
    } while (0)


struct CandOut { int* tok; int* edge; float* logp; int* cnt; };
CandOut shared_cand(atspeed_session* s);
SampleCfg sample_cfg(const atspeed_session* s);
// one forward of `m` over the batch in `b`: T tokens, logits for the R rows listed in rows_idx; S = KV slots scanned by
// attention (ignored in cohort forwards, where b.ckv carries it per user)
int forward(atspeed_session* s, ModelRT& m, const BatchDesc& b, int T, int S, const int* rows_idx, int R, cudaStream_t st);
// cohort.cu: packing of the target forwards of one scheduler step (see the definition)
struct PackItem { int T, R, waited; };
int plan_packs(const PackItem* items, int n, int T_max, int R_max, int max_users, bool defer, bool no_more_work, int* order,
               int* pack_of, unsigned char* run_now);
int run_topk(atspeed_session* s, ModelRT& m, int R, int B, const CandOut& o, cudaStream_t st);

}  // namespace atspeed
