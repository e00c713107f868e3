// Device-resident beam-tree state shared by beam.cu (kernels) and engine.cu (orchestration).
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "noise.cuh"

namespace atspeed {

// Per-user search state, all in HBM.  Levels: 0 = the round's root beams, 1.. = tokens proposed at each
// step.  A level entry is a beam (token, parent index in the previous level, cumulative score, trie
// node, KV slot, generated suffix, visibility mask over the accepted + tree KV slots).
struct TreeDev {
    int* cnt;          // [MAX_LEVELS]
    int* tok;          // [MAX_LEVELS][MAX_BEAMS]
    int* parent;       // [MAX_LEVELS][MAX_BEAMS]
    int* node;         // [MAX_LEVELS][MAX_BEAMS]
    int* slot;         // [MAX_LEVELS][MAX_BEAMS]
    float* score;      // [MAX_LEVELS][MAX_BEAMS]
    int* gen;          // [MAX_LEVELS][MAX_BEAMS][MAX_NEW]
    uint32_t* vis;     // [MAX_LEVELS][MAX_BEAMS][VIS_WORDS]
    // scalars: [0] P, [1] gen_len0 (tokens generated before this round), [2] first (roots == prompt),
    //          [3] acc_n (accepted KV slots in use), [4] n_matches of the last verify, [5] miss_n,
    //          [6] gather_n, [7] result_level
    int* scal;
    // verify trace (read back by tests): per verified level the K target picks and the hit positions
    int* tr_pick_parent;   // [MAX_LEVELS][MAX_K]  parent position in the draft's level list (or root index)
    int* tr_pick_tok;      // [MAX_LEVELS][MAX_K]
    float* tr_pick_score;  // [MAX_LEVELS][MAX_K]
    int* tr_hit_pos;       // [MAX_LEVELS][MAX_K]  draft position of each pick (-1 = not proposed)
    int* tr_npick;         // [MAX_LEVELS]
    // KV compaction lists produced by verify, consumed by kernel (c)
    int* gather_src;       // [MAX_LEVELS * MAX_K]
    int* gather_dst;
    // draft tokens whose KV is missing after a fully accepted round
    int* miss_tok;         // [MAX_K]
    int* miss_pos;
    int* miss_slot;
    uint32_t* miss_vis;    // [MAX_K][VIS_WORDS]
    // ---- AtSpeed-R (sampling) only ----
    // the draft's warped per-row candidate lists of every step (its q lives on them: beamSD.py:70, step_probs) and the
    // log-normaliser of its flat softmax; level l = the step that expanded level l into level l + 1
    int* dcand_tok;        // [MAX_LEVELS][MAX_BEAMS * MAX_BEAMS]: per level, kernel (a) output with row stride B
    float* dcand_logp;     // same layout
    int* dcand_edge;       // same layout
    int* dcand_cnt;        // [MAX_LEVELS][MAX_BEAMS]
    float* lse_q;          // [MAX_LEVELS]
    int* tr_acc;           // [MAX_LEVELS][MAX_BEAMS] acceptance flag of every draft pick (trace)
};
// sampling parameters of one launch (by value)
struct SampleCfg {
    int B;                 // warped candidates per row = max(top_k, min_tokens_to_keep)
    float inv_temp;        // 1 / temperature
    unsigned long long seed;
    unsigned long long stream_base;   // noise_stream(user_seq, round, 0, 0)
};
enum { SC_P = 0, SC_GEN0 = 1, SC_FIRST = 2, SC_ACC = 3, SC_NMATCH = 4, SC_MISS = 5, SC_GATHER = 6, SC_RESULT = 7,
       SC_FALLBACK = 8 /* rounds whose residual distribution was empty (reference undefined) */,
       SC_COUNT = 16 };

// Forward batch (one per model invocation) + logits-row selection, device arrays.
struct BatchDev {
    int* tok; int* pos; int* slot; int* prefix_len; uint32_t* vis;   // [T_max] ([T_max][VIS_WORDS])
    int* tok_user;    // [T_max] cohort forwards: which user of the cohort each token belongs to
    int* rows_idx;    // [R_max] batch index of each logits row
    int* row_node;    // [R_max] trie node of each logits row (-1 = padding row)
};

// Static (host-known) geometry of one search configuration.
struct TreeGeom {
    int K, N;          // target / draft beam widths
    int A_cap;         // accepted-region slots after the prompt (MAX_NEW * K rounded)
    int V;
    __host__ __device__ int lvl_off(int l) const { return l == 0 ? 0 : K + (l - 1) * N; }      // tree-slot offset of a level
    __host__ __device__ int tree_slot(int P, int l, int i) const { return P + A_cap + lvl_off(l) + i; }
};

// prompt batch: tok = prompt, pos = slot = i, prefix = i + 1
int tree_begin(const TreeDev& t, const BatchDev& b, const int* prompt, int P, cudaStream_t st);
// what a forward batch contains
struct BatchPlan {
    int with_prompt;   // 1: batch starts with the P prompt tokens (first round)
    int with_missing;  // 1: then the draft's missing ancestors (cap K)
    int l_from, l_to;  // then levels l_from..l_to (caps: level 0 -> K, others -> width)
    int width;         // cap of levels >= 1
    int rows_from;     // first level whose tokens produce logits rows; with_prompt adds the root row (P-1)
    int root_row;      // 1: emit the prompt's last position as row 0 (first round)
    int prompt_skip;   // with_prompt: the first prompt_skip prompt tokens are NOT in the batch -- their K/V already sit in the
                       // user's cache (shared prompt prefix, atspeed_session_set_shared_prefix); 0 otherwise
};
int tree_build_batch(const TreeDev& t, const BatchDev& b, const TreeGeom& g, const BatchPlan& plan, const int* prompt,
                     int P, int T_cap, int R_cap, cudaStream_t st);
// level l+1 = top-`width` of (logp + parent score) over the rows of level l
int tree_select(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int level, int row0, int B,
                const int* cand_tok, const int* cand_edge, const float* cand_logp, const int* cand_cnt, int width,
                int P, cudaStream_t st);
// kernel (b), strict mode
int tree_verify_strict(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int draft_len, int root_rows,
                       const int* cand_tok, const int* cand_edge, const float* cand_logp, const int* cand_cnt, int P,
                       cudaStream_t st);

// ---- cohort: several users' searches advanced by the same launches (cohort.cu) --------------------------------------
// What one launch does for one user; passed by value inside `Cohort` (kernel parameter, no H2D copy per launch).
struct UserCtx {
    int tree;              // index of the user's TreeDev in the session's device array
    int P;                 // prompt length
    int tok0, T;           // the user's token range in the forward batch
    int row0, R;           // the user's logits-row range
    int level, width;      // select: expand `level` into level + 1 keeping `width` beams
    int draft_len, root_rows;   // verify
    int mode;              // post-forward action: 1 = select (one_step_beam_search), 2 = verify, 0 = none
    int is_draft;          // sampling: the step belongs to the draft model (candidates are recorded, SITE_DRAFT noise)
    BatchPlan plan;
    unsigned long long stream_base;   // noise stream of (user, round)
};
struct Cohort {
    int n;
    UserCtx u[MAX_USERS];
};
int cohort_begin(const Cohort& c, const TreeDev* trees, cudaStream_t st);
int cohort_build_batch(const Cohort& c, const TreeDev* trees, const BatchDev& b, const TreeGeom& g, const int* prompts,
                       int prompt_stride, cudaStream_t st);
int cohort_check_prefix(const Cohort& c, const int* prompts, int prompt_stride, int prefix_row, int n, int* bad, cudaStream_t st);
// select / verify for every user of the cohort whose mode asks for it; candidates are read from cand_* rows
// [row0, row0 + R) of each user with row stride B
int cohort_select(const Cohort& c, const TreeDev* trees, const TreeGeom& g, const TrieCSR& trie, int B, const int* cand_tok,
                  const int* cand_edge, const float* cand_logp, const int* cand_cnt, bool sampling, const SampleCfg& sc,
                  cudaStream_t st);
int cohort_verify(const Cohort& c, const TreeDev* trees, const TreeGeom& g, const TrieCSR& trie, int B, const int* cand_tok,
                  const int* cand_edge, const float* cand_logp, const int* cand_cnt, bool sampling, const SampleCfg& sc,
                  cudaStream_t st);
// pack what the host needs after a verify: out[i] = {n_matches, miss_n, beams, fallbacks} of user i
int cohort_collect(const Cohort& c, const TreeDev* trees, int* out4, cudaStream_t st);
// final beams of every user (sorted by score first when `sort`): tokens [n][K][MAX_NEW], scores [n][K], counts [n]
int cohort_results(const Cohort& c, const TreeDev* trees, int K, bool sort, int* tokens, float* scores, int* counts,
                   cudaStream_t st);

// ---- AtSpeed-R ----
// sampling form of tree_select: N samples without replacement from softmax(logp / T + parent score) over the rows'
// warped candidates (beamSD.py:65-74); candidates are read from cand_* with row stride B = sc.B
int tree_select_sample(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int level, int row0, const int* cand_tok,
                       const int* cand_edge, const float* cand_logp, const int* cand_cnt, int width, int P,
                       const SampleCfg& sc, unsigned site, cudaStream_t st);
// kernel (b), relaxed mode (sequence-level speculative sampling, beamSD.py:293-321,332-369)
int tree_verify_relaxed(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int draft_len, int root_rows,
                        const int* cand_tok, const int* cand_edge, const float* cand_logp, const int* cand_cnt, int P,
                        const SampleCfg& sc, cudaStream_t st);
// sort the beams of `level` by score, descending (beamSD.py:529-531)
int tree_sort_level(const TreeDev& t, int level, cudaStream_t st);
// out[i] = noise(seed, stream, i): kind 0 = raw uint32 bits, 1 = uniform (0,1), 2 = Exp(1)
int noise_fill(unsigned long long seed, unsigned long long stream, int kind, int n, void* out, cudaStream_t st);

}  // namespace atspeed
