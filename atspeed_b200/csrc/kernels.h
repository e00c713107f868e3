// Internal C++ interface between the .cu files of libatspeed_b200 (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace atspeed {

// ---- gemm.cu ------------------------------------------------------------------------------------
constexpr int MAX_USERS = 16;     // users whose trees may share one forward (cohort.cu)
struct GemmWeights {
    int n;                 // 1..3 weight matrices sharing one input
    int K;                 // input features
    int rows[3];           // output features of each
    int colbase[3];        // first output column of each
    CUtensorMap tmap[3];   // [rows_i, K] bf16, K-major, box 64 x 128, SWIZZLE_128B
    CUtensorMap tmap256[3];   // same tensors, box 64 x 256: one TMA operation per 256-row tile (BM = 256)
    CUtensorMap tmap64[3];    // same tensors, box 64 x 64: gate / up halves of an interleaved SiLU tile (fused epilogue)
};
// Fused epilogues (gemm.cu): instead of fp32 partial-sum slices for a row-wise consumer kernel, the GEMM itself finishes
// the op that follows it in the LLaMA layer.  A tile whose k-range is cut across CTAs is completed INSIDE the kernel: the
// CTA that holds the tile's first k-block owns it; the other CTAs store their raw fp32 partial and raise a flag; the owner
// adds the partials in slice order (the order the consumer kernels used: same bits) and applies the epilogue.
enum : int { EPI_SLICES = 0, EPI_QKV_ROPE = 1, EPI_SILU_MUL = 2 };
struct FusedEpi {
    int kind;
    float* part;             // [workers x CTAs per worker x halves][T_pad][128] fp32 raw partial of a non-owner segment
    unsigned int* flags;     // [workers x CTAs per worker]: epoch of the launch whose partial is complete
    unsigned int epoch;      // unique per launch within the session (never 0)
    // EPI_QKV_ROPE: q|k|v projections -> RoPE(q, k) at pos -> q to qbuf [T, H*D], k/v to the cache rows slot[t]
    // (elementwise.cu qkv_rope_append restated per tile: one 128-row tile holds whole heads)
    const int* pos;
    const int* slot;
    const int* tok_user;     // cohort forward: user of each token (index into kv_off); nullptr otherwise
    const float* rope_cos;
    const float* rope_sin;
    int max_pos, head_dim, HD;
    __nv_bfloat16 *qbuf, *kcache, *vcache;
    long long kv_off[MAX_USERS];
    // EPI_SILU_MUL: gate|up projections -> m = bf16(bf16(silu(g)) * u)  (elementwise.cu silu_mul restated per tile: a tile
    // holds the gate rows AND the up rows of the same features)
    __nv_bfloat16* m;
    int mlp;
};
int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* base, long long rows, long long cols, int box_rows);
// One launch's work decomposition (host-computed, see gemm.cu): persistent CTAs each own a contiguous range of
// (tile, k-block) units; tiles cut across CTAs produce one fp32 partial-sum slice per CTA.
struct GemmPlan {
    int T, T_pad, KB;
    int BM;                 // 128 or 256 output features per tile
    int tiles[3], tilebase[3], total_tiles;
    int U;                  // units per CTA
    int grid;
    int max_slices;         // most slices any tile of this launch has
    int stages, tmem_cols, acc_stride, n_bufs, buf_stride;
    int two_cta, n_mma, N_mma;   // CTA-pair kernel (T > 256): see gemm_wx_tcgen05_2cta
    int cluster;                 // CTA-pair kernel: CTAs per cluster, 2 (one pair) or 4 (two pairs sharing the activations by multicast)
    int epi_kind;                // EPI_*: EPI_SLICES plans feed a consumer kernel, the others finish tiles in the kernel
};
// What a consumer of the partial sums needs to know: how many slices hold column `col`.  The host tabulates the count per tile
// (gemm_split_map), so the device side is a shift and a byte load; launches with more than SPLIT_TAB tiles fall back to the
// arithmetic (two integer divisions per query: they made the row-wise consumers issue-bound).
constexpr int SPLIT_TAB = 256;
struct SplitMap {
    int n;
    int colbase[3], tilebase[3];
    int BM, KB, U;
    int tps;                // tiles per unit row: 2 when the work unit is a super-tile (clusters of 4), else 1
    int tab_n, bm_shift;    // tab_n > 0: tab[tile] is valid for tile < tab_n and BM == 1 << bm_shift
    unsigned char tab[SPLIT_TAB];
    __host__ __device__ __forceinline__ int slices_of_tile(int tile) const {
        const int u0 = (tile / tps) * KB;
        return (u0 + KB - 1) / U - u0 / U + 1;
    }
    __host__ __device__ __forceinline__ int slices(int col) const {
        const int i = col >= colbase[2] ? 2 : (col >= colbase[1] ? 1 : 0);
        if (tab_n > 0) return tab[tilebase[i] + ((col - colbase[i]) >> bm_shift)];
        return slices_of_tile(tilebase[i] + (col - colbase[i]) / BM);
    }
};
struct XMap { CUtensorMap tm0, tm1; int T, K, box0, cluster; };   // activation operand [T, K] bf16 (two boxes: tokens < 256, >= 256)
int gemm_cluster_size(int num_sms);     // CTAs per cluster of the pair kernel on this device (2 or 4)
bool gemm_use_2cta(int T);
int gemm_make_plan(const GemmWeights& w, int T, int num_sms, bool allow_cut, GemmPlan* plan);
// plan of a fused-epilogue launch (kind = EPI_QKV_ROPE: w = {q, k, v}; EPI_SILU_MUL: w = {gate, up})
int gemm_make_plan_fused(const GemmWeights& w, int T, int num_sms, int kind, GemmPlan* plan);
size_t gemm_fused_part_elems(int T_max, int num_sms);     // floats of FusedEpi::part that cover every plan up to T_max
int gemm_wx_fused(const GemmWeights& w, const XMap& xm, const GemmPlan& plan, const FusedEpi& epi, cudaStream_t stream);
int gemm_max_slices(const GemmWeights& w, int T_max, int num_sms);
SplitMap gemm_split_map(const GemmWeights& w, const GemmPlan& plan);
int gemm_make_xmap(XMap* xm, const void* x, int T, int K, int cluster);   // cluster: GemmPlan::cluster of the launch that reads it
struct OMap { float* out; int ldo; long long slice_stride; int T; };   // fp32 partial-sum output: out[slice][t][column]
int gemm_make_omap(OMap* om, const GemmWeights& w, float* out, int ldo, long long slice_stride, int T, int max_slices);
int gemm_wx(const GemmWeights& w, const XMap& xm, const GemmPlan& plan, const OMap& om, cudaStream_t stream);
bool pdl_enabled();   // ATSPEED_PDL=0 disables programmatic dependent launch (debugging)

// ---- elementwise.cu -----------------------------------------------------------------------------
// Several users' trees in ONE forward ("cohort", cohort.cu): every user keeps its own KV cache, prompt length and
// attention extent; tokens of user i occupy batch positions [tok0[i], tok0[i] + T[i]).  n == 0: single-user forward.
struct CohortKV {
    int n;
    long long kv_off[MAX_USERS];   // element offset of the user's cache inside the model's KV allocation
    int S[MAX_USERS];              // KV slots the user's tokens may attend
    int vis_base[MAX_USERS];       // = the user's prompt length
    int tok0[MAX_USERS], T[MAX_USERS];
};

// Forward-batch descriptor, all device arrays of length >= T (built by beam.cu kernels).
struct BatchDesc {
    const int* tok;          // token ids
    const int* pos;          // RoPE positions
    const int* slot;         // KV slot each token writes
    const int* prefix_len;   // token attends KV slots [0, prefix_len)
    const uint32_t* vis;     // [T][VIS_WORDS] bit j: token attends slot vis_base + j
    int vis_base;            // first slot covered by the bit masks (= prompt length)
    const int* n_valid;      // device scalar: tokens >= *n_valid are padding (may be nullptr = all valid)
    const int* tok_user;     // cohort forward: user index (into ckv) of each token; nullptr otherwise
    CohortKV ckv;
};
int embed_rows(const __nv_bfloat16* table, const int* tok, int T, int hidden, int vocab, __nv_bfloat16* h,
               cudaStream_t st);
// x = rmsnorm(h) * g   (first layer / when no residual is pending)
int rmsnorm_rows(const __nv_bfloat16* h, const __nv_bfloat16* g, int T, int hidden, float eps, __nv_bfloat16* x,
                 const int* row_index, cudaStream_t st);
// h = bf16(h + bf16(sum_s part[s])) ; x = rmsnorm(h) * g  (g == nullptr: residual only)
int residual_rmsnorm(__nv_bfloat16* h, const float* part, const SplitMap& sm, long long split_stride, int ldp,
                     const __nv_bfloat16* g, int T, int hidden, float eps, __nv_bfloat16* x, cudaStream_t st);
// q,k,v = bf16(sum_s part[s]); RoPE(q,k) at pos; q -> qbuf [T, H*D]; k,v -> cache rows slot[t]
// rope_cos/rope_sin: [max_pos][head_dim/2] fp32 tables holding bf16-rounded values (computed on the host exactly
// as HF's LlamaRotaryEmbedding does, so the device never evaluates powf/cosf/sinf).
int qkv_rope_append(const float* part, const SplitMap& sm, long long split_stride, int ldp, const BatchDesc& b, int T,
                    int n_heads, int head_dim, const float* rope_cos, const float* rope_sin, int max_pos,
                    __nv_bfloat16* qbuf, __nv_bfloat16* kcache, __nv_bfloat16* vcache, cudaStream_t st);
// m = bf16(bf16(silu(g)) * u) with g,u = bf16(sum_s part[s]) at columns [0,mlp) and [mlp,2mlp)
int silu_mul(const float* part, const SplitMap& sm, long long split_stride, int ldp, int T, int mlp, __nv_bfloat16* m,
             cudaStream_t st);
// out[t][c] = sum of the slices of column c (stand-alone GEMM entry point / tests)
int reduce_slices(const float* part, const SplitMap& sm, long long split_stride, int ldp, int T, int cols, float* out,
                  int ldo, cudaStream_t st);

// ---- forward_f32.cu: fp32 exact-parity forward (weights handed over as fp32) -----------------------------
int f32_embed(const float* table, const int* tok, int T, int hidden, int vocab, float* h, cudaStream_t st);
int f32_rmsnorm(const float* h, const float* g, int T, int hidden, float eps, float* x, const int* row_index, cudaStream_t st);
int f32_gemm(const float* x, const float* w, int T, int N, int K, float* out, int ldo, bool accumulate, cudaStream_t st);
int f32_rope_append(const float* qkv, const BatchDesc& b, int T, int n_heads, int head_dim, const float* rope_cos,
                    const float* rope_sin, int max_pos, float* qbuf, float* kcache, float* vcache, cudaStream_t st);
int f32_tree_attention(const float* q, const float* kcache, const float* vcache, const BatchDesc& b, int T, int S, int n_heads,
                       int head_dim, float* out, cudaStream_t st);
int f32_silu_mul(const float* gu, int T, int mlp, float* m, cudaStream_t st);

// ---- attention.cu -------------------------------------------------------------------------------
int tree_attention(const __nv_bfloat16* q, const __nv_bfloat16* kcache, const __nv_bfloat16* vcache,
                   const BatchDesc& b, int T, int S, int n_heads, int head_dim, __nv_bfloat16* out, cudaStream_t st);

// ---- topk.cu: kernel (a) ------------------------------------------------------------------------
struct TrieCSR {
    const int* child_off;
    const int* child_tok;
    const int* child_node;
    int n_nodes, n_edges;
};
// logits: rows x ld (bf16 if logits_bf16 else fp32). row_node[r] < 0 or r >= *n_rows_dev: row skipped (count 0).
// Outputs per row r: cand_cnt[r] <= B; cand_tok/cand_edge/cand_logp[r*B + j] sorted by (logp desc, token asc);
// lse[r] = log sum exp over the full vocabulary.
int mask_logsoftmax_topk(const void* logits, int logits_bf16, int rows, int V, long long ld, const int* row_node,
                         const int* n_rows_dev, const TrieCSR& trie, int B, int* cand_tok, int* cand_edge,
                         float* cand_logp, int* cand_cnt, float* lse, cudaStream_t st);

// ---- kvgather.cu: kernel (c) --------------------------------------------------------------------
// For every plane p < n_planes and i < *n_rows_dev (<= max_rows): copy row_bytes from
// base + p*plane_stride + src[i]*row_bytes to base + p*plane_stride + dst[i]*row_bytes. Source and
// destination rows must not overlap as sets.
int kv_gather_rows(void* base, long long plane_stride, int n_planes, int row_bytes, const int* src, const int* dst,
                   const int* n_rows_dev, int max_rows, cudaStream_t st);
int kv_gather_rows_oop(const void* src_base, void* dst_base, long long src_plane_stride, long long dst_plane_stride,
                       int n_planes, int row_bytes, const int* src, const int* dst, const int* n_rows_dev,
                       int max_rows, cudaStream_t st);
// the same for several users in one launch: user i moves rows inside base + byte_off[i] using its own lists
struct GatherCohort {
    int n;
    long long byte_off[MAX_USERS];
    const int* src[MAX_USERS];
    const int* dst[MAX_USERS];
    const int* n_rows[MAX_USERS];
};
int kv_gather_rows_cohort(void* base, long long plane_stride, int n_planes, int row_bytes, const GatherCohort& gc,
                          int max_rows, cudaStream_t st);

}  // namespace atspeed
