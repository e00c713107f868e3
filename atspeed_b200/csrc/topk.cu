// Kernel (a): fused trie-constrained mask + log-softmax + top-B over the vocabulary.
//
// Replaces, per beam row (reference file:line relative to /root/reference/code):
//   F.log_softmax(logits)                                   beamSD.py:58, :285          (G1)
//   logits_processor(...) = PrefixConstrainedLogitsProcessor beamSD.py:62,64,288,291     (G2)
//     -> a Python loop with one `.tolist()` sync + one index_put_ per row
//   the per-row part of `(scores + beam_scores).view(-1).topk(B)` beamSD.py:69-78,297-325 (G3)
// Order of operations is the reference's: log-softmax over the FULL vocabulary, the constraint is
// applied afterwards (scores are not renormalised over allowed tokens).  The global top-B over
// rows x V is contained in the union of the per-row top-B, so this kernel emits [rows, B] candidates
// and beam.cu merges them after adding the per-row parent score: one HBM pass over the logits, no
// [rows, V] temporary, no host sync.
//
// HBM-bound: algorithmic bytes = rows * V * sizeof(logit).  One CTA per row, 128-bit streaming loads
// (ld.global.nc.L1::no_allocate) for the online max/sum-exp pass; the row's CSR children (<= 256 for
// the item tries) are then gathered from L2 and ranked in shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace atspeed {

static constexpr int TOPK_THREADS = 256;
static constexpr int TOPK_SMEM_CAND = 2048;

__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

struct MaxSum {
    float m, s;   // running max, sum of exp(x - m)
};
__device__ __forceinline__ void ms_add_chunk(MaxSum& a, const float* x, int n) {
    float cm = x[0];
    for (int i = 1; i < n; ++i) cm = fmaxf(cm, x[i]);
    if (cm == -INFINITY) return;
    if (cm > a.m) { a.s *= exp2f((a.m - cm) * 1.4426950408889634f); a.m = cm; }   // exp2f(-inf) = 0 for the first chunk
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += exp2f((x[i] - a.m) * 1.4426950408889634f);
    a.s += acc;
}
__device__ __forceinline__ MaxSum ms_merge(MaxSum a, MaxSum b) {
    if (b.m == -INFINITY) return a;
    if (a.m == -INFINITY) return b;
    MaxSum r;
    r.m = fmaxf(a.m, b.m);
    r.s = a.s * exp2f((a.m - r.m) * 1.4426950408889634f) + b.s * exp2f((b.m - r.m) * 1.4426950408889634f);
    return r;
}

// UNR = 128-bit loads in flight per thread in the streaming pass.  With fewer rows than ~2 CTAs per SM the kernel is bound by
// memory-level parallelism, not bandwidth: 8 loads in flight read a 121-row launch in 10.5 us instead of 16.6 us (fp32; bf16
// 9.8 vs 12.7 us); with thousands of rows occupancy provides the parallelism and UNR = 2 is faster (bf16, 8192 rows: 4.9 vs
// 4.0 TB/s).  Measured in profiles/r02_kernels_a_c_gbs.txt; the launcher picks by row count, ATSPEED_TOPK_UNROLL forces one.
template <typename T, int UNR>
__global__ void __launch_bounds__(TOPK_THREADS)
mask_logsoftmax_topk_kernel(const T* __restrict__ logits, int V, long long ld, const int* __restrict__ row_node,
                            const int* __restrict__ n_rows_dev, const int* __restrict__ child_off,
                            const int* __restrict__ child_tok, int n_nodes, int B, int* __restrict__ cand_tok,
                            int* __restrict__ cand_edge, float* __restrict__ cand_logp, int* __restrict__ cand_cnt,
                            float* __restrict__ lse_out) {
    __shared__ MaxSum red[TOPK_THREADS / 32];
    __shared__ float s_lse;
    __shared__ unsigned long long s_key[TOPK_SMEM_CAND];
    __shared__ int s_count;
    __shared__ unsigned long long s_best[TOPK_THREADS / 32];

    const int r = blockIdx.x;
    const int node = row_node[r];
    const bool active = (n_rows_dev == nullptr || r < *n_rows_dev) && node >= 0 && node < n_nodes;
    if (!active) {
        if (threadIdx.x == 0) { cand_cnt[r] = 0; if (lse_out) lse_out[r] = 0.f; }
        return;
    }
    const T* row = logits + static_cast<long long>(r) * ld;

    // ---- pass 1: online max / sum-exp over the full vocabulary, 128-bit streaming loads ----
    constexpr int EPV = 16 / sizeof(T);
    MaxSum acc{-INFINITY, 0.f};
    const uintptr_t addr = reinterpret_cast<uintptr_t>(row);
    int head = static_cast<int>(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > V) head = V;
    const int nvec = (V - head) / EPV;
    const int tail0 = head + nvec * EPV;
    if (threadIdx.x < head) { float x = to_f32<T>(row[threadIdx.x]); ms_add_chunk(acc, &x, 1); }
    if (threadIdx.x < V - tail0) { float x = to_f32<T>(row[tail0 + threadIdx.x]); ms_add_chunk(acc, &x, 1); }
    const uint4* vrow = reinterpret_cast<const uint4*>(row + head);
    int i = threadIdx.x;
    if (UNR > 2) {
        for (; i + (UNR - 1) * TOPK_THREADS < nvec; i += UNR * TOPK_THREADS) {
            uint4 u[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) u[j] = ld_stream_v4(vrow + i + j * TOPK_THREADS);
            float x[UNR * EPV];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const T* e = reinterpret_cast<const T*>(&u[j]);
#pragma unroll
                for (int k = 0; k < EPV; ++k) x[j * EPV + k] = to_f32<T>(e[k]);
            }
            ms_add_chunk(acc, x, UNR * EPV);
        }
    }
    for (; i + TOPK_THREADS < nvec; i += 2 * TOPK_THREADS) {     // two loads in flight per thread
        const uint4 u0 = ld_stream_v4(vrow + i), u1 = ld_stream_v4(vrow + i + TOPK_THREADS);
        float x[2 * EPV];
        const T* e0 = reinterpret_cast<const T*>(&u0);
        const T* e1 = reinterpret_cast<const T*>(&u1);
#pragma unroll
        for (int j = 0; j < EPV; ++j) { x[j] = to_f32<T>(e0[j]); x[EPV + j] = to_f32<T>(e1[j]); }
        ms_add_chunk(acc, x, 2 * EPV);
    }
    for (; i < nvec; i += TOPK_THREADS) {
        const uint4 u0 = ld_stream_v4(vrow + i);
        float x[EPV];
        const T* e0 = reinterpret_cast<const T*>(&u0);
#pragma unroll
        for (int j = 0; j < EPV; ++j) x[j] = to_f32<T>(e0[j]);
        ms_add_chunk(acc, x, EPV);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxSum other;
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
        acc = ms_merge(acc, other);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = acc;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        MaxSum t = red[0];
        for (int w = 1; w < TOPK_THREADS / 32; ++w) t = ms_merge(t, red[w]);
        s_lse = t.m + logf(t.s);
        if (lse_out) lse_out[r] = s_lse;
    }
    __syncthreads();
    const float lse = s_lse;

    // ---- pass 2: the row's allowed children (CSR), ranked by (logp desc, token asc) ----
    const int lo = child_off[node], hi = child_off[node + 1];
    const int C = hi - lo;
    int* o_tok = cand_tok + static_cast<long long>(r) * B;
    int* o_edge = cand_edge + static_cast<long long>(r) * B;
    float* o_lp = cand_logp + static_cast<long long>(r) * B;
    if (C <= TOPK_SMEM_CAND) {
        for (int e = threadIdx.x; e < C; e += TOPK_THREADS) {
            const int tok = child_tok[lo + e];
            const float lp = (tok >= 0 && tok < V) ? to_f32<T>(row[tok]) - lse : -INFINITY;
            // non-finite candidates are dropped: key 0 sorts below every real key
            const bool fin = lp > -INFINITY && lp < INFINITY;
            s_key[e] = fin ? rank_key(lp, static_cast<uint32_t>(e)) : 0ull;   // children ascend by token: e orders like tok
            if (fin) atomicAdd(&s_count, 1);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < C; e += TOPK_THREADS) {
            const unsigned long long k = s_key[e];
            if (k == 0ull) continue;
            int rank = 0;
            for (int j = 0; j < C; ++j) rank += s_key[j] > k;
            if (rank < B) {
                const int tok = child_tok[lo + e];
                o_tok[rank] = tok;
                o_edge[rank] = lo + e;
                o_lp[rank] = to_f32<T>(row[tok]) - lse;
            }
        }
        if (threadIdx.x == 0) cand_cnt[r] = s_count < B ? s_count : B;
    } else {
        // wide nodes (e.g. no constraint at all): B rounds of block-wide arg-max below the previous pick
        unsigned long long prev = ~0ull;
        int n_out = 0;
        for (int round = 0; round < B; ++round) {
            unsigned long long best = 0ull;
            for (int e = threadIdx.x; e < C; e += TOPK_THREADS) {
                const int tok = child_tok[lo + e];
                const float lp = (tok >= 0 && tok < V) ? to_f32<T>(row[tok]) - lse : -INFINITY;
                if (lp > -INFINITY && lp < INFINITY) {
                    const unsigned long long k = rank_key(lp, static_cast<uint32_t>(e));
                    if (k < prev && k > best) best = k;
                }
            }
            best = warp_max_u64(best);
            if (lane == 0) s_best[warp] = best;
            __syncthreads();
            best = s_best[0];
            for (int w = 1; w < TOPK_THREADS / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
            __syncthreads();
            if (best == 0ull) break;
            if (threadIdx.x == 0) {
                const int e = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffull));
                const int tok = child_tok[lo + e];
                o_tok[round] = tok;
                o_edge[round] = lo + e;
                o_lp[round] = to_f32<T>(row[tok]) - lse;
            }
            prev = best;
            ++n_out;
        }
        if (threadIdx.x == 0) cand_cnt[r] = n_out;
    }
}

int mask_logsoftmax_topk(const void* logits, int logits_bf16, int rows, int V, long long ld, const int* row_node,
                         const int* n_rows_dev, const TrieCSR& trie, int B, int* cand_tok, int* cand_edge,
                         float* cand_logp, int* cand_cnt, float* lse, cudaStream_t st) {
    ATS_CHECK_ARG(rows >= 1 && V >= 1 && B >= 1 && B <= MAX_BEAMS, "topk: rows=%d V=%d B=%d", rows, V, B);
    ATS_CHECK_ARG(ld >= V, "topk: row stride %lld < V %d", ld, V);
    static const int unr_forced = []() { const char* e = getenv("ATSPEED_TOPK_UNROLL"); return e ? atoi(e) : 0; }();
    const int unr_env = unr_forced == 8 || unr_forced == 2 ? unr_forced : (rows <= 320 ? 8 : 2);
#define ATS_TOPK(TT, UU)                                                                                                  \
    mask_logsoftmax_topk_kernel<TT, UU><<<rows, TOPK_THREADS, 0, st>>>(static_cast<const TT*>(logits), V, ld, row_node,    \
                                                                       n_rows_dev, trie.child_off, trie.child_tok,        \
                                                                       trie.n_nodes, B, cand_tok, cand_edge, cand_logp,   \
                                                                       cand_cnt, lse)
    if (logits_bf16) {
        if (unr_env == 8) ATS_TOPK(__nv_bfloat16, 8); else ATS_TOPK(__nv_bfloat16, 2);
    } else {
        if (unr_env == 8) ATS_TOPK(float, 8); else ATS_TOPK(float, 2);
    }
#undef ATS_TOPK
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed
