// Tree attention for the draft/verify forward (reference: the 4-D additive mask built at
// code/beamSD.py:87-91,203-211,396-400 and consumed by transformers' LlamaAttention).
//
// The reference lays all beams of all levels along the sequence axis and hands the model a dense
// [T, S] fp32 mask.  Here a token's visibility is (a) a causal prefix length over the prompt slots and
// (b) a 512-bit mask over the accepted/tree slots that follow the prompt -- a few words per token,
// produced on the device by the beam kernels (beam.cu), never a dense mask.
//
// One CTA = 64 queries of one head and one user (4 warps x 16 rows).  The prompt keys -- seen by every query of the block --
// are streamed 64 at a time through a double-buffered cp.async ring, fragments come from ldmatrix (transposing for V),
// flash-style online softmax in fp32, QK^T and PV on the warp-level tensor-core path (mma.sync m16n8k16 bf16; P is split
// into hi+lo bf16 parts so the PV product is accurate to ~2^-16, the fp32-softmax contract of oracle/llama_ref.py).  The
// accepted / tree slots are NOT run through dense tiles: a beam sees <= ~8 of them, so each query folds its visible slots
// in one by one (sparse phase, see the kernel).  Round 1 ran 64-key tiles over the tree region as well: 31 us per layer at
// T = 300..512, latency-bound on > 95 % masked MMAs (profiles/r02_att_bench.txt).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace atspeed {

static constexpr int ATT_BQ = 64;   // queries per CTA
static constexpr int ATT_BK = 64;   // keys per tile
static constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t n = valid ? 16u : 0u;     // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}

// Two phases per CTA (64 queries of one head, one user):
//   dense  : keys [0, max prefix_len) -- the prompt, which every query of the block sees (causally for prompt tokens) --
//            in 64-key tiles on the warp-level tensor-core path, flash-style online softmax;
//   sparse : the accepted / tree slots behind the prompt.  A beam sees only its own ancestor chain there (<= ~8 of the
//            ~170 slots), so a dense pass over those tiles would be > 95 % masked work: instead every query walks the set
//            bits of its 512-bit visibility mask and folds each visible key into its running (max, sum, output) state with
//            plain fp32 dot products -- two threads per query, each half of the head dimension.
// The dense phase hands its per-row softmax state to the sparse phase through shared memory (the K/V ring is free by then).
template <int D>
__global__ void __launch_bounds__(ATT_THREADS)
tree_attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                      const __nv_bfloat16* __restrict__ vcache, const int* __restrict__ prefix_len,
                      const uint32_t* __restrict__ vis, int vis_base, int T, int S, int n_heads, float scale,
                      __nv_bfloat16* __restrict__ out, CohortKV ckv) {
    constexpr int BQ = ATT_BQ;
    constexpr int LDS = D + 8;                     // padded row (bf16 elements): conflict-free ldmatrix, rows 16-byte aligned
    constexpr int LDO = D + 4;                     // padded fp32 row of the state hand-off
    constexpr int DH = D / 2;                      // dimensions per thread in the sparse phase
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* sKV = sQ + BQ * LDS;        // [2 buffers][K | V][ATT_BK][LDS]; afterwards: fp32 [BQ][LDO] output state
    __shared__ uint32_t sVis[BQ * VIS_WORDS];
    __shared__ int sPl[BQ];
    __shared__ int sMaxPl;
    __shared__ float sM[BQ], sL[BQ];

    pdl_launch_dependents();
    pdl_wait();
    const int head = blockIdx.y;
    int q0 = blockIdx.x * BQ;
    if (ckv.n > 0) {
        // cohort forward: a CTA's 64 queries belong to ONE user (own KV cache, prompt length and extent); blockIdx.x
        // enumerates the users' 64-query blocks in order
        int b = blockIdx.x, u = 0;
        for (; u < ckv.n; ++u) {
            const int nb = (ckv.T[u] + BQ - 1) / BQ;
            if (b < nb) break;
            b -= nb;
        }
        if (u >= ckv.n) return;
        q0 = ckv.tok0[u] + b * BQ;
        T = ckv.tok0[u] + ckv.T[u];            // rows >= T are padding of this user's last block
        S = ckv.S[u];
        vis_base = ckv.vis_base[u];
        kcache += ckv.kv_off[u];
        vcache += ckv.kv_off[u];
    }
    const int HD = n_heads * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;

    if (threadIdx.x == 0) sMaxPl = 0;
    __syncthreads();
    // stage the Q tile with cp.async (group 0), prefix lengths and visibility words directly
    for (int i = threadIdx.x; i < BQ * (D / 8); i += ATT_THREADS) {
        const int r = i / (D / 8), c = (i % (D / 8)) * 8;
        const bool ok = q0 + r < T;
        cp_async16(&sQ[r * LDS + c], q + static_cast<long long>(ok ? q0 + r : 0) * HD + head * D + c, ok);
    }
    cp_async_commit();
    for (int i = threadIdx.x; i < BQ; i += ATT_THREADS) {
        const int pl = (q0 + i < T) ? min(prefix_len[q0 + i], S) : 0;
        sPl[i] = pl;
        atomicMax(&sMaxPl, pl);
    }
    for (int i = threadIdx.x; i < BQ * VIS_WORDS; i += ATT_THREADS) {
        const int r = i / VIS_WORDS, w = i % VIS_WORDS;
        sVis[i] = (q0 + r < T) ? vis[static_cast<long long>(q0 + r) * VIS_WORDS + w] : 0u;
    }
    __syncthreads();
    const int n_tiles = (sMaxPl + ATT_BK - 1) / ATT_BK;      // dense phase: the keys some query sees through its prefix

    auto load_tile = [&](int buf, int key0) {
        __nv_bfloat16* sK = sKV + static_cast<size_t>(buf) * 2 * ATT_BK * LDS;
        __nv_bfloat16* sV = sK + ATT_BK * LDS;
        for (int i = threadIdx.x; i < ATT_BK * (D / 8); i += ATT_THREADS) {
            const int r = i / (D / 8), c = (i % (D / 8)) * 8;
            const bool ok = key0 + r < S;
            const long long off = static_cast<long long>(ok ? key0 + r : 0) * HD + head * D + c;
            cp_async16(&sK[r * LDS + c], kcache + off, ok);
            cp_async16(&sV[r * LDS + c], vcache + off, ok);
        }
    };

    const int r0 = warp * 16 + g, r1 = r0 + 8;         // the two query rows this thread owns in the dense phase
    const int pl0 = sPl[r0], pl1 = sPl[r1];
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    if (n_tiles > 0) load_tile(0, 0);
    cp_async_commit();
    // ldmatrix lane addressing (see the fragment layouts of mma.m16n8k16):
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;        // A operand (Q), 16x16 tiles
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;        // B operand (K), 2 n-blocks x 16 k
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_col = (lane >> 4) * 8;        // B operand (V, transposed)

    for (int it = 0; it < n_tiles; ++it) {
        const int key0 = it * ATT_BK;
        const int buf = it & 1;
        if (it + 1 < n_tiles) load_tile(buf ^ 1, key0 + ATT_BK);      // overlaps this tile's math
        cp_async_commit();
        cp_async_wait<1>();                                            // this tile (and Q) have landed
        __syncthreads();
        const __nv_bfloat16* sK = sKV + static_cast<size_t>(buf) * 2 * ATT_BK * LDS;
        const __nv_bfloat16* sV = sK + ATT_BK * LDS;

        // ---- S = Q K^T for this warp's 16 rows x 64 keys ----
        float s[ATT_BK / 8][4];
#pragma unroll
        for (int nb = 0; nb < ATT_BK / 8; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            uint32_t a[4];
            ldmatrix_x4(a, &sQ[(warp * 16 + a_row) * LDS + kk * 16 + a_col]);
#pragma unroll
            for (int nb = 0; nb < ATT_BK / 8; nb += 2) {
                uint32_t b[4];
                ldmatrix_x4(b, &sK[(nb * 8 + b_row) * LDS + kk * 16 + b_col]);
                mma_bf16_16816(s[nb], a[0], a[1], a[2], a[3], b[0], b[1]);
                mma_bf16_16816(s[nb + 1], a[0], a[1], a[2], a[3], b[2], b[3]);
            }
        }
        // ---- causal-prefix mask + online softmax ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < ATT_BK / 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = key0 + nb * 8 + 2 * t4 + e;
                s[nb][e] = key < pl0 ? s[nb][e] * scale : -INFINITY;
                s[nb][2 + e] = key < pl1 ? s[nb][2 + e] * scale : -INFINITY;
                mx0 = fmaxf(mx0, s[nb][e]);
                mx1 = fmaxf(mx1, s[nb][2 + e]);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float ref0 = mn0 == -INFINITY ? 0.f : mn0, ref1 = mn1 == -INFINITY ? 0.f : mn1;
        const float corr0 = m0 == -INFINITY ? 0.f : expf(m0 - ref0), corr1 = m1 == -INFINITY ? 0.f : expf(m1 - ref1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < ATT_BK / 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                s[nb][e] = expf(s[nb][e] - ref0);          // exp(-inf) = 0 for masked keys
                s[nb][2 + e] = expf(s[nb][2 + e] - ref1);
                rs0 += s[nb][e];
                rs1 += s[nb][2 + e];
            }
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * corr0 + rs0;
        l1 = l1 * corr1 + rs1;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }

        // ---- O += P V, P = hi + lo (two bf16 parts: the product is accurate to ~2^-16, the oracle's fp32-softmax contract) ----
#pragma unroll
        for (int kk = 0; kk < ATT_BK / 16; ++kk) {
            // S accumulator layout of n-blocks 2kk, 2kk+1 == A fragment layout of a 16x16 tile
            float ph[8], plo[8];
            const float src[8] = {s[2 * kk][0], s[2 * kk][1], s[2 * kk][2], s[2 * kk][3],
                                  s[2 * kk + 1][0], s[2 * kk + 1][1], s[2 * kk + 1][2], s[2 * kk + 1][3]};
#pragma unroll
            for (int i = 0; i < 8; ++i) { ph[i] = bf16_round(src[i]); plo[i] = src[i] - ph[i]; }
            const uint32_t ah0 = pack_bf16(ph[0], ph[1]), ah1 = pack_bf16(ph[2], ph[3]);
            const uint32_t ah2 = pack_bf16(ph[4], ph[5]), ah3 = pack_bf16(ph[6], ph[7]);
            const uint32_t al0 = pack_bf16(plo[0], plo[1]), al1 = pack_bf16(plo[2], plo[3]);
            const uint32_t al2 = pack_bf16(plo[4], plo[5]), al3 = pack_bf16(plo[6], plo[7]);
#pragma unroll
            for (int nb = 0; nb < D / 8; nb += 2) {
                // B[k = key][n = d] from V stored [key][d]: transposing ldmatrix, two d-blocks per instruction
                uint32_t b[4];
                ldmatrix_x4_trans(b, &sV[(kk * 16 + v_row) * LDS + nb * 8 + v_col]);
                mma_bf16_16816(o[nb], ah0, ah1, ah2, ah3, b[0], b[1]);
                mma_bf16_16816(o[nb], al0, al1, al2, al3, b[0], b[1]);
                mma_bf16_16816(o[nb + 1], ah0, ah1, ah2, ah3, b[2], b[3]);
                mma_bf16_16816(o[nb + 1], al0, al1, al2, al3, b[2], b[3]);
            }
        }
        __syncthreads();   // this buffer is refilled by the next iteration's prefetch
    }
    cp_async_wait<0>();
    __syncthreads();       // Q has landed for everybody (n_tiles == 0) and the K/V ring is free: it now carries the row states
    float* sO = reinterpret_cast<float*>(sKV);
#pragma unroll
    for (int nb = 0; nb < D / 8; ++nb) {
        const int c = nb * 8 + 2 * t4;
        sO[r0 * LDO + c] = o[nb][0]; sO[r0 * LDO + c + 1] = o[nb][1];
        sO[r1 * LDO + c] = o[nb][2]; sO[r1 * LDO + c + 1] = o[nb][3];
    }
    if (t4 == 0) { sM[r0] = m0; sL[r0] = l0; sM[r1] = m1; sL[r1] = l1; }
    __syncthreads();

    // ---- sparse phase: two threads per query row, each DH dimensions ----
    const int row = threadIdx.x >> 1, part = threadIdx.x & 1;
    const int tq = q0 + row;
    if (tq >= T) return;
    const unsigned pair_mask = 3u << (lane & ~1);
    const int pl = sPl[row];
    float m = sM[row], l = sL[row];
    float acc[DH], qf[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) {
        acc[i] = sO[row * LDO + part * DH + i];
        qf[i] = __bfloat162float(sQ[row * LDS + part * DH + i]);
    }
    const long long col0 = static_cast<long long>(head) * D + part * DH;
    for (int w = 0; w < VIS_WORDS; ++w) {
        uint32_t bits = sVis[row * VIS_WORDS + w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int key = vis_base + w * 32 + b;
            if (key >= S || key < pl) continue;               // out of this forward's extent / already seen through the prefix
            const uint4* kp = reinterpret_cast<const uint4*>(kcache + static_cast<long long>(key) * HD + col0);
            const uint4* vp = reinterpret_cast<const uint4*>(vcache + static_cast<long long>(key) * HD + col0);
            uint4 kr[DH / 8], vr[DH / 8];
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) kr[i] = __ldg(kp + i);
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) vr[i] = __ldg(vp + i);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) {
                const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kr[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 kf = __bfloat1622float2(k2[j]);
                    dot += qf[i * 8 + 2 * j] * kf.x + qf[i * 8 + 2 * j + 1] * kf.y;
                }
            }
            dot += __shfl_xor_sync(pair_mask, dot, 1);
            const float sc = dot * scale;
            const float mn = fmaxf(m, sc);
            const float corr = m == -INFINITY ? 0.f : expf(m - mn);
            const float pw = expf(sc - mn);
            l = l * corr + pw;
            m = mn;
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) {
                const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vr[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 vf = __bfloat1622float2(v2[j]);
                    acc[i * 8 + 2 * j] = acc[i * 8 + 2 * j] * corr + pw * vf.x;
                    acc[i * 8 + 2 * j + 1] = acc[i * 8 + 2 * j + 1] * corr + pw * vf.y;
                }
            }
        }
    }
    // ---- normalise and store (bf16) ----
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    __nv_bfloat16* dst = out + static_cast<long long>(tq) * HD + col0;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
        uint4 pk;
        pk.x = pack_bf16(acc[i * 8 + 0] * inv, acc[i * 8 + 1] * inv);
        pk.y = pack_bf16(acc[i * 8 + 2] * inv, acc[i * 8 + 3] * inv);
        pk.z = pack_bf16(acc[i * 8 + 4] * inv, acc[i * 8 + 5] * inv);
        pk.w = pack_bf16(acc[i * 8 + 6] * inv, acc[i * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i * 8) = pk;
    }
}

int tree_attention(const __nv_bfloat16* q, const __nv_bfloat16* kcache, const __nv_bfloat16* vcache,
                   const BatchDesc& b, int T, int S, int n_heads, int head_dim, __nv_bfloat16* out, cudaStream_t st) {
    ATS_CHECK_ARG(T >= 1 && S >= 1, "attention: T=%d S=%d", T, S);
    int q_blocks = (T + ATT_BQ - 1) / ATT_BQ;
    if (b.ckv.n > 0) {
        q_blocks = 0;
        for (int u = 0; u < b.ckv.n; ++u) q_blocks += (b.ckv.T[u] + ATT_BQ - 1) / ATT_BQ;
    }
    dim3 grid(q_blocks, n_heads);
    const float scale = 1.0f / sqrtf(static_cast<float>(head_dim));
    const size_t smem = static_cast<size_t>(ATT_BQ + 4 * ATT_BK) * (head_dim + 8) * sizeof(__nv_bfloat16);
#define ATS_ATT(DD)                                                                                                       \
    do {                                                                                                                  \
        static std::once_flag once;                                                                                       \
        static cudaError_t err = cudaSuccess;                                                                             \
        std::call_once(once, []() {                                                                                       \
            err = cudaFuncSetAttribute(tree_attention_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
        });                                                                                                               \
        ATS_CUDA(err);                                                                                                    \
        ATS_CUDA(launch_pdl(tree_attention_kernel<DD>, grid, dim3(ATT_THREADS), smem, st, q, kcache, vcache,              \
                            b.prefix_len, b.vis, b.vis_base, T, S, n_heads, scale, out, b.ckv));                        \
    } while (0)
    switch (head_dim) {
        case 16: ATS_ATT(16); break;
        case 32: ATS_ATT(32); break;
        case 64: ATS_ATT(64); break;
        case 128: ATS_ATT(128); break;
        default:
            set_error("attention: head_dim=%d not in {16,32,64,128}", head_dim);
            return ATS_ERR_ARG;
    }
#undef ATS_ATT
    return ATS_OK;
}

}  // namespace atspeed
