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
// accepted / tree slots are NOT run through dense tiles: a beam sees <= ~11 of them, so each query folds its visible slots
// in one by one (sparse phase, see the kernel).  Round 1 ran 64-key tiles over the tree region as well: 31 us per layer at
// T = 300..512, latency-bound on > 95 % masked MMAs (profiles/r02_att_bench.txt).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace atspeed {

// -DATT_TIMING (diagnostic builds only, tools/att_timing.py): thread 0 of CTA (0,0) leaves clock64 stamps at the phase boundaries
#ifdef ATT_TIMING
__device__ long long g_att_stamps[16];
#define ATT_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_att_stamps[i] = clock64(); } while (0)
#else
#define ATT_STAMP(i) do { } while (0)
#endif

static constexpr int ATT_BQ = 64;   // queries per CTA
static constexpr int ATT_BK = 64;   // keys per tile
static constexpr int ATT_THREADS = 128;
static constexpr int SPARSE_CAP = 16;   // visible tree / accepted slots folded per round of the sparse phase (a beam sees <= ~11)

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
// the 128-byte line at gptr -> L2: no destination, nothing to wait for.  (cp.async.bulk.prefetch.L2 was tried for the same
// purpose: hundreds of 256-byte requests per CTA queue in the TMA unit, +9 us per launch.)
__device__ __forceinline__ void l2_prefetch_line(const void* gptr) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(gptr) : "memory");
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
//   sparse : the accepted / tree slots behind the prompt.  A beam sees only its own ancestor chain there (<= ~11 of the
//            ~170 slots), so a dense pass over those tiles would be > 95 % masked work: instead every row's visible slots are
//            compacted into a short key list (before the dense phase, so their K/V lines can be prefetched into L2) and the
//            lists are folded into the rows' running (max, sum, output) states with plain fp32 dot products, 4 lanes per
//            row and 8 rows per warp at a time, K/V rows loaded two keys ahead.
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
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* sKV = sQ + BQ * LDS;        // [2 buffers][K | V][ATT_BK][LDS]; afterwards: fp32 [BQ][LDO] output state
    __shared__ uint32_t sVis[BQ * VIS_WORDS];
    __shared__ int sPl[BQ];
    __shared__ int sMaxPl;
    __shared__ float sM[BQ], sL[BQ];
    __shared__ unsigned short sKeys[BQ * SPARSE_CAP];
    __shared__ int sCnt[BQ];

    ATT_STAMP(0);
    pdl_launch_dependents();
    pdl_wait();
    ATT_STAMP(1);
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

    if (threadIdx.x == 0) sMaxPl = 0;
    __syncthreads();
    // stage the Q tile with cp.async (group 0) and -- without waiting to learn how many prompt tiles the block needs -- the
    // first K/V tile (group 1): the prefix lengths and visibility words below are a second global round trip that the first
    // tile would otherwise queue behind (a CTA has 4 warps: every exposed round trip is ~1 us of its ~10 us life)
    for (int i = threadIdx.x; i < BQ * (D / 8); i += ATT_THREADS) {
        const int r = i / (D / 8), c = (i % (D / 8)) * 8;
        const bool ok = q0 + r < T;
        cp_async16(&sQ[r * LDS + c], q + static_cast<long long>(ok ? q0 + r : 0) * HD + head * D + c, ok);
    }
    cp_async_commit();
    load_tile(0, 0);
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
    ATT_STAMP(2);

    // Sparse-phase key lists, built BEFORE the dense phase so that the rows can be requested from HBM now (prefetch.global.L2,
    // fire and forget): lane l < 16 of warp w compacts the visibility words of row 16 w + l into up to SPARSE_CAP slot
    // numbers, ascending (the order the keys are folded in).  The cursor stays in the lane's registers for further rounds.
    const int lrow = warp * 16 + (lane & 15);
    int vw = 0, my_cnt = 0;
    uint32_t vbits = 0u;
    auto fill_keys = [&]() {                                         // lanes < 16 only
        const int spl = sPl[lrow];
        int n = 0;
        while (n < SPARSE_CAP) {
            while (vbits == 0u && ++vw < VIS_WORDS) vbits = sVis[lrow * VIS_WORDS + vw];
            if (vbits == 0u) break;
            const int b = __ffs(vbits) - 1;
            vbits &= vbits - 1;
            const int key = vis_base + vw * 32 + b;
            if (key >= S || key < spl) continue;                     // out of this forward's extent / already seen through the prefix
            sKeys[lrow * SPARSE_CAP + n++] = static_cast<unsigned short>(key);
        }
        sCnt[lrow] = n;
        my_cnt = n;
    };
    if (lane < 16) {
        vbits = sVis[lrow * VIS_WORDS];                              // rows >= T hold zeros: empty list
        fill_keys();
        for (int i = 0; i < my_cnt; ++i) {
            const long long off = static_cast<long long>(sKeys[lrow * SPARSE_CAP + i]) * HD + head * D;
#pragma unroll
            for (int b = 0; b < D * 2; b += 128) {
                l2_prefetch_line(reinterpret_cast<const char*>(kcache + off) + b);
                l2_prefetch_line(reinterpret_cast<const char*>(vcache + off) + b);
            }
        }
    }
    __syncwarp();

    const int r0 = warp * 16 + g, r1 = r0 + 8;         // the two query rows this thread owns in the dense phase
    const int pl0 = sPl[r0], pl1 = sPl[r1];
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    // ldmatrix lane addressing (see the fragment layouts of mma.m16n8k16):
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;        // A operand (Q), 16x16 tiles
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;        // B operand (K), 2 n-blocks x 16 k
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_col = (lane >> 4) * 8;        // B operand (V, transposed)

    ATT_STAMP(3);
    for (int it = 0; it < n_tiles; ++it) {
        if (it < 4) ATT_STAMP(4 + it);
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
    ATT_STAMP(8);
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

    ATT_STAMP(9);
    // ---- sparse phase ----
    // A warp owns the 16 rows it owned in the dense phase and folds their key lists LPR lanes per row, RPP = 32 / LPR rows at a
    // time: lane j of a row holds the 16-byte chunks j, j + LPR, ... of the head dimension, so one load instruction of the
    // warp reads LPR * 16 contiguous bytes per row (the first version gave each thread half a row: every 16-byte load of a warp
    // touched 32 different lines and the phase was bound by L1 wavefronts, 2.2 us per key).  K and V rows are loaded two keys
    // ahead (three register sets) -- they sit in L2 by now (prefetched above), ~700 cycles away.
    constexpr int LPR = D >= 32 ? 4 : 2;
    constexpr int RPP = 32 / LPR;
    constexpr int NCH = D / (8 * LPR);                               // chunks (8 dimensions) per lane
    const int sub = lane % LPR, rsel = lane / LPR;
    const unsigned row_mask = ((1u << LPR) - 1u) << (lane - sub);
    struct KVRegs { uint4 k[NCH], v[NCH]; };
    for (;;) {
        // rounds of up to SPARSE_CAP keys per row (one round unless a row sees more): is this the last one?
        const bool last = !__any_sync(0xffffffffu, lane < 16 && my_cnt == SPARSE_CAP);
#pragma unroll 1
        for (int pass = 0; pass < 16 / RPP; ++pass) {
            const int row = warp * 16 + pass * RPP + rsel;
            const int tq = q0 + row;
            const int n = sCnt[row];
            const unsigned short* keys = sKeys + row * SPARSE_CAP;
            float m = sM[row], l = sL[row];
            float acc[NCH][8];
            uint4 qp[NCH];                                          // this lane's query dimensions, packed bf16
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int d0 = 8 * (sub + LPR * c);
                const float4 a0 = *reinterpret_cast<const float4*>(sO + row * LDO + d0), a1 = *reinterpret_cast<const float4*>(sO + row * LDO + d0 + 4);
                acc[c][0] = a0.x; acc[c][1] = a0.y; acc[c][2] = a0.z; acc[c][3] = a0.w;
                acc[c][4] = a1.x; acc[c][5] = a1.y; acc[c][6] = a1.z; acc[c][7] = a1.w;
                qp[c] = *reinterpret_cast<const uint4*>(sQ + row * LDS + d0);
            }
            auto load_kv = [&](int key, KVRegs& r) {
                const long long off = static_cast<long long>(key) * HD + head * D + 8 * sub;
#pragma unroll
                for (int c = 0; c < NCH; ++c) r.k[c] = __ldg(reinterpret_cast<const uint4*>(kcache + off + 8 * LPR * c));
#pragma unroll
                for (int c = 0; c < NCH; ++c) r.v[c] = __ldg(reinterpret_cast<const uint4*>(vcache + off + 8 * LPR * c));
            };
            auto fold = [&](const KVRegs& r) {
                float dot = 0.f;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&r.k[c]);
                    const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qp[c]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 kf = __bfloat1622float2(k2[j]), qf = __bfloat1622float2(q2[j]);
                        dot += qf.x * kf.x + qf.y * kf.y;
                    }
                }
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(row_mask, dot, o);
                const float sc = dot * scale;
                const float mn = fmaxf(m, sc);
                const float corr = m == -INFINITY ? 0.f : expf(m - mn);
                const float pw = expf(sc - mn);
                l = l * corr + pw;
                m = mn;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&r.v[c]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 vf = __bfloat1622float2(v2[j]);
                        acc[c][2 * j] = acc[c][2 * j] * corr + pw * vf.x;
                        acc[c][2 * j + 1] = acc[c][2 * j + 1] * corr + pw * vf.y;
                    }
                }
            };
            ATT_STAMP(10);
            KVRegs r0, r1, r2;
            if (n > 0) load_kv(keys[0], r0);
            if (n > 1) load_kv(keys[1], r1);
            for (int i = 0; i < n; i += 3) {
                if (i + 2 < n) load_kv(keys[i + 2], r2);
                fold(r0);
                if (i + 3 < n) load_kv(keys[i + 3], r0);
                if (i + 1 < n) fold(r1);
                if (i + 4 < n) load_kv(keys[i + 4], r1);
                if (i + 2 < n) fold(r2);
            }
            if (last) {
                // ---- normalise and store (bf16) ----
                if (tq < T) {
                    const float inv = l > 0.f ? 1.0f / l : 0.f;
                    __nv_bfloat16* dst = out + static_cast<long long>(tq) * HD + head * D;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        uint4 pk;
                        pk.x = pack_bf16(acc[c][0] * inv, acc[c][1] * inv);
                        pk.y = pack_bf16(acc[c][2] * inv, acc[c][3] * inv);
                        pk.z = pack_bf16(acc[c][4] * inv, acc[c][5] * inv);
                        pk.w = pack_bf16(acc[c][6] * inv, acc[c][7] * inv);
                        *reinterpret_cast<uint4*>(dst + 8 * (sub + LPR * c)) = pk;
                    }
                }
            } else {
                // more keys to come: the row state goes back to shared memory for the next round
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int d0 = 8 * (sub + LPR * c);
                    *reinterpret_cast<float4*>(sO + row * LDO + d0) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                    *reinterpret_cast<float4*>(sO + row * LDO + d0 + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
                }
                if (sub == 0) { sM[row] = m; sL[row] = l; }
            }
        }
        if (last) break;
        __syncwarp();
        if (lane < 16) fill_keys();
        __syncwarp();
    }
    ATT_STAMP(11);
}

#ifdef ATT_TIMING
extern "C" int atspeed_debug_att_stamps(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_att_stamps, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
}
#endif

int tree_attention(const __nv_bfloat16* q, const __nv_bfloat16* kcache, const __nv_bfloat16* vcache,
                   const BatchDesc& b, int T, int S, int n_heads, int head_dim, __nv_bfloat16* out, cudaStream_t st) {
    ATS_CHECK_ARG(T >= 1 && S >= 1 && S < 65536, "attention: T=%d S=%d", T, S);
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
