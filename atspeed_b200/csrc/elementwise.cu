// Row-wise kernels of the LLaMA forward between the tcgen05 GEMMs: embedding gather,
// RMSNorm, residual add, RoPE + KV-cache append, SiLU-gated product.  Each one also performs the
// fixed-order reduction of the GEMM's split-K fp32 slices and the bf16 roundings of the numerical
// contract documented in oracle/llama_ref.py (the places where an HF bf16 module rounds).
// All are HBM/L2-bound streaming kernels: one CTA per token row, 128-bit accesses where aligned.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace atspeed {

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

// a[k] = sum over the ns[k] partial-sum slices of the 4 columns at base[k], IN SLICE ORDER (deterministic, the order every
// consumer has always used), for N column groups at once: the loads of slice s of ALL groups are issued before any of them is
// used, slices 0 and 1 together.  Per-group loops (add_slices4 once per group) made a thread wait for one HBM round trip per
// group and slice -- 6 in a row in qkv_rope_append, which ran at 2.4 TB/s with every byte it needed addressable up front.
template <int N>
__device__ __forceinline__ void sum_slices(float4 (&a)[N], const float* const (&base)[N], const int (&ns)[N], long long split_stride) {
    float4 b[N];
    int smax = 1;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (ns[k] > 0) a[k] = __ldg(reinterpret_cast<const float4*>(base[k]));
        smax = ns[k] > smax ? ns[k] : smax;
    }
#pragma unroll
    for (int k = 0; k < N; ++k)
        if (ns[k] > 1) b[k] = __ldg(reinterpret_cast<const float4*>(base[k] + split_stride));
#pragma unroll
    for (int k = 0; k < N; ++k)
        if (ns[k] > 1) { a[k].x += b[k].x; a[k].y += b[k].y; a[k].z += b[k].z; a[k].w += b[k].w; }
    for (int s2 = 2; s2 < smax; ++s2) {
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (s2 < ns[k]) b[k] = __ldg(reinterpret_cast<const float4*>(base[k] + static_cast<long long>(s2) * split_stride));
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (s2 < ns[k]) { a[k].x += b[k].x; a[k].y += b[k].y; a[k].z += b[k].z; a[k].w += b[k].w; }
    }
}

__global__ void embed_rows_kernel(const __nv_bfloat16* __restrict__ table, const int* __restrict__ tok, int hidden,
                                  int vocab, __nv_bfloat16* __restrict__ h) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    int id = tok[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const __nv_bfloat16* src = table + static_cast<long long>(id) * hidden;
    __nv_bfloat16* dst = h + static_cast<long long>(t) * hidden;
    if ((hidden & 7) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = threadIdx.x; i < hidden; i += blockDim.x) dst[i] = src[i];
    }
}

int embed_rows(const __nv_bfloat16* table, const int* tok, int T, int hidden, int vocab, __nv_bfloat16* h,
               cudaStream_t st) {
    ATS_CUDA(launch_pdl(embed_rows_kernel, dim3(T), dim3(128), 0, st, table, tok, hidden, vocab, h));
    return ATS_OK;
}

// x[r] = bf16(g * bf16(h[row] * rsqrt(mean(h[row]^2) + eps)))
__global__ void rmsnorm_rows_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ g,
                                    int hidden, float eps, __nv_bfloat16* __restrict__ x,
                                    const int* __restrict__ row_index) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    const int row = row_index ? row_index[r] : r;
    const __nv_bfloat16* src = h + static_cast<long long>(row) * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        const float v = __bfloat162float(src[i]);
        ss += v * v;
    }
    ss = block_sum(ss, red);
    const float rstd = 1.0f / sqrtf(ss / static_cast<float>(hidden) + eps);
    __nv_bfloat16* dst = x + static_cast<long long>(r) * hidden;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        const float y = bf16_round(__bfloat162float(src[i]) * rstd);
        dst[i] = __float2bfloat16_rn(__bfloat162float(g[i]) * y);
    }
}

int rmsnorm_rows(const __nv_bfloat16* h, const __nv_bfloat16* g, int T, int hidden, float eps, __nv_bfloat16* x,
                 const int* row_index, cudaStream_t st) {
    ATS_CUDA(launch_pdl(rmsnorm_rows_kernel, dim3(T), dim3(256), 0, st, h, g, hidden, eps, x, row_index));
    return ATS_OK;
}

__global__ void residual_rmsnorm_kernel(__nv_bfloat16* __restrict__ h, const float* __restrict__ part, SplitMap sm,
                                        long long split_stride, int ldp, const __nv_bfloat16* __restrict__ g,
                                        int hidden, float eps, __nv_bfloat16* __restrict__ x) {
    extern __shared__ float row_buf[];   // hidden floats: the updated residual stream of this row
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    __nv_bfloat16* hrow = h + static_cast<long long>(t) * hidden;
    const float* prow = part + static_cast<long long>(t) * ldp;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        float acc = prow[i];
        const int splits = sm.slices(i);
        for (int s = 1; s < splits; ++s) acc += prow[s * split_stride + i];
        const float o = bf16_round(acc);
        const float hn = bf16_round(__bfloat162float(hrow[i]) + o);
        hrow[i] = __float2bfloat16_rn(hn);
        row_buf[i] = hn;
        ss += hn * hn;
    }
    if (g == nullptr) return;
    ss = block_sum(ss, red);
    const float rstd = 1.0f / sqrtf(ss / static_cast<float>(hidden) + eps);
    __nv_bfloat16* dst = x + static_cast<long long>(t) * hidden;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        const float y = bf16_round(row_buf[i] * rstd);
        dst[i] = __float2bfloat16_rn(__bfloat162float(g[i]) * y);
    }
}

// Vectorised variant: hidden % 4 == 0, ldp % 4 == 0, hidden <= THREADS * 4 * CH.  Each thread keeps its CH float4
// chunks of the row in registers, so the row is read once and all split slices are fetched with independent 128-bit
// loads (memory-level parallelism instead of a dependent scalar loop).
template <int THREADS, int CH>
__global__ void __launch_bounds__(THREADS)
residual_rmsnorm_vec_kernel(__nv_bfloat16* __restrict__ h, const float* __restrict__ part, SplitMap sm,
                            long long split_stride, int ldp, const __nv_bfloat16* __restrict__ g, int hidden, float eps,
                            __nv_bfloat16* __restrict__ x) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    __nv_bfloat16* hrow = h + static_cast<long long>(t) * hidden;
    const float* prow = part + static_cast<long long>(t) * ldp;
    const int nchunk = hidden >> 2;
    float4 v[CH];
    float ss = 0.f;
    float4 accs[CH];
    const float* bases[CH];
    int ns[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int i = threadIdx.x + c * THREADS;
        bases[c] = prow + 4 * (i < nchunk ? i : 0);
        ns[c] = i < nchunk ? sm.slices(4 * i) : 0;
    }
    sum_slices<CH>(accs, bases, ns, split_stride);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int i = threadIdx.x + c * THREADS;
        if (i < nchunk) {
            const float4 acc = accs[c];
            const uint2 hb = *reinterpret_cast<const uint2*>(hrow + 4 * i);
            const __nv_bfloat162 h01 = *reinterpret_cast<const __nv_bfloat162*>(&hb.x);
            const __nv_bfloat162 h23 = *reinterpret_cast<const __nv_bfloat162*>(&hb.y);
            float4 hn;
            hn.x = bf16_round(__bfloat162float(h01.x) + bf16_round(acc.x));
            hn.y = bf16_round(__bfloat162float(h01.y) + bf16_round(acc.y));
            hn.z = bf16_round(__bfloat162float(h23.x) + bf16_round(acc.z));
            hn.w = bf16_round(__bfloat162float(h23.y) + bf16_round(acc.w));
            __nv_bfloat162 o01 = __floats2bfloat162_rn(hn.x, hn.y), o23 = __floats2bfloat162_rn(hn.z, hn.w);
            uint2 ob;
            ob.x = *reinterpret_cast<uint32_t*>(&o01); ob.y = *reinterpret_cast<uint32_t*>(&o23);
            *reinterpret_cast<uint2*>(hrow + 4 * i) = ob;
            v[c] = hn;
            ss += hn.x * hn.x + hn.y * hn.y + hn.z * hn.z + hn.w * hn.w;
        }
    }
    if (g == nullptr) return;
    ss = block_sum(ss, red);
    const float rstd = 1.0f / sqrtf(ss / static_cast<float>(hidden) + eps);
    __nv_bfloat16* dst = x + static_cast<long long>(t) * hidden;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int i = threadIdx.x + c * THREADS;
        if (i < nchunk) {
            const uint2 gb = __ldg(reinterpret_cast<const uint2*>(g + 4 * i));
            const __nv_bfloat162 g01 = *reinterpret_cast<const __nv_bfloat162*>(&gb.x);
            const __nv_bfloat162 g23 = *reinterpret_cast<const __nv_bfloat162*>(&gb.y);
            __nv_bfloat162 o01 = __floats2bfloat162_rn(__bfloat162float(g01.x) * bf16_round(v[c].x * rstd),
                                                       __bfloat162float(g01.y) * bf16_round(v[c].y * rstd));
            __nv_bfloat162 o23 = __floats2bfloat162_rn(__bfloat162float(g23.x) * bf16_round(v[c].z * rstd),
                                                       __bfloat162float(g23.y) * bf16_round(v[c].w * rstd));
            uint2 ob;
            ob.x = *reinterpret_cast<uint32_t*>(&o01); ob.y = *reinterpret_cast<uint32_t*>(&o23);
            *reinterpret_cast<uint2*>(dst + 4 * i) = ob;
        }
    }
}

int residual_rmsnorm(__nv_bfloat16* h, const float* part, const SplitMap& sm, long long split_stride, int ldp,
                     const __nv_bfloat16* g, int T, int hidden, float eps, __nv_bfloat16* x, cudaStream_t st) {
    if ((hidden & 3) == 0 && (ldp & 3) == 0 && (split_stride & 3) == 0 && hidden <= 512 * 4 * 4 &&
        (reinterpret_cast<uintptr_t>(part) & 15) == 0) {
        if (hidden <= 256 * 4)
            ATS_CUDA(launch_pdl(residual_rmsnorm_vec_kernel<256, 1>, dim3(T), dim3(256), 0, st, h, part, sm, split_stride, ldp,
                                g, hidden, eps, x));
        else if (hidden <= 512 * 4 * 2)
            ATS_CUDA(launch_pdl(residual_rmsnorm_vec_kernel<512, 2>, dim3(T), dim3(512), 0, st, h, part, sm, split_stride, ldp,
                                g, hidden, eps, x));
        else
            ATS_CUDA(launch_pdl(residual_rmsnorm_vec_kernel<512, 4>, dim3(T), dim3(512), 0, st, h, part, sm, split_stride, ldp,
                                g, hidden, eps, x));
        return ATS_OK;
    }
    ATS_CHECK_ARG(hidden * 4 <= 96 * 1024, "residual_rmsnorm: hidden=%d too large", hidden);
    static std::once_flag once;          // sessions on several host threads (bench.py lanes) may arrive here together
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, []() {
        attr_err = cudaFuncSetAttribute(residual_rmsnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    ATS_CUDA(attr_err);
    ATS_CUDA(launch_pdl(residual_rmsnorm_kernel, dim3(T), dim3(256), hidden * sizeof(float), st, h, part, sm, split_stride,
                        ldp, g, hidden, eps, x));
    return ATS_OK;
}

__global__ void qkv_rope_append_kernel(const float* __restrict__ part, SplitMap sm, long long split_stride, int ldp,
                                       const int* __restrict__ pos, const int* __restrict__ slot, int n_heads,
                                       int head_dim, const float* __restrict__ rope_cos,
                                       const float* __restrict__ rope_sin, int max_pos,
                                       __nv_bfloat16* __restrict__ qbuf, __nv_bfloat16* __restrict__ kcache,
                                       __nv_bfloat16* __restrict__ vcache, const int* __restrict__ tok_user, CohortKV ckv) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    const int HD = n_heads * head_dim, half = head_dim >> 1;
    int p = pos[t];
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
    const long long srow = static_cast<long long>(slot[t]) * HD + (tok_user ? ckv.kv_off[tok_user[t]] : 0);
    const float* prow = part + static_cast<long long>(t) * ldp;
    const float* ct = rope_cos + static_cast<long long>(p) * half;
    const float* sn = rope_sin + static_cast<long long>(p) * half;
    for (int e = threadIdx.x; e < n_heads * half; e += blockDim.x) {
        const int hd = e / half, i = e - hd * half;
        const int c0 = hd * head_dim + i, c1 = c0 + half;
        float q0 = prow[c0], q1 = prow[c1], k0 = prow[HD + c0], k1 = prow[HD + c1];
        float v0 = prow[2 * HD + c0], v1 = prow[2 * HD + c1];
        const int cols[6] = {c0, c1, HD + c0, HD + c1, 2 * HD + c0, 2 * HD + c1};
        float* vals[6] = {&q0, &q1, &k0, &k1, &v0, &v1};
        for (int k = 0; k < 6; ++k) {
            const int splits = sm.slices(cols[k]);
            for (int s = 1; s < splits; ++s) *vals[k] += prow[s * split_stride + cols[k]];
        }
        q0 = bf16_round(q0); q1 = bf16_round(q1); k0 = bf16_round(k0); k1 = bf16_round(k1);
        const float c = ct[i], s_ = sn[i];
        // HF apply_rotary_pos_emb in bf16: (x * cos) + (rotate_half(x) * sin), every op rounded
        const float qo0 = bf16_round(bf16_round(q0 * c) + bf16_round(-q1 * s_));
        const float qo1 = bf16_round(bf16_round(q1 * c) + bf16_round(q0 * s_));
        const float ko0 = bf16_round(bf16_round(k0 * c) + bf16_round(-k1 * s_));
        const float ko1 = bf16_round(bf16_round(k1 * c) + bf16_round(k0 * s_));
        qbuf[static_cast<long long>(t) * HD + c0] = __float2bfloat16_rn(qo0);
        qbuf[static_cast<long long>(t) * HD + c1] = __float2bfloat16_rn(qo1);
        kcache[srow + c0] = __float2bfloat16_rn(ko0);
        kcache[srow + c1] = __float2bfloat16_rn(ko1);
        vcache[srow + c0] = __float2bfloat16_rn(v0);
        vcache[srow + c1] = __float2bfloat16_rn(v1);
    }
}

// Vectorised variant: one thread = 4 consecutive rotary pairs (i..i+3, i+half..i+half+3) of one head.
__global__ void __launch_bounds__(256)
qkv_rope_append_vec_kernel(const float* __restrict__ part, SplitMap sm, long long split_stride, int ldp,
                           const int* __restrict__ pos, const int* __restrict__ slot, int n_heads, int head_dim,
                           const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int max_pos,
                           __nv_bfloat16* __restrict__ qbuf, __nv_bfloat16* __restrict__ kcache,
                           __nv_bfloat16* __restrict__ vcache, const int* __restrict__ tok_user, CohortKV ckv) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    const int HD = n_heads * head_dim, half = head_dim >> 1, q4 = half >> 2;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= n_heads * q4) return;
    const int hd = e / q4, i = (e - hd * q4) * 4;
    int p = pos[t];
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
    const long long srow = static_cast<long long>(slot[t]) * HD + (tok_user ? ckv.kv_off[tok_user[t]] : 0);
    const float* prow = part + static_cast<long long>(t) * ldp;
    const int c0 = hd * head_dim + i, c1 = c0 + half;
    float4 a[6];
    const int cols[6] = {c0, c1, HD + c0, HD + c1, 2 * HD + c0, 2 * HD + c1};
    const float* const bases[6] = {prow + cols[0], prow + cols[1], prow + cols[2], prow + cols[3], prow + cols[4], prow + cols[5]};
    // c0 and c1 = c0 + head_dim / 2 lie in the same head, hence (heads never straddle a GEMM tile) in the same tile
    const int nq = sm.slices(cols[0]), nk = sm.slices(cols[2]), nv = sm.slices(cols[4]);
    const int ns[6] = {nq, nq, nk, nk, nv, nv};
    sum_slices<6>(a, bases, ns, split_stride);
    const float4 cs = __ldg(reinterpret_cast<const float4*>(rope_cos + static_cast<long long>(p) * half + i));
    const float4 sn = __ldg(reinterpret_cast<const float4*>(rope_sin + static_cast<long long>(p) * half + i));
    auto rope = [](float x0, float x1, float c, float s_, float& o0, float& o1) {
        x0 = bf16_round(x0); x1 = bf16_round(x1);
        o0 = bf16_round(bf16_round(x0 * c) + bf16_round(-x1 * s_));
        o1 = bf16_round(bf16_round(x1 * c) + bf16_round(x0 * s_));
    };
    auto pack4 = [](float x, float y, float z, float w) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(x, y), hi = __floats2bfloat162_rn(z, w);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&lo); r.y = *reinterpret_cast<uint32_t*>(&hi);
        return r;
    };
    float4 q0, q1, k0, k1;
    rope(a[0].x, a[1].x, cs.x, sn.x, q0.x, q1.x); rope(a[0].y, a[1].y, cs.y, sn.y, q0.y, q1.y);
    rope(a[0].z, a[1].z, cs.z, sn.z, q0.z, q1.z); rope(a[0].w, a[1].w, cs.w, sn.w, q0.w, q1.w);
    rope(a[2].x, a[3].x, cs.x, sn.x, k0.x, k1.x); rope(a[2].y, a[3].y, cs.y, sn.y, k0.y, k1.y);
    rope(a[2].z, a[3].z, cs.z, sn.z, k0.z, k1.z); rope(a[2].w, a[3].w, cs.w, sn.w, k0.w, k1.w);
    *reinterpret_cast<uint2*>(qbuf + static_cast<long long>(t) * HD + c0) = pack4(q0.x, q0.y, q0.z, q0.w);
    *reinterpret_cast<uint2*>(qbuf + static_cast<long long>(t) * HD + c1) = pack4(q1.x, q1.y, q1.z, q1.w);
    *reinterpret_cast<uint2*>(kcache + srow + c0) = pack4(k0.x, k0.y, k0.z, k0.w);
    *reinterpret_cast<uint2*>(kcache + srow + c1) = pack4(k1.x, k1.y, k1.z, k1.w);
    *reinterpret_cast<uint2*>(vcache + srow + c0) = pack4(a[4].x, a[4].y, a[4].z, a[4].w);
    *reinterpret_cast<uint2*>(vcache + srow + c1) = pack4(a[5].x, a[5].y, a[5].z, a[5].w);
}

int qkv_rope_append(const float* part, const SplitMap& sm, long long split_stride, int ldp, const BatchDesc& b, int T,
                    int n_heads, int head_dim, const float* rope_cos, const float* rope_sin, int max_pos,
                    __nv_bfloat16* qbuf, __nv_bfloat16* kcache, __nv_bfloat16* vcache, cudaStream_t st) {
    ATS_CHECK_ARG((head_dim & 1) == 0, "head_dim=%d must be even", head_dim);
    if ((head_dim & 7) == 0 && (ldp & 3) == 0 && (split_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(rope_cos) & 15) == 0 && (reinterpret_cast<uintptr_t>(rope_sin) & 15) == 0) {
        const int work = n_heads * (head_dim >> 3);
        dim3 grid(T, (work + 255) / 256);
        ATS_CUDA(launch_pdl(qkv_rope_append_vec_kernel, grid, dim3(256), 0, st, part, sm, split_stride, ldp, b.pos, b.slot,
                            n_heads, head_dim, rope_cos, rope_sin, max_pos, qbuf, kcache, vcache, b.tok_user, b.ckv));
        return ATS_OK;
    }
    ATS_CUDA(launch_pdl(qkv_rope_append_kernel, dim3(T), dim3(256), 0, st, part, sm, split_stride, ldp, b.pos, b.slot,
                        n_heads, head_dim, rope_cos, rope_sin, max_pos, qbuf, kcache, vcache, b.tok_user, b.ckv));
    return ATS_OK;
}

__global__ void silu_mul_kernel(const float* __restrict__ part, SplitMap sm, long long split_stride, int ldp, int mlp,
                                __nv_bfloat16* __restrict__ m) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    const float* prow = part + static_cast<long long>(t) * ldp;
    __nv_bfloat16* dst = m + static_cast<long long>(t) * mlp;
    for (int i = threadIdx.x; i < mlp; i += blockDim.x) {
        float g = prow[i], u = prow[mlp + i];
        const int sg = sm.slices(i), su = sm.slices(mlp + i);
        for (int s = 1; s < sg; ++s) g += prow[s * split_stride + i];
        for (int s = 1; s < su; ++s) u += prow[s * split_stride + mlp + i];
        dst[i] = __float2bfloat16_rn(silu_mul_bf16(g, u));
    }
}

// One thread = 4 consecutive columns of NR consecutive token rows (same columns: the slice count and the column arithmetic are
// computed once).  All 2 * NR * min(slices, 2) 16-byte loads of a thread are issued before the first is used: with one row per
// thread the kernel kept ~50 KB in flight per SM and streamed 2.2 TB/s (tools/rowwise_bench.py); HBM at full rate needs ~100 KB.
template <int NR>
__global__ void __launch_bounds__(256, 3)
silu_mul_vec_kernel(const float* __restrict__ part, SplitMap sm, long long split_stride, int ldp, int T, int mlp,
                    __nv_bfloat16* __restrict__ m) {
    pdl_launch_dependents();
    pdl_wait();
    const int t0 = blockIdx.x * NR;
    const int i = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if (i >= mlp) return;
    const int sg = sm.slices(i), su = sm.slices(mlp + i);
    const float* p0 = part + static_cast<long long>(t0) * ldp + i;
    float4 gu[2 * NR];                        // gu[2r] = gate, gu[2r + 1] = up of row t0 + r
    const float* bases[2 * NR];
    int ns[2 * NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const bool ok = t0 + r < T;
        bases[2 * r] = p0 + static_cast<long long>(ok ? r : 0) * ldp;
        bases[2 * r + 1] = bases[2 * r] + mlp;
        ns[2 * r] = ok ? sg : 0;
        ns[2 * r + 1] = ok ? su : 0;
    }
    sum_slices<2 * NR>(gu, bases, ns, split_stride);
    auto f = [](float gg, float uu) { return silu_mul_bf16(gg, uu); };
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (t0 + r >= T) break;
        const float4 g = gu[2 * r], u = gu[2 * r + 1];
        __nv_bfloat162 lo = __floats2bfloat162_rn(f(g.x, u.x), f(g.y, u.y)), hi = __floats2bfloat162_rn(f(g.z, u.z), f(g.w, u.w));
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(m + static_cast<long long>(t0 + r) * mlp + i) = o;
    }
}

int silu_mul(const float* part, const SplitMap& sm, long long split_stride, int ldp, int T, int mlp, __nv_bfloat16* m,
             cudaStream_t st) {
    if ((mlp & 3) == 0 && (ldp & 3) == 0 && (split_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0) {
        constexpr int NR = 2;
        ATS_CUDA(launch_pdl(silu_mul_vec_kernel<NR>, dim3((T + NR - 1) / NR, (mlp / 4 + 255) / 256), dim3(256), 0, st, part, sm, split_stride,
                            ldp, T, mlp, m));
        return ATS_OK;
    }
    ATS_CUDA(launch_pdl(silu_mul_kernel, dim3(T), dim3(256), 0, st, part, sm, split_stride, ldp, mlp, m));
    return ATS_OK;
}

__global__ void reduce_slices_kernel(const float* __restrict__ part, SplitMap sm, long long split_stride, int ldp, int cols,
                                     float* __restrict__ out, int ldo) {
    const int t = blockIdx.x;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float acc = part[static_cast<long long>(t) * ldp + c];
        const int n = sm.slices(c);
        for (int s = 1; s < n; ++s) acc += part[s * split_stride + static_cast<long long>(t) * ldp + c];
        out[static_cast<long long>(t) * ldo + c] = acc;
    }
}

int reduce_slices(const float* part, const SplitMap& sm, long long split_stride, int ldp, int T, int cols, float* out,
                  int ldo, cudaStream_t st) {
    reduce_slices_kernel<<<T, 256, 0, st>>>(part, sm, split_stride, ldp, cols, out, ldo);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed

// Diagnostics (tools/rowwise_bench.py): microseconds per launch of one row-wise consumer kernel at a cohort-forward size, every
// column held in `slices` partial-sum slices, inputs cycled through buffers larger than L2 so each launch reads HBM.
// kind 0: qkv_rope_append, 1: silu_mul, 2: residual_rmsnorm.
extern "C" int atspeed_debug_rowwise_us(int32_t kind, int32_t T, int32_t hidden, int32_t mlp, int32_t n_heads, int32_t slices,
                                        int32_t iters, float* us_out, void* stream) {
    using namespace atspeed;
    ATS_CHECK_ARG(kind >= 0 && kind <= 2 && T >= 1 && slices >= 1 && iters >= 1 && us_out && n_heads >= 1 && hidden % n_heads == 0,
                  "rowwise bench: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int head_dim = hidden / n_heads;
    const int cols = kind == 0 ? 3 * hidden : (kind == 1 ? 2 * mlp : hidden);
    const long long stride = static_cast<long long>(T) * cols;
    const size_t part_bytes = sizeof(float) * stride * slices;
    const int nbuf = static_cast<int>((300ull << 20) / part_bytes) + 2;
    SplitMap sm;
    memset(&sm, 0, sizeof(sm));
    sm.n = 1; sm.colbase[0] = 0; sm.colbase[1] = sm.colbase[2] = 0x7fffffff; sm.BM = 256; sm.U = 8; sm.KB = 8 * slices; sm.tps = 1;
    sm.bm_shift = 8; sm.tab_n = (cols + 255) / 256 <= SPLIT_TAB ? (cols + 255) / 256 : 0;
    for (int t = 0; t < sm.tab_n; ++t) sm.tab[t] = static_cast<unsigned char>(slices);
    struct DevBufs {                      // freed on every return path
        std::vector<void*> p;
        ~DevBufs() { for (void* q : p) cudaFree(q); }
        cudaError_t take(void** out, size_t bytes) {
            cudaError_t e = cudaMalloc(out, bytes);
            if (e == cudaSuccess) p.push_back(*out);
            return e;
        }
    } bufs;
    float* part = nullptr; __nv_bfloat16 *h = nullptr, *x = nullptr, *g = nullptr, *q = nullptr, *kv = nullptr, *mm = nullptr;
    float* rope = nullptr; int* meta = nullptr;
    const int max_pos = 1024, S = T + 8;
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&part), part_bytes * nbuf));
    ATS_CUDA(cudaMemsetAsync(part, 0, part_bytes * nbuf, st));
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&h), sizeof(__nv_bfloat16) * T * hidden * 2));
    x = h + static_cast<long long>(T) * hidden;
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&g), sizeof(__nv_bfloat16) * hidden));
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&q), sizeof(__nv_bfloat16) * T * hidden));
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&kv), sizeof(__nv_bfloat16) * 2 * S * hidden));
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&mm), sizeof(__nv_bfloat16) * T * (mlp > 0 ? mlp : 1)));
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&rope), sizeof(float) * 2 * max_pos * (head_dim / 2)));
    ATS_CUDA(bufs.take(reinterpret_cast<void**>(&meta), sizeof(int) * 2 * T));
    ATS_CUDA(cudaMemsetAsync(h, 0, sizeof(__nv_bfloat16) * T * hidden * 2, st));
    ATS_CUDA(cudaMemsetAsync(g, 0, sizeof(__nv_bfloat16) * hidden, st));
    ATS_CUDA(cudaMemsetAsync(rope, 0, sizeof(float) * 2 * max_pos * (head_dim / 2), st));
    std::vector<int> hm(2 * T);
    for (int t = 0; t < T; ++t) { hm[t] = t % max_pos; hm[T + t] = t; }
    ATS_CUDA(cudaMemcpyAsync(meta, hm.data(), sizeof(int) * 2 * T, cudaMemcpyHostToDevice, st));
    BatchDesc b;
    memset(&b, 0, sizeof(b));
    b.pos = meta; b.slot = meta + T;
    auto run = [&](int i) -> int {
        const float* pp = part + static_cast<long long>(i % nbuf) * stride * slices;
        if (kind == 0)
            return qkv_rope_append(pp, sm, stride, cols, b, T, n_heads, head_dim, rope, rope + max_pos * (head_dim / 2), max_pos, q, kv,
                                   kv + static_cast<long long>(S) * hidden, st);
        if (kind == 1) return silu_mul(pp, sm, stride, cols, T, mlp, mm, st);
        return residual_rmsnorm(h, pp, sm, stride, cols, g, T, hidden, 1e-5f, x, st);
    };
    for (int i = 0; i < 3; ++i) ATS_TRY(run(i));
    cudaEvent_t e0, e1;
    ATS_CUDA(cudaEventCreate(&e0));
    ATS_CUDA(cudaEventCreate(&e1));
    ATS_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) ATS_TRY(run(i + 3));
    ATS_CUDA(cudaEventRecord(e1, st));
    ATS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ATS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *us_out = ms * 1e3f / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ATS_OK;
}

