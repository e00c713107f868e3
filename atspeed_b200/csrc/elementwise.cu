// Row-wise kernels of the LLaMA forward between the tcgen05 GEMMs: embedding gather,
// RMSNorm, residual add, RoPE + KV-cache append, SiLU-gated product.  Each one also performs the
// fixed-order reduction of the GEMM's split-K fp32 slices and the bf16 roundings of the numerical
// contract documented in oracle/llama_ref.py (the places where an HF bf16 module rounds).
// All are HBM/L2-bound streaming kernels: one CTA per token row, 128-bit accesses where aligned.
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

__global__ void embed_rows_kernel(const __nv_bfloat16* __restrict__ table, const int* __restrict__ tok, int hidden,
                                  int vocab, __nv_bfloat16* __restrict__ h) {
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
    embed_rows_kernel<<<T, 128, 0, st>>>(table, tok, hidden, vocab, h);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// x[r] = bf16(g * bf16(h[row] * rsqrt(mean(h[row]^2) + eps)))
__global__ void rmsnorm_rows_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ g,
                                    int hidden, float eps, __nv_bfloat16* __restrict__ x,
                                    const int* __restrict__ row_index) {
    __shared__ float red[32];
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
    rmsnorm_rows_kernel<<<T, 256, 0, st>>>(h, g, hidden, eps, x, row_index);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void residual_rmsnorm_kernel(__nv_bfloat16* __restrict__ h, const float* __restrict__ part, int splits,
                                        long long split_stride, int ldp, const __nv_bfloat16* __restrict__ g,
                                        int hidden, float eps, __nv_bfloat16* __restrict__ x) {
    extern __shared__ float row_buf[];   // hidden floats: the updated residual stream of this row
    __shared__ float red[32];
    const int t = blockIdx.x;
    __nv_bfloat16* hrow = h + static_cast<long long>(t) * hidden;
    const float* prow = part + static_cast<long long>(t) * ldp;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        float acc = prow[i];
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

int residual_rmsnorm(__nv_bfloat16* h, const float* part, int splits, long long split_stride, int ldp,
                     const __nv_bfloat16* g, int T, int hidden, float eps, __nv_bfloat16* x, cudaStream_t st) {
    ATS_CHECK_ARG(hidden * 4 <= 96 * 1024, "residual_rmsnorm: hidden=%d too large", hidden);
    static bool attr_set = false;
    if (!attr_set) {
        ATS_CUDA(cudaFuncSetAttribute(residual_rmsnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    residual_rmsnorm_kernel<<<T, 256, hidden * sizeof(float), st>>>(h, part, splits, split_stride, ldp, g, hidden,
                                                                     eps, x);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void qkv_rope_append_kernel(const float* __restrict__ part, int splits, long long split_stride, int ldp,
                                       const int* __restrict__ pos, const int* __restrict__ slot, int n_heads,
                                       int head_dim, const float* __restrict__ rope_cos,
                                       const float* __restrict__ rope_sin, int max_pos,
                                       __nv_bfloat16* __restrict__ qbuf, __nv_bfloat16* __restrict__ kcache,
                                       __nv_bfloat16* __restrict__ vcache) {
    const int t = blockIdx.x;
    const int HD = n_heads * head_dim, half = head_dim >> 1;
    int p = pos[t];
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
    const long long srow = static_cast<long long>(slot[t]) * HD;
    const float* prow = part + static_cast<long long>(t) * ldp;
    const float* ct = rope_cos + static_cast<long long>(p) * half;
    const float* sn = rope_sin + static_cast<long long>(p) * half;
    for (int e = threadIdx.x; e < n_heads * half; e += blockDim.x) {
        const int hd = e / half, i = e - hd * half;
        const int c0 = hd * head_dim + i, c1 = c0 + half;
        float q0 = prow[c0], q1 = prow[c1], k0 = prow[HD + c0], k1 = prow[HD + c1];
        float v0 = prow[2 * HD + c0], v1 = prow[2 * HD + c1];
        for (int s = 1; s < splits; ++s) {
            const float* ps = prow + s * split_stride;
            q0 += ps[c0]; q1 += ps[c1]; k0 += ps[HD + c0]; k1 += ps[HD + c1];
            v0 += ps[2 * HD + c0]; v1 += ps[2 * HD + c1];
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

int qkv_rope_append(const float* part, int splits, long long split_stride, int ldp, const BatchDesc& b, int T,
                    int n_heads, int head_dim, const float* rope_cos, const float* rope_sin, int max_pos,
                    __nv_bfloat16* qbuf, __nv_bfloat16* kcache, __nv_bfloat16* vcache, cudaStream_t st) {
    ATS_CHECK_ARG((head_dim & 1) == 0, "head_dim=%d must be even", head_dim);
    qkv_rope_append_kernel<<<T, 256, 0, st>>>(part, splits, split_stride, ldp, b.pos, b.slot, n_heads, head_dim,
                                              rope_cos, rope_sin, max_pos, qbuf, kcache, vcache);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void silu_mul_kernel(const float* __restrict__ part, int splits, long long split_stride, int ldp, int mlp,
                                __nv_bfloat16* __restrict__ m) {
    const int t = blockIdx.x;
    const float* prow = part + static_cast<long long>(t) * ldp;
    __nv_bfloat16* dst = m + static_cast<long long>(t) * mlp;
    for (int i = threadIdx.x; i < mlp; i += blockDim.x) {
        float g = prow[i], u = prow[mlp + i];
        for (int s = 1; s < splits; ++s) { g += prow[s * split_stride + i]; u += prow[s * split_stride + mlp + i]; }
        g = bf16_round(g); u = bf16_round(u);
        const float a = bf16_round(g / (1.0f + expf(-g)));
        dst[i] = __float2bfloat16_rn(a * u);
    }
}

int silu_mul(const float* part, int splits, long long split_stride, int ldp, int T, int mlp, __nv_bfloat16* m,
             cudaStream_t st) {
    silu_mul_kernel<<<T, 256, 0, st>>>(part, splits, split_stride, ldp, mlp, m);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed
