// fp32 forward of the LLaMA decoder: the exact-parity mode of the path.
//
// The bf16 production forward (gemm.cu / attention.cu / elementwise.cu) cannot reproduce the reference's ranked lists
// bit for bit: tensor-core accumulation order flips individual bf16 roundings and beam search prunes discontinuously
// (hard part 1).  The north star asks for exact lists and scores within 1e-3 "in fp32", so models handed over
// as fp32 (atspeed_model_desc.weights_f32 = 1) run through these plain SIMT fp32 kernels instead: same batch
// descriptors, same KV-slot layout, same tree masks, every reduction in fp32 -- only the summation order differs from
// HF's fp32 modules (~1e-6 relative), which leaves near-ties measure-zero.  Used for the tiny parity configuration
// (BASELINE.json configs[0]); it is not a performance path and not a fallback (bf16 models never come here).
#include "common.cuh"
#include "kernels.h"

namespace atspeed {

__global__ void f32_embed_kernel(const float* __restrict__ table, const int* __restrict__ tok, int hidden, int vocab,
                                 float* __restrict__ h) {
    const int t = blockIdx.x;
    int id = tok[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    for (int i = threadIdx.x; i < hidden; i += blockDim.x)
        h[static_cast<long long>(t) * hidden + i] = table[static_cast<long long>(id) * hidden + i];
}

// x[r] = g * (h[row] * rsqrt(mean(h[row]^2) + eps))      (LlamaRMSNorm in fp32)
__global__ void f32_rmsnorm_kernel(const float* __restrict__ h, const float* __restrict__ g, int hidden, float eps,
                                   float* __restrict__ x, const int* __restrict__ row_index) {
    __shared__ float red[32];
    const int r = blockIdx.x;
    const int row = row_index ? row_index[r] : r;
    const float* src = h + static_cast<long long>(row) * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) ss += src[i] * src[i];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    const float rstd = rsqrtf(tot / static_cast<float>(hidden) + eps);
    for (int i = threadIdx.x; i < hidden; i += blockDim.x)
        x[static_cast<long long>(r) * hidden + i] = g[i] * (src[i] * rstd);
}

// out[t][n] (+)= sum_k x[t][k] * w[n][k]; 32 x 32 output tile per CTA, k in chunks of 32 through shared memory
template <bool ACCUM>
__global__ void __launch_bounds__(256) f32_gemm_kernel(const float* __restrict__ x, const float* __restrict__ w, int T, int N,
                                                       int K, float* __restrict__ out, int ldo) {
    __shared__ float sx[32][33], sw[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8 threads, 4 outputs each
    const int t0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = ty; i < 32; i += 8) {
            const int k = k0 + tx;
            sx[i][tx] = (t0 + i < T && k < K) ? x[static_cast<long long>(t0 + i) * K + k] : 0.f;
            sw[i][tx] = (n0 + i < N && k < K) ? w[static_cast<long long>(n0 + i) * K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float wv = sw[tx][k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += sx[ty + 8 * j][k] * wv;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = t0 + ty + 8 * j, n = n0 + tx;
        if (t < T && n < N) {
            float* o = out + static_cast<long long>(t) * ldo + n;
            *o = ACCUM ? *o + acc[j] : acc[j];
        }
    }
}

// RoPE (HF rotate_half form) on q and k, q -> qbuf, k/v -> cache rows slot[t]; qkv: [T][3*HD]
__global__ void f32_rope_append_kernel(const float* __restrict__ qkv, const int* __restrict__ pos, const int* __restrict__ slot,
                                       int n_heads, int head_dim, const float* __restrict__ rope_cos,
                                       const float* __restrict__ rope_sin, int max_pos, float* __restrict__ qbuf,
                                       float* __restrict__ kcache, float* __restrict__ vcache,
                                       const int* __restrict__ tok_user, CohortKV ckv) {
    const int t = blockIdx.x;
    const int HD = n_heads * head_dim, half = head_dim >> 1;
    int p = pos[t];
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
    const long long srow = static_cast<long long>(slot[t]) * HD + (tok_user ? ckv.kv_off[tok_user[t]] : 0);   // cohort: own cache
    const float* row = qkv + static_cast<long long>(t) * 3 * HD;
    for (int e = threadIdx.x; e < n_heads * half; e += blockDim.x) {
        const int hd = e / half, i = e - hd * half;
        const int c0 = hd * head_dim + i, c1 = c0 + half;
        const float c = rope_cos[static_cast<long long>(p) * half + i], s = rope_sin[static_cast<long long>(p) * half + i];
        const float q0 = row[c0], q1 = row[c1], k0 = row[HD + c0], k1 = row[HD + c1];
        qbuf[static_cast<long long>(t) * HD + c0] = q0 * c - q1 * s;
        qbuf[static_cast<long long>(t) * HD + c1] = q1 * c + q0 * s;
        kcache[srow + c0] = k0 * c - k1 * s;
        kcache[srow + c1] = k1 * c + k0 * s;
        vcache[srow + c0] = row[2 * HD + c0];
        vcache[srow + c1] = row[2 * HD + c1];
    }
}

// one warp per (token, head): two passes over the visible keys (max, then exp-sum and PV) -- fp32 throughout
__global__ void __launch_bounds__(128) f32_tree_attention_kernel(const float* __restrict__ q, const float* __restrict__ kcache,
                                                                 const float* __restrict__ vcache,
                                                                 const int* __restrict__ prefix_len,
                                                                 const uint32_t* __restrict__ vis, int vis_base, int T,
                                                                 int S, int n_heads, int D, float scale,
                                                                 float* __restrict__ out, const int* __restrict__ tok_user,
                                                                 CohortKV ckv) {
    extern __shared__ float f32_att_smem[];                  // [4 warps][S] scores
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * 4 + warp;
    if (item >= T * n_heads) return;
    const int t = item / n_heads, head = item - t * n_heads;
    const int HD = n_heads * D;
    float* sc = f32_att_smem + static_cast<size_t>(warp) * S;   // S = the launch's largest extent
    if (tok_user) {
        // cohort forward: the token attends its own user's cache, extent and prompt length (block-diagonal tree mask)
        const int u = tok_user[t];
        kcache += ckv.kv_off[u];
        vcache += ckv.kv_off[u];
        S = ckv.S[u];
        vis_base = ckv.vis_base[u];
    }
    const float* qv = q + static_cast<long long>(t) * HD + head * D;
    const int pl = prefix_len[t];
    const uint32_t* vrow = vis + static_cast<long long>(t) * VIS_WORDS;
    float mx = -INFINITY;
    for (int key = 0; key < S; ++key) {
        bool ok = key < pl;
        if (!ok && key >= vis_base) {
            const int b = key - vis_base;
            ok = (b >> 5) < VIS_WORDS && ((vrow[b >> 5] >> (b & 31)) & 1u);
        }
        float s = -INFINITY;
        if (ok) {                                              // warp-uniform
            const float* kv = kcache + static_cast<long long>(key) * HD + head * D;
            float d = 0.f;
            for (int i = lane; i < D; i += 32) d += qv[i] * kv[i];
            s = warp_sum(d) * scale;
        }
        if (lane == 0) sc[key] = s;
        mx = fmaxf(mx, s);
    }
    __syncwarp();
    float l = 0.f;
    for (int key = lane; key < S; key += 32) {
        const float e = sc[key] == -INFINITY ? 0.f : expf(sc[key] - mx);
        sc[key] = e;
        l += e;
    }
    l = warp_sum(l);
    __syncwarp();
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    for (int i = lane; i < D; i += 32) {
        float acc = 0.f;
        for (int key = 0; key < S; ++key) {
            const float p = sc[key];
            if (p != 0.f) acc += p * vcache[static_cast<long long>(key) * HD + head * D + i];
        }
        out[static_cast<long long>(t) * HD + head * D + i] = acc * inv;
    }
}

__global__ void f32_silu_mul_kernel(const float* __restrict__ gu, int mlp, float* __restrict__ m) {
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < mlp; i += blockDim.x) {
        const float g = gu[static_cast<long long>(t) * 2 * mlp + i], u = gu[static_cast<long long>(t) * 2 * mlp + mlp + i];
        m[static_cast<long long>(t) * mlp + i] = (g / (1.0f + expf(-g))) * u;
    }
}

int f32_embed(const float* table, const int* tok, int T, int hidden, int vocab, float* h, cudaStream_t st) {
    f32_embed_kernel<<<T, 128, 0, st>>>(table, tok, hidden, vocab, h);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}
int f32_rmsnorm(const float* h, const float* g, int T, int hidden, float eps, float* x, const int* row_index, cudaStream_t st) {
    f32_rmsnorm_kernel<<<T, 128, 0, st>>>(h, g, hidden, eps, x, row_index);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}
int f32_gemm(const float* x, const float* w, int T, int N, int K, float* out, int ldo, bool accumulate, cudaStream_t st) {
    dim3 grid((N + 31) / 32, (T + 31) / 32);
    if (accumulate) f32_gemm_kernel<true><<<grid, 256, 0, st>>>(x, w, T, N, K, out, ldo);
    else f32_gemm_kernel<false><<<grid, 256, 0, st>>>(x, w, T, N, K, out, ldo);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}
int f32_rope_append(const float* qkv, const BatchDesc& b, int T, int n_heads, int head_dim, const float* rope_cos,
                    const float* rope_sin, int max_pos, float* qbuf, float* kcache, float* vcache, cudaStream_t st) {
    f32_rope_append_kernel<<<T, 128, 0, st>>>(qkv, b.pos, b.slot, n_heads, head_dim, rope_cos, rope_sin, max_pos, qbuf, kcache,
                                              vcache, b.tok_user, b.ckv);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}
int f32_tree_attention(const float* q, const float* kcache, const float* vcache, const BatchDesc& b, int T, int S, int n_heads,
                       int head_dim, float* out, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(4) * S * sizeof(float);
    ATS_CHECK_ARG(smem <= 48 * 1024, "fp32 attention: S=%d too long for the parity path", S);
    const int items = T * n_heads;
    f32_tree_attention_kernel<<<(items + 3) / 4, 128, smem, st>>>(q, kcache, vcache, b.prefix_len, b.vis, b.vis_base, T, S,
                                                                  n_heads, head_dim, 1.0f / sqrtf(static_cast<float>(head_dim)),
                                                                  out, b.tok_user, b.ckv);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}
int f32_silu_mul(const float* gu, int T, int mlp, float* m, cudaStream_t st) {
    f32_silu_mul_kernel<<<T, 256, 0, st>>>(gu, mlp, m);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed
