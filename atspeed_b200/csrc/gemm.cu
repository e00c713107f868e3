// Weight-streaming bf16 GEMM on tcgen05 / TMEM / TMA for sm_100a.
//
// Role on the path: every dense contraction of the draft/verify forward (reference
// code/beamSD.py:52,221 -> transformers LlamaForCausalLM: q/k/v/o, gate/up/down and lm_head
// projections).  The token count T of a forward is small (10..~300: the K x gamma
// beam tree, plus the prompt in the first round) while the weights are large, so the kernel is laid
// out "swap-AB": the WEIGHT tile is the 128-row MMA A operand (M = 128 output features), the
// activations are the B operand (N = all T tokens, up to 2 x 256 accumulator columns in TMEM), and
// each CTA streams its slice of the weight matrix through a TMA/mbarrier ring exactly once.
//
//   out[s][t][colbase_i + n] = sum_{k in split s} X[t][k] * W_i[n][k]        (fp32 partial sums)
//
// Up to three weight matrices that share the same input (q|k|v, gate|up) are handled by one launch so
// small projections still fill the machine; split-K (grid.y) does the same for the [hidden x K]
// projections.  Partial sums are written as fp32 slices and reduced IN A FIXED ORDER by the consumer
// kernel (elementwise.cu), so results are run-to-run deterministic.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld 32 lanes each -> coalesced fp32 stores along the feature axis).
#include "common.cuh"
#include "kernels.h"

namespace atspeed {

static constexpr int BLOCK_M = 128;    // output features per CTA (UMMA M)
static constexpr int BLOCK_K = 64;     // 64 bf16 = 128 B = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;
static constexpr int GEMM_THREADS = 192;
static constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
static constexpr int MAX_STAGES = 12;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4, [16,30) LBO>>4 (=1, unused for swizzled K-major), [32,46) SBO>>4 (=1024 B between
//   8-row groups), [46,48) version=1, [61,64) layout=2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((BLOCK_M >> 4) << 24);
}

struct GemmParams {
    float* out;
    long long split_stride;   // elements between split slices
    int ldo;                  // row stride (elements) of out
    int T, T_pad;             // tokens, padded to 16
    int K;
    int n_kblocks;            // ceil(K / 64)
    int n_rows[3];            // rows (features) of each weight
    int tiles[3];             // 128-row tiles of each weight
    int colbase[3];           // output column of each weight's row 0
    int stages;
    int tmem_cols;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wx_tcgen05(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmX1, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_tile_bytes = p.T_pad * BLOCK_K * 2;
    const int stage_bytes = A_TILE_BYTES + b_tile_bytes;

    // which weight / tile
    int tile = blockIdx.x, wid = 0;
    if (tile >= p.tiles[0]) { tile -= p.tiles[0]; wid = 1; }
    if (wid == 1 && tile >= p.tiles[1]) { tile -= p.tiles[1]; wid = 2; }
    const CUtensorMap* tmW = wid == 0 ? &tmW0 : (wid == 1 ? &tmW1 : &tmW2);
    const int m0 = tile * BLOCK_M;
    // balanced split-K partition: every split gets >= 1 k-block (host guarantees splits <= n_kblocks)
    const int kb_begin = static_cast<int>((static_cast<long long>(p.n_kblocks) * blockIdx.y) / gridDim.y);
    const int kb_end = static_cast<int>((static_cast<long long>(p.n_kblocks) * (blockIdx.y + 1)) / gridDim.y);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(tmW); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmX1); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(static_cast<uint32_t>(p.tmem_cols)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a_dst = smem + static_cast<size_t>(s) * stage_bytes;
                uint8_t* b_dst = a_dst + A_TILE_BYTES;
                mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(stage_bytes));
                tma_load_2d(tmW, &full_bar[s], a_dst, kb * BLOCK_K, m0);
                tma_load_2d(&tmX, &full_bar[s], b_dst, kb * BLOCK_K, 0);
                if (p.T_pad > 256) tma_load_2d(&tmX1, &full_bar[s], b_dst + 256 * BLOCK_K * 2, kb * BLOCK_K, 256);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t n0 = p.T_pad > 256 ? 256 : p.T_pad;
            const uint32_t n1 = p.T_pad > 256 ? p.T_pad - 256 : 0;
            const uint32_t idesc0 = make_idesc(n0), idesc1 = make_idesc(n1 ? n1 : 16);
            int s = 0; uint32_t ph = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
                const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint32_t acc = (kb > kb_begin || k > 0) ? 1u : 0u;
                    const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2);
                    umma_bf16(tmem_base, da, make_smem_desc(b_addr + k * UMMA_K * 2), idesc0, acc);
                    if (n1) umma_bf16(tmem_base + 256, da, make_smem_desc(b_addr + 256 * BLOCK_K * 2 + k * UMMA_K * 2),
                                      idesc1, acc);
                }
                umma_commit(&empty_bar[s]);   // frees the smem slot when these MMAs retire
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            umma_commit(&accum_bar);          // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global fp32 =====
        mbar_wait(&accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;           // output feature
        const bool row_ok = row < p.n_rows[wid];
        float* out = p.out + static_cast<long long>(blockIdx.y) * p.split_stride + p.colbase[wid] + row;
        for (int c = 0; c < p.T_pad; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c + j < p.T) out[static_cast<long long>(c + j) * p.ldo] = __uint_as_float(r[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(static_cast<uint32_t>(p.tmem_cols)));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* base, long long rows, long long cols, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return ATS_ERR_CUDA; }
    ATS_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
    ATS_CHECK_ARG(cols % 8 == 0, "TMA inner dimension %lld must be a multiple of 8 bf16", cols);
    ATS_CHECK_ARG(box_rows >= 1 && box_rows <= 256, "TMA box rows %d out of range", box_rows);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", static_cast<int>(r)); return ATS_ERR_CUDA; }
    return ATS_OK;
}

int gemm_plan_splits(int total_tiles, int n_kblocks, int num_sms) {
    // Pick the split-K factor with the cheapest estimated schedule: waves x (k-blocks per CTA + a fixed per-CTA cost of
    // ~6 k-block times for prologue/epilogue), measured on B200 with tools/gemm_bench.py.  Splitting multiplies the fp32
    // partial-sum traffic, so wide outputs (many tiles) are only split when a wave would otherwise be mostly empty.
    const int max_s = total_tiles >= num_sms ? 1 : 8;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= max_s && s <= n_kblocks; ++s) {
        const int ctas = total_tiles * s;
        const int waves = (ctas + num_sms - 1) / num_sms;
        const double cost = waves * ((n_kblocks + s - 1) / s + 6.0) + 0.5 * (s - 1);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    return best;
}

// out[s][t][...] for s < splits. X: [T, K] bf16 row-major. W_i: [n_rows_i, K] bf16 row-major.
int gemm_wx(const GemmWeights& w, const void* x, int T, float* out, int ldo, long long split_stride, int splits,
            cudaStream_t stream) {
    ATS_CHECK_ARG(T >= 1 && T <= 512, "gemm: T=%d out of range [1,512]", T);
    ATS_CHECK_ARG(w.n >= 1 && w.n <= 3, "gemm: %d weight matrices", w.n);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.ldo = ldo; p.split_stride = split_stride;
    p.T = T; p.T_pad = (T + 15) & ~15; p.K = w.K;
    p.n_kblocks = (w.K + BLOCK_K - 1) / BLOCK_K;
    ATS_CHECK_ARG(splits >= 1 && splits <= p.n_kblocks, "gemm: splits=%d vs %d k-blocks", splits, p.n_kblocks);
    int total_tiles = 0;
    for (int i = 0; i < 3; ++i) {
        p.n_rows[i] = i < w.n ? w.rows[i] : 0;
        p.tiles[i] = i < w.n ? (w.rows[i] + BLOCK_M - 1) / BLOCK_M : 0;
        p.colbase[i] = i < w.n ? w.colbase[i] : 0;
        total_tiles += p.tiles[i];
    }
    const int stage_bytes = A_TILE_BYTES + p.T_pad * BLOCK_K * 2;
    int stages = (220 * 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > p.n_kblocks) stages = p.n_kblocks < 2 ? 2 : p.n_kblocks;
    ATS_CHECK_ARG(stages >= 2, "gemm: T=%d leaves room for %d pipeline stages", T, stages);
    p.stages = stages;
    p.tmem_cols = p.T_pad <= 32 ? 32 : p.T_pad <= 64 ? 64 : p.T_pad <= 128 ? 128 : p.T_pad <= 256 ? 256 : 512;
    // activations: tokens 0..255 through tmX (box = min(T_pad,256) rows), tokens 256..T_pad-1 through tmX1
    // (box = T_pad-256 rows); rows past T are zero-filled by TMA, and each box always delivers its full byte count.
    CUtensorMap tmX, tmX1;
    ATS_TRY(make_tmap_bf16_kmajor(&tmX, x, T, w.K, p.T_pad > 256 ? 256 : p.T_pad));
    ATS_TRY(make_tmap_bf16_kmajor(&tmX1, x, T, w.K, p.T_pad > 256 ? p.T_pad - 256 : 16));
    const size_t smem_bytes = static_cast<size_t>(p.stages) * stage_bytes + 1024;
    static int max_dyn = 0;
    if (!max_dyn) {
        cudaFuncAttributes fa;
        ATS_CUDA(cudaFuncGetAttributes(&fa, gemm_wx_tcgen05));
        const int want = 227 * 1024 - static_cast<int>(fa.sharedSizeBytes);   // static barriers share the 227 KiB budget
        ATS_CUDA(cudaFuncSetAttribute(gemm_wx_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
        max_dyn = want;
    }
    ATS_CHECK_ARG(static_cast<int>(smem_bytes) <= max_dyn, "gemm: %zu bytes of shared memory > %d", smem_bytes, max_dyn);
    dim3 grid(total_tiles, splits);
    gemm_wx_tcgen05<<<grid, GEMM_THREADS, smem_bytes, stream>>>(w.tmap[0], w.tmap[w.n > 1 ? 1 : 0],
                                                                  w.tmap[w.n > 2 ? 2 : 0], tmX, tmX1, p);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed
