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
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace atspeed {

static constexpr int BLOCK_M = 128;    // output features per CTA (UMMA M)
static constexpr int BLOCK_K = 64;     // 64 bf16 = 128 B = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;
static constexpr int GEMM_THREADS = 192;
static constexpr size_t EXCLUSIVE_SMEM_BYTES = 120 * 1024;   // > 227 KiB / 2: at most one GEMM CTA per SM
static constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
static constexpr int MAX_STAGES = 12;
static constexpr int OUT_STAGE_BYTES = 2 * 16 * 128 * 4;    // fused epilogues only: two [16 tokens][128 features] fp32 staging tiles

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
// every wait of the pipelines is bounded (common.cuh): a lost arrival traps with a HangDiag record instead of spinning
struct WaitCtx {
    SpinGuard g;
    unsigned kernel, role, u_begin, u_end, T;
};
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, const WaitCtx& w, unsigned barrier, unsigned index,
                                          unsigned unit) {
    mbar_wait_guarded(smem_u32(bar), parity, w.g, w.kernel, w.role, barrier, index, unit, w.u_begin, w.u_end, w.T);
}
// L2 eviction-priority hints of the TMA loads (the encoded createpolicy values CUTLASS uses, cute/arch/copy_sm90_desc.hpp).
// A weight tile is read by exactly one CTA (pair) once per launch and a launch streams 100-270 MB of them through the 126 MB
// L2: evict-first, or they push out what IS reused -- the activation tile every CTA re-reads for every k-block (evict-last)
// and the fp32 partial sums the consumer kernel is about to read (ncu, session I: the row-wise consumers hit L2 for 0.1-10 %
// of their reads and ran at 2-3 TB/s of DRAM bandwidth on data the GEMM had written microseconds earlier).
static constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
static constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
// ---- cta_group::2 (CTA pair) forms ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs of the pair; the transaction bytes are credited to the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
// the same, delivered to every CTA of `cta_mask` (same shared-memory offset in each); each destination's bytes are credited to
// the full barrier of ITS pair leader (peer bit cleared).  CUTLASS: SM100_TMA_2SM_LOAD_MULTICAST.
__device__ __forceinline__ void tma_load_2d_2sm_mc(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, uint16_t cta_mask,
                                                   uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5, %6;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(cta_mask), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once the issuing thread's MMAs retire) on the barrier at the same shared-memory offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// arrive on the barrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_m(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
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

// Drain one segment's accumulators (one 128-row half: this CTA's 128 TMEM lanes x n_chunks x 16 columns) into its fp32
// partial-sum slice.  Every thread stores its feature's 16 tokens of a chunk straight from registers: a warp's store instruction
// writes 32 consecutive features of one token, one full 128-byte line.  The tcgen05.ld of the next chunk is in flight while this
// one is stored (two register sets), and `release` (the next segment's MMAs may overwrite TMEM) runs as soon as the last load has
// completed.  The epilogue of a T > 256 segment is exposed -- one accumulator fills TMEM -- so its length is kernel time.
// (Round 2 also had a shared-memory path: 16 x 128 staging tiles + cp.async.bulk.tensor stores, ring of three, one barrier per
// chunk.  Same-box A/B, profiles/r02_gemm_epilogue_ab.txt: the register path is 4-10 % faster per launch at every T >= 50 --
// four independent warps beat a barrier, a proxy fence and an elected thread per 8 KB -- so the staged path was removed.)
template <typename ReleaseFn>
__device__ __forceinline__ void drain_segment_stg(float* out_f, long long ldo, int T, bool row_ok, uint32_t tb0, int n_chunks,
                                                  ReleaseFn release) {
    auto emit = [&](const uint32_t (&r)[16], int idx) {
        if (!row_ok) return;
        float* o = out_f + static_cast<long long>(idx << 4) * ldo;
        if ((idx << 4) + 16 <= T) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[static_cast<long long>(j) * ldo] = __uint_as_float(r[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if ((idx << 4) + j < T) o[static_cast<long long>(j) * ldo] = __uint_as_float(r[j]);
        }
    };
    uint32_t ra[16], rb[16];
    tmem_ld16(tb0, ra);
    for (int idx = 0; idx < n_chunks; idx += 2) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (idx + 1 < n_chunks) tmem_ld16(tb0 + static_cast<uint32_t>((idx + 1) << 4), rb);
        else release();
        emit(ra, idx);
        if (idx + 1 < n_chunks) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (idx + 2 < n_chunks) tmem_ld16(tb0 + static_cast<uint32_t>((idx + 2) << 4), ra);
            else release();
            emit(rb, idx + 1);
        }
    }
}

struct GemmParams {
    float* out;
    long long slice_stride;   // elements between partial-sum slices
    int ldo;                  // row stride (elements) of out
    int T, T_pad;             // tokens, padded to 16
    int n_kblocks;            // ceil(K / 64)
    int n_rows[3];            // rows (features) of each weight
    int tiles[3];             // BM-row tiles of each weight
    int colbase[3];           // output column of each weight's row 0
    int BM;                   // 128, or 256 = two stacked 128-row MMAs sharing one activation tile
    int U;                    // work units (tile, k-block) per CTA; CTA c owns units [c*U, (c+1)*U)
    int total_units;
    int stages;
    int tmem_cols, acc_stride;
    int n_bufs, buf_stride;   // accumulator double buffering in TMEM: segment s uses buffer s % n_bufs
    int b_box_bytes;          // bytes the activation TMA box(es) deliver per stage (== T_pad * 128 except in timing experiments)
    int n_mma, N_mma;         // 2-CTA kernel: the token axis (padded to 64) is covered by n_mma MMAs of N_mma columns each
    SpinGuard guard;          // bound + diagnostic record of every mbarrier wait
    GemmTrace trace;          // optional per-CTA progress words (ATSPEED_GEMM_TRACE=1)
    FusedEpi epi;             // kind != EPI_SLICES: tiles are finished inside the kernel (kernels.h)
    int l2_hints;             // TMA loads carry L2 eviction priorities (ATSPEED_GEMM_L2HINT=0: plain loads)
};
__device__ __forceinline__ void trace_put(const GemmTrace& t, int word, unsigned v) {
    if (t.buf != nullptr && blockIdx.x < TRACE_CTAS) {
        volatile unsigned int* w = t.buf + blockIdx.x * TRACE_WORDS + word;
        *w = v;
    }
}
__device__ __forceinline__ unsigned sm_id() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

// ---------------------------------------------------------------------------------------------
// Fused epilogues (FusedEpi, kernels.h): shared by the single-CTA and the CTA-pair kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void flag_release(unsigned int* flag, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int flag_acquire(const unsigned int* flag) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
// bounded like every other wait of the kernel: the CTA waited for is one with a HIGHER index that writes its partial as the
// first thing it does, so this never waits long unless that CTA is not resident yet (it is dispatched in index order)
__device__ __forceinline__ void flag_wait(const unsigned int* flag, unsigned int epoch, const WaitCtx& w, unsigned worker, unsigned unit) {
    if (flag_acquire(flag) == epoch) return;
    const unsigned long long t0 = global_timer_ns();
    for (uint32_t spins = 1;; ++spins) {
        if (flag_acquire(flag) == epoch) return;
        if ((spins & 63u) == 0u && w.g.limit_ns != 0ull) {
            const unsigned long long dt = global_timer_ns() - t0;
            if (dt > w.g.limit_ns) hang_report(w.g, w.kernel, w.role, HANG_B_FLAG, worker, epoch, unit, w.u_begin, w.u_end, w.T, dt);
        }
        __nanosleep(64);
    }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 r;
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
    return r;
}
// Drain every segment of this CTA's unit range [u_begin, u_end).  `worker` = owner of the unit range (the CTA, or the CTA pair),
// `sub` = this CTA's rank inside the worker.  One 128-row half of a tile sits in TMEM lanes 0..127 (lane = feature row),
// token t in accumulator column t.  Called by the four epilogue warps (threads 64..191).
//   non-owner segment (does not start at the tile's first k-block): raw fp32 partial -> FusedEpi::part, then the flag;
//   owner segment: wait for the flags of the CTAs that hold the rest of the tile, add their partials in slice order, then
//     EPI_QKV_ROPE : stage [16 tokens][128 features] in shared memory (a RoPE pair sits head_dim/2 lanes apart), rotate,
//                    write q to qbuf and k / v to the KV-cache rows with 16-byte stores;
//     EPI_SILU_MUL : interleaved tile (lanes 0..63 gate, 64..127 up of the same 64 features) through the same staging tile,
//                    or stacked tile (BM = 256: gate in the first accumulator, up in the second, same lane) from registers.
template <bool PAIR>
__device__ __forceinline__ void fused_epilogue(const GemmParams& p, int worker, int sub, int u_begin, int u_end, uint32_t tmem_base,
                                               float* stage_out, uint64_t* accum_full, uint64_t* accum_empty, const WaitCtx& wc,
                                               int* s_pos, int* s_slotuser, long long* s_kvoff) {
    const FusedEpi& e = p.epi;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3;
    const int f = q * 32 + lane;                    // this thread's TMEM lane = feature row of the 128-row half
    const int ti = threadIdx.x - 64;                // 0..127 among the epilogue threads
    const bool elected = ti == 0;
    const int KB = p.n_kblocks;
    constexpr int NSUB = PAIR ? 2 : 1;
    const int halves = PAIR ? 1 : (p.BM >> 7);
    const int BMW = PAIR ? 256 : p.BM;              // weight rows per tile of a worker
    const bool stacked = !PAIR && e.kind == EPI_SILU_MUL && p.BM == 256;
    const long long half_elems = static_cast<long long>(p.T_pad) * 128;
    float* my_part = e.part + static_cast<long long>(worker * NSUB + sub) * halves * half_elems;
    const int tok = ti >> 3, g8 = (ti & 7) * 8;      // staged chunk: this thread's token and first of 8 consecutive features
    const int hd = e.head_dim, hhalf = e.head_dim >> 1;
    const int nchunk = (p.T + 15) >> 4;
    // RoPE: the 8 features [g8, g8 + 8) and [g8 + 64, g8 + 72) of a staged tile use the SAME 8 cos / sin values (their indices
    // inside the head differ by 64: a multiple of head_dim, or exactly head_dim / 2 for heads of 128), so one set per token
    // serves both groups; whether a group is the low or the high half of its rotary pairs is decided per group below
    const int ci = (g8 % hd) % hhalf;
    if (e.kind == EPI_QKV_ROPE) {
        // per-token metadata once per CTA (positions, cache rows): the chunk loop below must not wait for global memory
        for (int i = ti; i < p.T; i += 128) {
            int pp = e.pos[i];
            s_pos[i] = pp < 0 ? 0 : (pp >= e.max_pos ? e.max_pos - 1 : pp);
            s_slotuser[i] = e.slot[i] | ((e.tok_user ? e.tok_user[i] : 0) << 24);
        }
        if (ti < MAX_USERS) s_kvoff[ti] = e.tok_user ? e.kv_off[ti] : 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    int seg = 0, chunk = 0;
    for (int u = u_begin; u < u_end; ++seg) {
        const int tile = u / KB;
        const int seg_end = min((tile + 1) * KB, u_end);
        const bool owner = u == tile * KB;
        const int buf = p.n_bufs == 2 ? (seg & 1) : 0, use = p.n_bufs == 2 ? (seg >> 1) : seg;
        mbar_wait(&accum_full[buf], use & 1, wc, HANG_B_ACCUM_FULL, buf, u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tb = tmem_base + buf * p.buf_stride + (static_cast<uint32_t>(q * 32) << 16);
        if (!owner) {
            for (int half = 0; half < halves; ++half) {
                float* dst = my_part + half * half_elems + f;
                for (int c = 0; c < p.T; c += 16) {
                    uint32_t r[16];
                    tmem_ld16(tb + half * p.acc_stride + static_cast<uint32_t>(c), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[static_cast<long long>(c + j) * 128] = __uint_as_float(r[j]);   // rows < T_pad
                }
            }
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (elected) flag_release(e.flags + worker * NSUB + sub, e.epoch);
        } else {
            const int n_extra = ((tile + 1) * KB - 1) / p.U - worker;     // workers after this one that hold a part of the tile
            if (n_extra > 0) {
                if (elected)
                    for (int s2 = 1; s2 <= n_extra; ++s2) flag_wait(e.flags + (worker + s2) * NSUB + sub, e.epoch, wc, worker + s2, u);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            // raw partial of slice s2 (the worker s2 places after this one), half `half`, this thread's feature row
            auto part_of = [&](int s2, int half) -> const float* {
                return e.part + (static_cast<long long>((worker + s2) * NSUB + sub) * halves + half) * half_elems + f;
            };
            int wid = 0, tw = tile;                   // weight and tile index inside it (EPI_QKV_ROPE)
            if (e.kind == EPI_QKV_ROPE) {
                if (tw >= p.tiles[0]) { tw -= p.tiles[0]; wid = 1; }
                if (wid == 1 && tw >= p.tiles[1]) { tw -= p.tiles[1]; wid = 2; }
            }
            if (stacked) {
                // gate = first accumulator, up = second, same lane: straight from registers.  The partials of the next chunk
                // are requested before this chunk is processed (an L2 round trip per chunk would otherwise be exposed).
                const int feat = tile * 128 + f;
                float png[16], pnu[16];
                if (n_extra > 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { png[j] = __ldcg(part_of(1, 0) + static_cast<long long>(j) * 128); pnu[j] = __ldcg(part_of(1, 1) + static_cast<long long>(j) * 128); }
                }
                for (int c = 0; c < p.T; c += 16) {
                    uint32_t rg[16], ru[16];
                    tmem_ld16(tb + static_cast<uint32_t>(c), rg);
                    tmem_ld16(tb + p.acc_stride + static_cast<uint32_t>(c), ru);
                    float g[16], uu[16];
                    if (n_extra > 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { g[j] = png[j]; uu[j] = pnu[j]; }
                        if (c + 16 < p.T) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                png[j] = __ldcg(part_of(1, 0) + static_cast<long long>(c + 16 + j) * 128);
                                pnu[j] = __ldcg(part_of(1, 1) + static_cast<long long>(c + 16 + j) * 128);
                            }
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        g[j] = n_extra > 0 ? __uint_as_float(rg[j]) + g[j] : __uint_as_float(rg[j]);
                        uu[j] = n_extra > 0 ? __uint_as_float(ru[j]) + uu[j] : __uint_as_float(ru[j]);
                    }
                    for (int s2 = 2; s2 <= n_extra; ++s2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            g[j] += __ldcg(part_of(s2, 0) + static_cast<long long>(c + j) * 128);
                            uu[j] += __ldcg(part_of(s2, 1) + static_cast<long long>(c + j) * 128);
                        }
                    }
                    if (feat < e.mlp) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c + j < p.T) e.m[static_cast<long long>(c + j) * e.mlp + feat] = __float2bfloat16_rn(silu_mul_bf16(g[j], uu[j]));
                    }
                }
            } else {
                const int total = halves * nchunk;
                float pn[16];                                          // slice-1 partial of the NEXT chunk (prefetched)
                float4 cnx[2], snx[2];                                 // cos / sin of this thread's token in the NEXT chunk
                auto fetch_partial = [&](int idx) {
                    const float* src = part_of(1, idx / nchunk) + static_cast<long long>((idx % nchunk) * 16) * 128;
#pragma unroll
                    for (int j = 0; j < 16; ++j) pn[j] = __ldcg(src + static_cast<long long>(j) * 128);
                };
                auto fetch_rope = [&](int idx) {
                    const int t = (idx % nchunk) * 16 + tok;
                    if (e.kind == EPI_QKV_ROPE && wid != 2 && t < p.T) {
                        const float* ct = e.rope_cos + static_cast<long long>(s_pos[t]) * hhalf + ci;
                        const float* sn = e.rope_sin + static_cast<long long>(s_pos[t]) * hhalf + ci;
                        cnx[0] = __ldg(reinterpret_cast<const float4*>(ct)); cnx[1] = __ldg(reinterpret_cast<const float4*>(ct) + 1);
                        snx[0] = __ldg(reinterpret_cast<const float4*>(sn)); snx[1] = __ldg(reinterpret_cast<const float4*>(sn) + 1);
                    }
                };
                if (n_extra > 0) fetch_partial(0);
                fetch_rope(0);
                for (int idx = 0; idx < total; ++idx, ++chunk) {
                    const int half = idx / nchunk, c = (idx - half * nchunk) * 16;
                    uint32_t r[16];
                    tmem_ld16(tb + half * p.acc_stride + static_cast<uint32_t>(c), r);
                    float v[16];
                    const float4 cc0 = cnx[0], cc1 = cnx[1], ss0 = snx[0], ss1 = snx[1];
                    if (n_extra > 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = pn[j];
                        if (idx + 1 < total) fetch_partial(idx + 1);
                    }
                    if (idx + 1 < total) fetch_rope(idx + 1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = n_extra > 0 ? __uint_as_float(r[j]) + v[j] : __uint_as_float(r[j]);
                    for (int s2 = 2; s2 <= n_extra; ++s2) {
                        const float* src = part_of(s2, half) + static_cast<long long>(c) * 128;
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += __ldcg(src + static_cast<long long>(j) * 128);
                    }
                    // double-buffered staging tile: chunk k+2 overwrites buffer k only after every thread has passed the
                    // barrier of chunk k+1, i.e. after it finished reading buffer k
                    float* stg = stage_out + (chunk & 1) * (16 * 128);
#pragma unroll
                    for (int j = 0; j < 16; ++j) stg[j * 128 + f] = v[j];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int t = c + tok;
                    if (t >= p.T) continue;
                    const float* srow = stg + tok * 128;
                    if (e.kind == EPI_SILU_MUL) {
                        // lanes 0..63 = gate rows, 64..127 = up rows of features [f0, f0 + 64)
                        const int feat = tile * (PAIR ? 128 : 64) + sub * 64 + g8;
                        if (feat < e.mlp) {
                            const float4 ga = *reinterpret_cast<const float4*>(srow + g8), gb = *reinterpret_cast<const float4*>(srow + g8 + 4);
                            const float4 ua = *reinterpret_cast<const float4*>(srow + 64 + g8), ub = *reinterpret_cast<const float4*>(srow + 64 + g8 + 4);
                            float o[8] = {silu_mul_bf16(ga.x, ua.x), silu_mul_bf16(ga.y, ua.y), silu_mul_bf16(ga.z, ua.z), silu_mul_bf16(ga.w, ua.w),
                                          silu_mul_bf16(gb.x, ub.x), silu_mul_bf16(gb.y, ub.y), silu_mul_bf16(gb.z, ub.z), silu_mul_bf16(gb.w, ub.w)};
                            *reinterpret_cast<uint4*>(e.m + static_cast<long long>(t) * e.mlp + feat) = pack_bf16x8(o);
                        }
                    } else {
                        const int su = s_slotuser[t];
                        const long long srow_kv = static_cast<long long>(su & 0xffffff) * e.HD + s_kvoff[(su >> 24) & (MAX_USERS - 1)];
                        const float cs[8] = {cc0.x, cc0.y, cc0.z, cc0.w, cc1.x, cc1.y, cc1.z, cc1.w};
                        const float sn[8] = {ss0.x, ss0.y, ss0.z, ss0.w, ss1.x, ss1.y, ss1.z, ss1.w};
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            const int grp = g8 + gi * 64;                         // 8 features [grp, grp + 8) of the staged 128
                            const int col = tw * BMW + sub * 128 + half * 128 + grp;   // column inside the weight's output
                            if (col >= p.n_rows[wid]) continue;
                            float o[8];
                            const float4 xa = *reinterpret_cast<const float4*>(srow + grp), xb = *reinterpret_cast<const float4*>(srow + grp + 4);
                            const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                            if (wid == 2) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) o[k] = x[k];
                            } else {
                                // HF apply_rotary_pos_emb in bf16: (x * cos) + (rotate_half(x) * sin), every op rounded
                                const bool rope_lo = (grp % hd) < hhalf;
                                const int pgrp = rope_lo ? grp + hhalf : grp - hhalf;    // the RoPE partners, same staged tile
                                const float4 ya = *reinterpret_cast<const float4*>(srow + pgrp), yb = *reinterpret_cast<const float4*>(srow + pgrp + 4);
                                const float y[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const float xx = bf16_round(x[k]), yy = bf16_round(y[k]);
                                    o[k] = rope_lo ? bf16_round(bf16_round(xx * cs[k]) + bf16_round(-yy * sn[k]))
                                                   : bf16_round(bf16_round(xx * cs[k]) + bf16_round(yy * sn[k]));
                                }
                            }
                            __nv_bfloat16* dst = wid == 0 ? e.qbuf + static_cast<long long>(t) * e.HD + col
                                               : (wid == 1 ? e.kcache : e.vcache) + srow_kv + col;
                            *reinterpret_cast<uint4*>(dst) = pack_bf16x8(o);
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(&accum_empty[buf], 0);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty[buf])) : "memory");
        }
        u = seg_end;
    }
}

// Work decomposition ("stream-K with consumer-side fix-up").  The (tile, k-block) units of the whole GEMM are
// numbered tile-major and cut into gridDim.x equal contiguous ranges, one per persistent CTA, so every SM streams the
// same number of weight bytes whatever the tile count (no partial last wave).  A tile whose k-range is cut across CTAs
// gets one fp32 partial-sum slice per CTA: slice index = cta - first_cta_of_tile.  The consumer kernels
// (elementwise.cu) know the same arithmetic (SplitMap) and add a tile's slices in index order, so the result is
// deterministic.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wx_tcgen05(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmX1, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t accum_full[2], accum_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) int s_pos[512], s_slotuser[512];     // fused RoPE epilogue: per-token position / cache row
    __shared__ __align__(8) long long s_kvoff[MAX_USERS];

    // let the next kernel of the stream start its own prologue / weight prefetch as early as resources allow (PDL)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_bytes = p.BM * BLOCK_K * 2;
    const int b_tile_bytes = p.T_pad * BLOCK_K * 2;
    const int stage_bytes = a_bytes + b_tile_bytes;
    const int KB = p.n_kblocks;
    const int u_begin = blockIdx.x * p.U;
    const int u_end = min(u_begin + p.U, p.total_units);
    const int n_units = u_end - u_begin;          // host guarantees >= 1

    auto tile_of = [&](int tile, int& wid, int& m0) {
        wid = 0;
        if (tile >= p.tiles[0]) { tile -= p.tiles[0]; wid = 1; }
        if (wid == 1 && tile >= p.tiles[1]) { tile -= p.tiles[1]; wid = 2; }
        m0 = tile * p.BM;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&accum_full[b], 1);
            mbar_init(&accum_empty[b], 4);        // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmX1);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(static_cast<uint32_t>(p.tmem_cols)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    WaitCtx wc;
    wc.g = p.guard; wc.kernel = HANG_K_GEMM; wc.u_begin = u_begin; wc.u_end = u_end; wc.T = p.T;
    wc.role = warp == 0 ? HANG_R_PRODUCER : (warp == 1 ? HANG_R_MMA : HANG_R_EPILOGUE);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            // (tile, kb) of the next unit to load, advanced incrementally: the loops of the two single-thread roles are on
            // the kernel's critical path, so they contain no integer division
            int tile = u_begin / KB, kb = u_begin - tile * KB, wid, m0;
            tile_of(tile, wid, m0);
            const uint64_t hint_w = p.l2_hints ? L2_EVICT_FIRST : 0x1000000000000000ull;      // weights: streamed once
            const uint64_t hint_x = p.l2_hints ? L2_EVICT_LAST : 0x1000000000000000ull;       // activations: re-read by all
            auto advance = [&]() {
                if (++kb == KB) { kb = 0; ++tile; tile_of(tile, wid, m0); }
            };
            auto load_a = [&](int s) {
                const CUtensorMap* tmW = wid == 0 ? &tmW0 : (wid == 1 ? &tmW1 : &tmW2);
                uint8_t* a_dst = smem + static_cast<size_t>(s) * stage_bytes;
                mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(a_bytes + p.b_box_bytes));
                if (p.epi.kind == EPI_SILU_MUL) {
                    // interleaved tile: the gate rows and the up rows of the SAME features (two boxes of BM/2 rows)
                    const int rows = p.BM >> 1;
                    tma_load_2d(&tmW0, &full_bar[s], a_dst, kb * BLOCK_K, tile * rows, hint_w);
                    tma_load_2d(&tmW1, &full_bar[s], a_dst + rows * BLOCK_K * 2, kb * BLOCK_K, tile * rows, hint_w);
                } else {
                    tma_load_2d(tmW, &full_bar[s], a_dst, kb * BLOCK_K, m0, hint_w);      // BM = 256: the map's box is 256 rows
                }
            };
            auto load_b = [&](int s, int kbb) {
                uint8_t* b_dst = smem + static_cast<size_t>(s) * stage_bytes + a_bytes;
                tma_load_2d(&tmX, &full_bar[s], b_dst, kbb * BLOCK_K, 0, hint_x);
                if (p.T_pad > 256) tma_load_2d(&tmX1, &full_bar[s], b_dst + 256 * BLOCK_K * 2, kbb * BLOCK_K, 256, hint_x);
            };
            // The weights do not depend on the previous kernel: fill the ring with weight tiles first, then wait for
            // the producer of the activations (griddepcontrol.wait), then add the activation tiles.
            const int npre = n_units < p.stages ? n_units : p.stages;
            const int kb_first = kb;
            for (int i = 0; i < npre; ++i) { load_a(i); advance(); }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int i = 0, kbb = kb_first; i < npre; ++i) { load_b(i, kbb); if (++kbb == KB) kbb = 0; }
            int s = npre == p.stages ? 0 : npre;
            uint32_t ph = npre == p.stages ? 1 : 0;
            for (int u = u_begin + npre; u < u_end; ++u) {
                mbar_wait(&empty_bar[s], ph ^ 1, wc, HANG_B_EMPTY, s, u);
                load_a(s);
                load_b(s, kb);
                advance();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t n0 = p.T_pad > 256 ? 256 : p.T_pad;
            const uint32_t n1 = p.T_pad > 256 ? p.T_pad - 256 : 0;
            const uint32_t idesc0 = make_idesc(n0), idesc1 = make_idesc(n1 ? n1 : 16);
            int s = 0; uint32_t ph = 0;
            int kb = u_begin % KB;
            int buf = 0;                       // accumulator buffer of the current segment (alternates when double-buffered)
            uint32_t use0 = 0, use1 = 0;       // how many segments each buffer has held so far (scalars: no local memory)
            for (int u = u_begin; u < u_end; ++u) {
                const bool seg_start = (u == u_begin) || kb == 0;
                const uint32_t used = buf ? use1 : use0;
                if (seg_start && used > 0) {
                    // the epilogue must have drained this buffer's previous segment before it is overwritten
                    mbar_wait(&accum_empty[buf], (used - 1) & 1, wc, HANG_B_ACCUM_EMPTY, buf, u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t tacc = tmem_base + buf * p.buf_stride;
                mbar_wait(&full_bar[s], ph, wc, HANG_B_FULL, s, u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint32_t acc = (!seg_start || k > 0) ? 1u : 0u;
                    const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2);
                    const uint64_t db = make_smem_desc(b_addr + k * UMMA_K * 2);
                    umma_bf16(tacc, da, db, idesc0, acc);
                    if (p.BM == 256)
                        umma_bf16(tacc + p.acc_stride, make_smem_desc(a_addr + A_TILE_BYTES + k * UMMA_K * 2), db, idesc0, acc);
                    if (n1) umma_bf16(tacc + 256, da, make_smem_desc(b_addr + 256 * BLOCK_K * 2 + k * UMMA_K * 2),
                                      idesc1, acc);
                }
                umma_commit(&empty_bar[s]);   // frees the smem slot when these MMAs retire
                if (++s == p.stages) { s = 0; ph ^= 1; }
                if (++kb == KB) kb = 0;
                if (u + 1 == u_end || kb == 0) {             // segment complete
                    umma_commit(&accum_full[buf]);
                    if (buf) ++use1; else ++use0;
                    if (p.n_bufs == 2) buf ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> global fp32 partial-sum slice =====
        asm volatile("griddepcontrol.wait;" ::: "memory");    // `out` may still be read by the previous consumer
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        float* stage_out = reinterpret_cast<float*>(smem + static_cast<size_t>(p.stages) * stage_bytes);   // fused epilogues only
        int seg = 0;
        if (p.epi.kind != EPI_SLICES) {
            fused_epilogue<false>(p, blockIdx.x, 0, u_begin, u_end, tmem_base, stage_out, accum_full, accum_empty, wc, s_pos, s_slotuser, s_kvoff);
        } else
        for (int u = u_begin; u < u_end; ++seg) {
            const int tile = u / KB;
            const int seg_end = min((tile + 1) * KB, u_end);
            int wid, m0;
            tile_of(tile, wid, m0);
            const int slice = blockIdx.x - (tile * KB) / p.U;
            const int buf = p.n_bufs == 2 ? (seg & 1) : 0, use = p.n_bufs == 2 ? (seg >> 1) : seg;
            mbar_wait(&accum_full[buf], use & 1, wc, HANG_B_ACCUM_FULL, buf, u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const int halves = p.BM >> 7;
                for (int half = 0; half < halves; ++half) {
                    const int row = m0 + half * BLOCK_M + q * 32 + lane;           // output feature
                    float* out = p.out + static_cast<long long>(slice) * p.slice_stride + p.colbase[wid] + row;
                    const uint32_t tbase = tmem_base + buf * p.buf_stride + half * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
                    const bool last_half = half + 1 == halves;
                    drain_segment_stg(out, p.ldo, p.T, row < p.n_rows[wid], tbase, p.T_pad >> 4, [&]() {
                        if (!last_half) return;
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty[buf])) : "memory");
                    });
                }
                u = seg_end;
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
// CTA-pair variant (cta_group::2) for large token counts (cohort forwards, T > 256).
//
// Two CTAs on the SMs of one TPC own one 256-row weight tile: each streams ITS 128 rows (A) and HALF of the activation
// tile (B: the tokens are split across the pair), and the leader's single MMA thread issues M = 256 MMAs that read both
// CTAs' shared memory and write both CTAs' TMEM.  Per CTA a k-block stage is 16 KB + T/2 x 128 B instead of
// 16 KB + T x 128 B, so the ring is twice as deep at T = 512 (4 stages instead of 2) and the L2->SM traffic of the
// activations is halved.  Work decomposition, slices and epilogue are those of gemm_wx_tcgen05 with "CTA" read as
// "pair" and BM = 256.
// Barriers (same offsets in both CTAs): full[s] lives in the leader (its producer posts the expected bytes of BOTH CTAs,
// every TMA of the pair completes on it); empty[s] and accum_full[b] are signalled in both CTAs by the leader's
// multicast commit; accum_empty[b] lives in the leader and collects the 8 epilogue warps of the pair.
// ---------------------------------------------------------------------------------------------
//
// CL = 4: two pairs in one cluster work on ADJACENT tiles (a 512-row "super-tile") over the same k-blocks, in lock step, and
// share the activation stream: each CTA fetches only HALF of its half of the activation tile and multicasts it to the CTA of
// the same pair rank in the other pair.  L2 -> SM requests per CTA and k-block: 16 KB + T/4 x 128 B instead of 16 KB + T/2 x
// 128 B (ncu, T = 512, CL = 2: the launch moves 585 MB out of L2 for 190 MB of weights and runs at ~4800 B/clk of the
// ~6300 B/clk the L2 delivers to TMA; the tensor pipe is busy 63 % of a CTA's life).  A stage may be refilled only when BOTH
// pairs have consumed it (a CTA's multicast writes into the other pair's shared memory): empty[s] collects one commit from
// each leader.  The work unit is (super-tile, k-block) and the worker is the cluster; an odd tile count leaves the last
// super-tile with a phantom tile (zero-filled loads, epilogue skipped).
template <int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wx_tcgen05_2cta(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
                     const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t accum_full[2], accum_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) int s_pos[512], s_slotuser[512];     // fused RoPE epilogue: per-token position / cache row
    __shared__ __align__(8) long long s_kvoff[MAX_USERS];

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        trace_put(p.trace, 0, p.trace.seq);
        trace_put(p.trace, 1, (2u << 24) | (rank << 16) | sm_id());
        trace_put(p.trace, 7, (gridDim.x << 16) | static_cast<unsigned>(p.T));
        trace_put(p.trace, 2, 1);
    }
    const uint32_t prank = rank & 1u;                         // rank inside the pair
    const int cpair = static_cast<int>(rank >> 1);            // pair inside the cluster (0 when CL == 2)
    const bool leader = prank == 0;
    const uint32_t lead_rank = rank & ~1u;
    const int pair = blockIdx.x / CL;                         // the worker: owner of a unit range (a pair, or a cluster of two pairs)
    constexpr int TPS = CL / 2;                               // tiles per (super-)tile of a unit
    const int n_tiles_total = p.tiles[0] + p.tiles[1] + p.tiles[2];
    const int T64 = p.n_mma * p.N_mma;                       // tokens padded for the MMA / TMA boxes (pair_pad_tokens)
    const int b_half_rows = p.N_mma >> 1;                    // tokens of one MMA held by this CTA
    const int b_mma_bytes = b_half_rows * BLOCK_K * 2;
    const int a_bytes = A_TILE_BYTES;
    const int stage_bytes = a_bytes + p.n_mma * b_mma_bytes;
    const int KB = p.n_kblocks;
    const int u_begin = pair * p.U;
    const int u_end = min(u_begin + p.U, p.total_units);
    const int n_units = u_end - u_begin;

    auto tile_of = [&](int tile, int& wid, int& m0) {
        wid = 0;
        if (tile >= p.tiles[0]) { tile -= p.tiles[0]; wid = 1; }
        if (wid == 1 && tile >= p.tiles[1]) { tile -= p.tiles[1]; wid = 2; }
        m0 = tile * 256 + static_cast<int>(prank) * BLOCK_M;  // this CTA's 128 rows of the pair's 256-row tile
    };
    // this pair's tile of super-tile `st` (>= n_tiles_total: the phantom tile behind an odd tile count)
    auto my_tile = [&](int st) { return st * TPS + cpair; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TPS); }   // one commit per leader
        for (int b = 0; b < 2; ++b) {
            mbar_init(&accum_full[b], 1);
            mbar_init(&accum_empty[b], 8);        // 4 epilogue warps of each CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmX);
    }
    // BOTH CTAs of the pair must be running before tcgen05.alloc.cta_group::2: the allocation is a handshake between the
    // two CTAs through mbarriers in their reserved shared memory (the leader allocates in both SMs and posts the address to
    // the peer).  Co-scheduling guarantees that the peer WILL be resident, not that it has started: a leader that runs ahead
    // posts into a CTA that does not exist yet and both then wait forever inside tcgen05.alloc -- the round-1 device stall
    // (no mbarrier of this kernel involved, GPU 98 % busy).  CUTLASS / DeepGEMM open their 2-SM kernels
    // with this same cluster barrier.
    cluster_sync_all();
    if (warp == 1) {
        if (lane == 0) trace_put(p.trace, 3, 1);
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(static_cast<uint32_t>(p.tmem_cols)));
        if (lane == 0) trace_put(p.trace, 3, 2);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        if (lane == 0) trace_put(p.trace, 3, 3);
    }
    if (threadIdx.x == 0) trace_put(p.trace, 2, 2);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                      // barriers of BOTH CTAs initialised, TMEM allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) trace_put(p.trace, 2, 4);

    WaitCtx wc;
    wc.g = p.guard; wc.kernel = HANG_K_GEMM_PAIR; wc.u_begin = u_begin; wc.u_end = u_end; wc.T = p.T;
    wc.role = warp == 0 ? HANG_R_PRODUCER : (warp == 1 ? HANG_R_MMA : HANG_R_EPILOGUE);

    if (warp == 0) {
        // ===== TMA producer (one thread in each CTA) =====
        if (lane == 0) {
            int stile = u_begin / KB, kb = u_begin - stile * KB, wid, m0;
            // a phantom tile loads fully out-of-bounds boxes: zero fill, the byte count still arrives
            auto tile_or_phantom = [&](int st) {
                const int t = my_tile(st);
                if (t < n_tiles_total) tile_of(t, wid, m0);
                else { wid = 0; m0 = p.n_rows[0] + 256; }
            };
            tile_or_phantom(stile);
            const uint64_t hint_w = p.l2_hints ? L2_EVICT_FIRST : 0x1000000000000000ull;      // weights: streamed once
            const uint64_t hint_x = p.l2_hints ? L2_EVICT_LAST : 0x1000000000000000ull;       // activations: re-read by all
            auto advance = [&]() {
                if (++kb == KB) { kb = 0; ++stile; tile_or_phantom(stile); }
            };
            auto load_a = [&](int s) {
                const CUtensorMap* tmW = wid == 0 ? &tmW0 : (wid == 1 ? &tmW1 : &tmW2);
                if (leader) mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(2 * stage_bytes));
                uint8_t* a_dst = smem + static_cast<size_t>(s) * stage_bytes;
                if (p.epi.kind == EPI_SILU_MUL) {
                    // this CTA's 128 rows = the gate rows and the up rows of ITS 64 features of the pair's 128-feature tile
                    const int r0 = stile * 128 + static_cast<int>(prank) * 64;    // fused plans: CL == 2, stile is the tile
                    tma_load_2d_2sm(&tmW0, &full_bar[s], a_dst, kb * BLOCK_K, r0, hint_w);
                    tma_load_2d_2sm(&tmW1, &full_bar[s], a_dst + 64 * BLOCK_K * 2, kb * BLOCK_K, r0, hint_w);
                } else {
                    tma_load_2d_2sm(tmW, &full_bar[s], a_dst, kb * BLOCK_K, m0, hint_w);
                }
            };
            auto load_b = [&](int s, int kbb) {
                uint8_t* b_dst = smem + static_cast<size_t>(s) * stage_bytes + a_bytes;
                for (int i = 0; i < p.n_mma; ++i) {          // tokens [i*N_mma + prank*N_mma/2, +N_mma/2): rows past T are zeros
                    if (CL == 2) {
                        tma_load_2d_2sm(&tmX, &full_bar[s], b_dst + i * b_mma_bytes, kbb * BLOCK_K,
                                        i * p.N_mma + static_cast<int>(prank) * b_half_rows, hint_x);
                    } else {
                        // this CTA fetches quarter `cpair` of those tokens for itself and for its twin in the other pair
                        const int qrows = b_half_rows >> 1;
                        tma_load_2d_2sm_mc(&tmX, &full_bar[s], b_dst + i * b_mma_bytes + cpair * (b_mma_bytes >> 1), kbb * BLOCK_K,
                                           i * p.N_mma + static_cast<int>(prank) * b_half_rows + cpair * qrows,
                                           static_cast<uint16_t>(0x5u << prank), hint_x);
                    }
                }
            };
            const int npre = n_units < p.stages ? n_units : p.stages;
            const int kb_first = kb;
            for (int i = 0; i < npre; ++i) { load_a(i); advance(); }
            trace_put(p.trace, 4, 0x10000u);                              // waiting for the predecessor grid
            asm volatile("griddepcontrol.wait;" ::: "memory");
            trace_put(p.trace, 4, 0x20000u);
            for (int i = 0, kbb = kb_first; i < npre; ++i) { load_b(i, kbb); if (++kbb == KB) kbb = 0; }
            int s = npre == p.stages ? 0 : npre;
            uint32_t ph = npre == p.stages ? 1 : 0;
            for (int u = u_begin + npre; u < u_end; ++u) {
                trace_put(p.trace, 4, 0x30000u | static_cast<unsigned>(u - u_begin));
                mbar_wait(&empty_bar[s], ph ^ 1, wc, HANG_B_EMPTY, s, u);
                load_a(s);
                load_b(s, kb);
                advance();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            trace_put(p.trace, 4, 0xFFFFFu);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the LEADER CTA =====
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc_m(256, static_cast<uint32_t>(p.N_mma));
            int s = 0; uint32_t ph = 0;
            int kb = u_begin % KB;
            int buf = 0;
            uint32_t use0 = 0, use1 = 0;
            for (int u = u_begin; u < u_end; ++u) {
                const bool seg_start = (u == u_begin) || kb == 0;
                const uint32_t used = buf ? use1 : use0;
                if (seg_start && used > 0) {
                    mbar_wait(&accum_empty[buf], (used - 1) & 1, wc, HANG_B_ACCUM_EMPTY, buf, u);   // both CTAs' epilogues drained it
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t tacc = tmem_base + buf * p.buf_stride;
                trace_put(p.trace, 5, 0x30000u | static_cast<unsigned>(u - u_begin));
                mbar_wait(&full_bar[s], ph, wc, HANG_B_FULL, s, u);                              // both CTAs' tiles of this stage have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint32_t acc = (!seg_start || k > 0) ? 1u : 0u;
                    const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2);
                    umma_bf16_2sm(tacc, da, make_smem_desc(b_addr + k * UMMA_K * 2), idesc, acc);
                    if (p.n_mma == 2)
                        umma_bf16_2sm(tacc + p.N_mma, da, make_smem_desc(b_addr + b_mma_bytes + k * UMMA_K * 2), idesc, acc);
                }
                umma_commit_2sm(&empty_bar[s], static_cast<uint16_t>((1u << CL) - 1u));   // this pair is done with the slot: every CTA of the cluster hears it
                if (++s == p.stages) { s = 0; ph ^= 1; }
                if (++kb == KB) kb = 0;
                if (u + 1 == u_end || kb == 0) {
                    umma_commit_2sm(&accum_full[buf], static_cast<uint16_t>(3u << lead_rank));   // accumulators complete in both CTAs of the pair
                    if (buf) ++use1; else ++use0;
                    if (p.n_bufs == 2) buf ^= 1;
                }
            }
            trace_put(p.trace, 5, 0xFFFFFu);
        }
        __syncwarp();                                        // .aligned cluster barrier / dealloc below need the whole warp
    } else {
        // ===== epilogue (both CTAs): this CTA's 128 rows of the pair's tile =====
        if (threadIdx.x == 64) trace_put(p.trace, 6, 0x10000u);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (threadIdx.x == 64) trace_put(p.trace, 6, 0x20000u);
        const int q = warp & 3;
        float* stage_out = reinterpret_cast<float*>(smem + static_cast<size_t>(p.stages) * stage_bytes);   // fused epilogues only
        int seg = 0;
        const bool elected = threadIdx.x == 64;
        const int f = q * 32 + lane;
        if (p.epi.kind != EPI_SLICES) {
            fused_epilogue<true>(p, pair, static_cast<int>(prank), u_begin, u_end, tmem_base, stage_out, accum_full, accum_empty, wc, s_pos, s_slotuser, s_kvoff);
        } else
        for (int u = u_begin; u < u_end; ++seg) {
            const int stile = u / KB;
            const int seg_end = min((stile + 1) * KB, u_end);
            const int tile = my_tile(stile);
            int wid = 0, m0 = 0;
            const bool phantom = tile >= n_tiles_total;
            if (!phantom) tile_of(tile, wid, m0);
            const int slice = pair - (stile * KB) / p.U;
            const int buf = p.n_bufs == 2 ? (seg & 1) : 0, use = p.n_bufs == 2 ? (seg >> 1) : seg;
            if (elected) trace_put(p.trace, 6, 0x30000u | static_cast<unsigned>(seg));
            mbar_wait(&accum_full[buf], use & 1, wc, HANG_B_ACCUM_FULL, buf, u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (phantom) {                                            // nothing to store: hand the accumulator back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) mbar_arrive_cluster(&accum_empty[buf], lead_rank);
                u = seg_end;
                continue;
            }
            const uint32_t tbase = tmem_base + buf * p.buf_stride + (static_cast<uint32_t>(q * 32) << 16);
            {
                const int row = m0 + f;
                float* out = p.out + static_cast<long long>(slice) * p.slice_stride + p.colbase[wid] + row;
                drain_segment_stg(out, p.ldo, p.T, row < p.n_rows[wid], tbase, (min(T64, p.T) + 15) >> 4, [&]() {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0) mbar_arrive_cluster(&accum_empty[buf], lead_rank);
                });
                u = seg_end;
            }
        }
        if (elected) trace_put(p.trace, 6, 0xFFFFFu);
    }
    if (threadIdx.x == 0) trace_put(p.trace, 2, 5);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                      // nobody frees TMEM / exits while the peer may still touch it
    if (threadIdx.x == 0) trace_put(p.trace, 2, 6);
    if (warp == 1) {
        if (lane == 0) trace_put(p.trace, 3, 7);
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(static_cast<uint32_t>(p.tmem_cols)));
        if (lane == 0) trace_put(p.trace, 3, 8);
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

// Forwards of more than 256 tokens (cohort forwards) run the CTA-pair kernel (cta_group::2); ATSPEED_GEMM_2CTA=0 keeps the
// single-CTA kernel for every T, ATSPEED_GEMM_2CTA_MIN sets the smallest padded token count that uses the pair kernel
// (experiments).  The same rule holds at every GPU count.  (Round 1's device stall of this kernel was the missing cluster
// barrier in front of tcgen05.alloc.cta_group::2 -- see the kernel's prologue and DESIGN.md section 6; tests/test_zz_gpu_soak.py
// is its regression test.)  The environment is read per call: tests toggle it inside one process.
bool gemm_use_2cta(int T) {
    const char* e = getenv("ATSPEED_GEMM_2CTA");
    if (e && atoi(e) == 0) return false;
    const char* m = getenv("ATSPEED_GEMM_2CTA_MIN");
    const int T_pad = (T + 15) & ~15;
    return T_pad > (m ? atoi(m) : 256);
}

// Cluster size of the CTA-pair kernel: 4 (two pairs sharing the activation stream by multicast) when ATSPEED_GEMM_CLUSTER=4 and
// the device can keep num_sms / 4 such clusters resident (a cluster lives inside one GPC: 148 SMs do not always tile into 37
// clusters of 4); 2 otherwise.  Read per call like ATSPEED_GEMM_2CTA.
// Token padding of a CTA-pair launch: every TMA box of the activations must be a whole number of 8-row swizzle atoms and the
// MMA's N a multiple of 16.  Two MMAs (more than 256 tokens): N_mma = Tp / 2 and a CTA's box holds N_mma / 2 rows (N_mma / 4 in
// a cluster of 4) -> Tp is a multiple of 32 (64).  One MMA: N_mma = Tp, boxes of Tp / 2 (Tp / 4) rows -> 16 (32).
static int pair_pad_tokens(int T, int cluster, int* n_mma) {
    const int T16 = (T + 15) & ~15;
    *n_mma = T16 > 256 ? 2 : 1;
    const int q = (*n_mma == 2 ? 32 : 16) * (cluster == 4 ? 2 : 1);
    return (T + q - 1) / q * q;
}
static int g_max_clusters4[64];      // per device: co-resident 4-CTA clusters of gemm_wx_tcgen05_2cta<4> (0 = not queried / none)
static int gemm_init_device(int* max_dyn_out);
// workers (clusters of 4) a cluster-of-4 plan may use on this device, 0 = use CTA pairs
static int gemm_cluster4_workers(int num_sms) {
    const char* e = getenv("ATSPEED_GEMM_CLUSTER");
    if (!e || atoi(e) != 4 || num_sms < 4) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return num_sms / 4; }   // no device (CPU tests of the plan arithmetic)
    int md = 0;
    if (dev < 0 || dev >= 64 || gemm_init_device(&md) != ATS_OK) return 0;
    const int n = g_max_clusters4[dev] < num_sms / 4 ? g_max_clusters4[dev] : num_sms / 4;
    return n >= 1 ? n : 0;
}
int gemm_cluster_size(int num_sms) { return gemm_cluster4_workers(num_sms) > 0 ? 4 : 2; }

// Choose the tile height, the persistent grid and the unit range of every CTA for one GEMM shape.
//   allow_cut : tiles may be cut along K across CTAs (partial-sum slices; consumers reduce via SplitMap).  When false
//               (lm_head: its consumer, kernel (a), reads plain fp32 logits) whole tiles are dealt out instead.
int gemm_make_plan(const GemmWeights& w, int T, int num_sms, bool allow_cut, GemmPlan* pl) {
    ATS_CHECK_ARG(T >= 1 && T <= 512, "gemm: T=%d out of range [1,512]", T);
    ATS_CHECK_ARG(w.n >= 1 && w.n <= 3, "gemm: %d weight matrices", w.n);
    ATS_CHECK_ARG(num_sms >= 1, "gemm: num_sms=%d", num_sms);
    memset(pl, 0, sizeof(*pl));
    pl->T = T;
    pl->T_pad = (T + 15) & ~15;
    pl->KB = (w.K + BLOCK_K - 1) / BLOCK_K;
    if (gemm_use_2cta(T) && num_sms >= 2) {
        // CTA-pair kernel (gemm_wx_tcgen05_2cta): 256-row tiles owned by SM pairs, tokens padded to 64 and covered by one
        // or two M=256 MMAs of N_mma columns (each CTA holds N_mma/2 tokens of each).  cluster = 4: the worker is a cluster of
        // two pairs on adjacent tiles (a super-tile) sharing the activation stream by multicast.
        pl->two_cta = 1;
        pl->BM = 256;
        pl->total_tiles = 0;
        for (int i = 0; i < 3; ++i) {
            pl->tiles[i] = i < w.n ? (w.rows[i] + 255) / 256 : 0;
            pl->tilebase[i] = pl->total_tiles;
            pl->total_tiles += pl->tiles[i];
        }
        const int w4 = gemm_cluster4_workers(num_sms);
        pl->cluster = w4 > 0 && pl->total_tiles >= 2 ? 4 : 2;
        const int T64 = pair_pad_tokens(T, pl->cluster, &pl->n_mma);   // (named for its round-1 granularity)
        pl->N_mma = T64 / pl->n_mma;
        const int tps = pl->cluster / 2;
        const int n_super = (pl->total_tiles + tps - 1) / tps;         // units are (super-tile, k-block)
        const int pairs = pl->cluster == 4 ? w4 : num_sms / 2;         // workers
        const int units = n_super * pl->KB;
        if (allow_cut) {
            const int grid = units < pairs ? units : pairs;
            pl->U = (units + grid - 1) / grid;
            const int min_u = pl->KB < 8 ? pl->KB : 8;
            if (pl->U < min_u) pl->U = min_u;
            const int s = pairs / n_super;              // narrow outputs: whole k-splits per tile (one segment per worker)
            if (s >= 2 && s <= pl->KB && n_super * s * 10 >= pairs * 8) pl->U = (pl->KB + s - 1) / s;
        } else {
            pl->U = ((n_super + pairs - 1) / pairs) * pl->KB;
        }
        pl->grid = pl->cluster * ((units + pl->U - 1) / pl->U);
        pl->max_slices = 1;
        for (int t = 0; t < n_super; ++t) {
            const int n = (t * pl->KB + pl->KB - 1) / pl->U - (t * pl->KB) / pl->U + 1;
            if (n > pl->max_slices) pl->max_slices = n;
        }
        const int stage_bytes = A_TILE_BYTES + (T64 / 2) * BLOCK_K * 2;
        int stages = (220 * 1024) / stage_bytes;
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        if (stages > pl->U) stages = pl->U < 2 ? 2 : pl->U;
        pl->stages = stages;
        int acc = 32;
        while (acc < T64) acc <<= 1;
        pl->acc_stride = acc;
        pl->n_bufs = 2 * acc <= 512 ? 2 : 1;
        pl->buf_stride = acc;
        pl->tmem_cols = pl->n_bufs * acc;
        return ATS_OK;
    }
    int tiles128 = 0;
    for (int i = 0; i < w.n; ++i) tiles128 += (w.rows[i] + 127) / 128;
    // Two stacked 128-row MMAs per activation tile halve the L2->SM re-reads of the activations (the limiter at
    // T >~ 100, profiles/r01_ncu_full_v2.txt); they need 2 x T_pad TMEM columns and only pay off on wide outputs.
    // (gemm_sweep: at T_pad <= 128 BM = 256 streams 4.7-5.9 TB/s where BM = 128 stalls at 3.3-3.8 -- one CTA per SM keeps too
    // few TMA boxes in flight; beyond that the [T_pad, 64] activation box dominates and the extra partial-sum traffic of
    // 256-row tiles costs more than it saves)
    pl->BM = (pl->T_pad <= 128 && tiles128 >= 2) ? 256 : 128;
    pl->total_tiles = 0;
    for (int i = 0; i < 3; ++i) {
        pl->tiles[i] = i < w.n ? (w.rows[i] + pl->BM - 1) / pl->BM : 0;
        pl->tilebase[i] = pl->total_tiles;
        pl->total_tiles += pl->tiles[i];
    }
    const int units = pl->total_tiles * pl->KB;
    if (allow_cut) {
        int grid = units < num_sms ? units : num_sms;
        pl->U = (units + grid - 1) / grid;
        // small projections (the 68M draft): a CTA needs a few k-blocks to amortise its prologue / epilogue, and every extra
        // cut is one more slice for the consumer
        const int min_u = pl->KB < 8 ? pl->KB : 8;
        if (pl->U < min_u) pl->U = min_u;
        // Narrow outputs at large T (o / down projections): the partial-sum slices and their epilogues cost as much as the
        // weights, so prefer a whole number of k-splits per tile (one segment per CTA, `s` slices) when that still fills
        // >= 80 % of the SMs (gemm_sweep: o 26 -> 21 us, down 41 -> 31 us at T = 220).
        const int s = num_sms / pl->total_tiles;
        if (pl->T_pad > 128 && s >= 2 && s <= pl->KB && pl->total_tiles * s * 10 >= num_sms * 8) pl->U = (pl->KB + s - 1) / s;
    } else {
        const int per = (pl->total_tiles + num_sms - 1) / num_sms;
        pl->U = per * pl->KB;
    }
    pl->grid = (units + pl->U - 1) / pl->U;
    pl->max_slices = 1;
    for (int t = 0; t < pl->total_tiles; ++t) {
        const int n = (t * pl->KB + pl->KB - 1) / pl->U - (t * pl->KB) / pl->U + 1;
        if (n > pl->max_slices) pl->max_slices = n;
    }
    const int stage_bytes = pl->BM * BLOCK_K * 2 + pl->T_pad * BLOCK_K * 2;
    int stages = (220 * 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > pl->U) stages = pl->U < 2 ? 2 : pl->U;
    ATS_CHECK_ARG(stages >= 2, "gemm: T=%d leaves room for %d pipeline stages", T, stages);
    pl->stages = stages;
    int acc = 32;
    while (acc < pl->T_pad) acc <<= 1;                       // columns of one 128-row accumulator (power of two)
    pl->acc_stride = acc;
    const int per_buf = pl->BM == 256 ? 2 * acc : acc;
    pl->n_bufs = 2 * per_buf <= 512 ? 2 : 1;                 // double-buffer so an epilogue overlaps the next segment
    pl->buf_stride = per_buf;
    pl->tmem_cols = pl->n_bufs * per_buf;
    ATS_CHECK_ARG(pl->tmem_cols <= 512, "gemm: %d TMEM columns", pl->tmem_cols);
    return ATS_OK;
}

// Plan of a fused-epilogue launch: plain stream-K (equal contiguous unit ranges, no whole-k-split variant: a tile cut in two
// costs one partial store + one partial load inside the kernel instead of a consumer pass).  EPI_SILU_MUL tiles hold the gate
// rows and the up rows of the same features: BM/2 features per tile of a single CTA, 128 per tile of a CTA pair.
int gemm_make_plan_fused(const GemmWeights& w, int T, int num_sms, int kind, GemmPlan* pl) {
    ATS_CHECK_ARG(T >= 1 && T <= 512, "gemm: T=%d out of range [1,512]", T);
    ATS_CHECK_ARG(kind == EPI_QKV_ROPE ? w.n == 3 : (kind == EPI_SILU_MUL && w.n == 2 && w.rows[0] == w.rows[1]),
                  "gemm: fused epilogue %d does not fit %d weight matrices", kind, w.n);
    ATS_CHECK_ARG(num_sms >= 1, "gemm: num_sms=%d", num_sms);
    memset(pl, 0, sizeof(*pl));
    pl->epi_kind = kind;
    pl->T = T;
    pl->T_pad = (T + 15) & ~15;
    pl->KB = (w.K + BLOCK_K - 1) / BLOCK_K;
    pl->max_slices = 1;
    const bool pair = gemm_use_2cta(T) && num_sms >= 2;
    int workers = num_sms, stage_bytes;
    if (pair) {
        pl->two_cta = 1;
        pl->cluster = 2;
        const int T64 = pair_pad_tokens(T, 2, &pl->n_mma);
        pl->N_mma = T64 / pl->n_mma;
        pl->BM = 256;
        workers = num_sms / 2;
        stage_bytes = A_TILE_BYTES + (T64 / 2) * BLOCK_K * 2;
        int acc = 32;
        while (acc < T64) acc <<= 1;
        pl->acc_stride = acc;
        pl->n_bufs = 2 * acc <= 512 ? 2 : 1;
        pl->buf_stride = acc;
        pl->tmem_cols = pl->n_bufs * acc;
    } else {
        int tiles128 = 0;
        for (int i = 0; i < w.n; ++i) tiles128 += (w.rows[i] + 127) / 128;
        pl->BM = (pl->T_pad <= 128 && tiles128 >= 2) ? 256 : 128;
        stage_bytes = pl->BM * BLOCK_K * 2 + pl->T_pad * BLOCK_K * 2;
        int acc = 32;
        while (acc < pl->T_pad) acc <<= 1;
        pl->acc_stride = acc;
        const int per_buf = pl->BM == 256 ? 2 * acc : acc;
        pl->n_bufs = 2 * per_buf <= 512 ? 2 : 1;
        pl->buf_stride = per_buf;
        pl->tmem_cols = pl->n_bufs * per_buf;
    }
    pl->total_tiles = 0;
    for (int i = 0; i < 3; ++i) {
        pl->tiles[i] = 0;
        if (kind == EPI_QKV_ROPE && i < w.n) pl->tiles[i] = (w.rows[i] + pl->BM - 1) / pl->BM;
        if (kind == EPI_SILU_MUL && i == 0) pl->tiles[i] = (w.rows[0] + pl->BM / 2 - 1) / (pl->BM / 2);
        pl->tilebase[i] = pl->total_tiles;
        pl->total_tiles += pl->tiles[i];
    }
    const int units = pl->total_tiles * pl->KB;
    const int grid = units < workers ? units : workers;
    pl->U = (units + grid - 1) / grid;
    const int min_u = pl->KB < 8 ? pl->KB : 8;
    if (pl->U < min_u) pl->U = min_u;
    pl->grid = (pair ? 2 : 1) * ((units + pl->U - 1) / pl->U);
    int stages = (220 * 1024 - OUT_STAGE_BYTES) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > pl->U) stages = pl->U < 2 ? 2 : pl->U;
    ATS_CHECK_ARG(stages >= 2, "gemm: T=%d leaves room for %d pipeline stages", T, stages);
    pl->stages = stages;
    ATS_CHECK_ARG(pl->tmem_cols <= 512, "gemm: %d TMEM columns", pl->tmem_cols);
    return ATS_OK;
}

// floats of FusedEpi::part: one [T_pad][128] tile per CTA and 128-row half (single CTA, BM = 256 at T <= 128: two halves)
size_t gemm_fused_part_elems(int T_max, int num_sms) {
    const size_t t = static_cast<size_t>(((T_max + 15) & ~15) > 256 ? ((T_max + 15) & ~15) : 256);
    return static_cast<size_t>(num_sms) * t * 128;
}

// upper bound of partial-sum slices over every T the plan function can be asked for (workspace sizing)
int gemm_max_slices(const GemmWeights& w, int T_max, int num_sms) {
    int best = 1;
    for (int T : {1, 32, 64, 128, 256, T_max}) {
        if (T > T_max) continue;
        GemmPlan pl;
        if (gemm_make_plan(w, T, num_sms, true, &pl) == ATS_OK && pl.max_slices > best) best = pl.max_slices;
    }
    return best + 1;
}

SplitMap gemm_split_map(const GemmWeights& w, const GemmPlan& pl) {
    SplitMap m;
    memset(&m, 0, sizeof(m));
    m.n = w.n;
    for (int i = 0; i < 3; ++i) { m.colbase[i] = i < w.n ? w.colbase[i] : 0x7fffffff; m.tilebase[i] = pl.tilebase[i]; }
    m.BM = pl.BM; m.KB = pl.KB; m.U = pl.U;
    m.tps = pl.two_cta && pl.cluster == 4 ? 2 : 1;
    m.tab_n = 0; m.bm_shift = 0;
    memset(m.tab, 0, sizeof(m.tab));
    if (pl.total_tiles <= SPLIT_TAB && (pl.BM & (pl.BM - 1)) == 0) {
        while ((1 << m.bm_shift) < pl.BM) ++m.bm_shift;
        for (int t = 0; t < pl.total_tiles; ++t) m.tab[t] = static_cast<unsigned char>(m.slices_of_tile(t));
        m.tab_n = pl.total_tiles;
    }
    return m;
}

int gemm_make_xmap(XMap* xm, const void* x, int T, int K, int cluster) {
    if (gemm_use_2cta(T)) {
        // CTA-pair kernel: one box = the N_mma/2 tokens of one MMA that one CTA of the pair holds (cluster of 4: half of them,
        // the other half arrives by multicast from the twin CTA)
        int n_mma = 1;
        const int T64 = pair_pad_tokens(T, cluster, &n_mma);
        const int box = T64 / n_mma / 2 / (cluster == 4 ? 2 : 1);
        ATS_TRY(make_tmap_bf16_kmajor(&xm->tm0, x, T, K, box));
        xm->tm1 = xm->tm0;
        xm->T = T; xm->K = K; xm->box0 = box; xm->cluster = cluster;
        return ATS_OK;
    }
    xm->cluster = 0;
    const int T_pad = (T + 15) & ~15;
    // activations: tokens 0..255 through tm0 (box = min(T_pad,256) rows), tokens 256..T_pad-1 through tm1
    // (box = T_pad-256 rows); rows past T are zero-filled by TMA, and each box always delivers its full byte count.
    int box0 = T_pad > 256 ? 256 : T_pad;
    xm->box0 = box0;
    ATS_TRY(make_tmap_bf16_kmajor(&xm->tm0, x, T, K, box0));
    ATS_TRY(make_tmap_bf16_kmajor(&xm->tm1, x, T, K, T_pad > 256 ? T_pad - 256 : 16));
    xm->T = T; xm->K = K;
    return ATS_OK;
}

// Where one launch's fp32 partial sums go: out[slice][t][colbase_i + n], slices `slice_stride` elements apart.
int gemm_make_omap(OMap* om, const GemmWeights& w, float* out, int ldo, long long slice_stride, int T, int max_slices) {
    (void)w; (void)max_slices;
    memset(om, 0, sizeof(*om));
    om->out = out; om->ldo = ldo; om->slice_stride = slice_stride; om->T = T;
    return ATS_OK;
}

// Per-device one-time set-up: both kernels may use the whole 227 KiB of shared memory.  Done under std::call_once so the
// lane threads of one process (bench.py runs three) cannot race on it.
static int gemm_init_device(int* max_dyn_out) {
    static std::once_flag once[64];
    static int max_dyn[64];
    static cudaError_t err[64];
    int dev = 0;
    ATS_CUDA(cudaGetDevice(&dev));
    ATS_CHECK_ARG(dev >= 0 && dev < 64, "gemm: device ordinal %d", dev);
    std::call_once(once[dev], [dev]() {
        cudaFuncAttributes fa;
        int md = 227 * 1024;
        err[dev] = cudaSuccess;
        for (const void* fn : {reinterpret_cast<const void*>(gemm_wx_tcgen05), reinterpret_cast<const void*>(gemm_wx_tcgen05_2cta<2>),
                               reinterpret_cast<const void*>(gemm_wx_tcgen05_2cta<4>)}) {
            cudaError_t e = cudaFuncGetAttributes(&fa, fn);
            const int want = 227 * 1024 - static_cast<int>(fa.sharedSizeBytes);   // static barriers share the 227 KiB budget
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            if (e != cudaSuccess) { err[dev] = e; return; }
            if (want < md) md = want;
        }
        max_dyn[dev] = md;
        // how many clusters of 4 (one GEMM CTA per SM) can be resident at once: decides whether the cluster-of-4 plans are used
        {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            int sms = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cfg.gridDim = dim3((sms / 4) * 4); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = md;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, gemm_wx_tcgen05_2cta<4>, &cfg) == cudaSuccess) g_max_clusters4[dev] = n;
            else cudaGetLastError();
        }
    });
    if (err[dev] != cudaSuccess) { set_error("gemm: shared-memory set-up failed: %s", cudaGetErrorString(err[dev])); return ATS_ERR_CUDA; }
    *max_dyn_out = max_dyn[dev];
    return ATS_OK;
}

// out[slice][t][...]. X: [T, K] bf16 row-major (through xm). W_i: [n_rows_i, K] bf16 row-major.  om (EPI_SLICES plans) or
// epi (fused plans) says where the result goes.
static int gemm_launch(const GemmWeights& w, const XMap& xm, const GemmPlan& pl, const OMap* om, const FusedEpi* epi,
                       cudaStream_t stream) {
    ATS_CHECK_ARG((om != nullptr) != (epi != nullptr) && (epi ? epi->kind == pl.epi_kind : pl.epi_kind == EPI_SLICES),
                  "gemm: output description does not match the plan (epilogue kind %d)", pl.epi_kind);
    ATS_CHECK_ARG(!om || om->T == pl.T, "gemm: output map built for T=%d, plan for T=%d", om ? om->T : 0, pl.T);
    ATS_CHECK_ARG(xm.T == pl.T && xm.K == w.K, "gemm: activation map (%d x %d) does not match the plan (%d x %d)", xm.T,
                  xm.K, pl.T, w.K);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    if (om) { p.out = om->out; p.ldo = om->ldo; p.slice_stride = om->slice_stride; }
    if (epi) {
        p.epi = *epi;
        ATS_CHECK_ARG(epi->part && epi->flags && epi->epoch != 0, "gemm: fused epilogue without workspace / epoch");
        if (epi->kind == EPI_QKV_ROPE)
            ATS_CHECK_ARG(epi->head_dim >= 16 && 128 % epi->head_dim == 0 && epi->HD % 8 == 0 && epi->pos && epi->slot && epi->qbuf &&
                              epi->kcache && epi->vcache && epi->rope_cos && epi->rope_sin,
                          "gemm: RoPE epilogue needs head_dim in {16,32,64,128} (got %d) and all of its buffers", epi->head_dim);
        else
            ATS_CHECK_ARG(epi->m && epi->mlp % 8 == 0 && epi->mlp == w.rows[0], "gemm: SiLU epilogue: mlp=%d", epi->mlp);
    }
    p.T = pl.T; p.T_pad = pl.T_pad; p.n_kblocks = pl.KB;
    for (int i = 0; i < 3; ++i) {
        p.n_rows[i] = i < w.n ? w.rows[i] : 0;
        p.tiles[i] = pl.tiles[i];
        p.colbase[i] = i < w.n ? w.colbase[i] : 0;
    }
    const int tps = pl.two_cta && pl.cluster == 4 ? 2 : 1;
    p.BM = pl.BM; p.U = pl.U; p.total_units = ((pl.total_tiles + tps - 1) / tps) * pl.KB;
    p.stages = pl.stages; p.tmem_cols = pl.tmem_cols; p.acc_stride = pl.acc_stride;
    p.n_bufs = pl.n_bufs; p.buf_stride = pl.buf_stride;
    p.b_box_bytes = (pl.T_pad > 256 ? pl.T_pad : xm.box0) * BLOCK_K * 2;
    p.n_mma = pl.n_mma; p.N_mma = pl.N_mma;
    p.guard = spin_guard();
    { static const bool on = []() { const char* e = getenv("ATSPEED_GEMM_L2HINT"); return !(e && atoi(e) == 0); }(); p.l2_hints = on ? 1 : 0; }
    p.trace = pl.two_cta ? gemm_trace() : GemmTrace{nullptr, 0};
    const int stage_bytes = pl.two_cta ? A_TILE_BYTES + (pl.n_mma * pl.N_mma / 2) * BLOCK_K * 2
                                       : pl.BM * BLOCK_K * 2 + pl.T_pad * BLOCK_K * 2;
    size_t smem_bytes = static_cast<size_t>(p.stages) * stage_bytes + (epi ? OUT_STAGE_BYTES : 0) + 1024;
    // One GEMM CTA per SM, always: a CTA holds its TMEM columns from prologue to exit and, with PDL and several streams,
    // CTAs of different launches overlap in time.  A CTA that was launched early (PDL) and already holds TMEM while it
    // waits for its predecessor must never share an SM with a predecessor CTA that has not allocated yet; asking for more
    // than half of the SM's shared memory makes co-residency impossible.
    if (smem_bytes < EXCLUSIVE_SMEM_BYTES) smem_bytes = EXCLUSIVE_SMEM_BYTES;
    int max_dyn = 0;
    ATS_TRY(gemm_init_device(&max_dyn));
    ATS_CHECK_ARG(static_cast<int>(smem_bytes) <= max_dyn, "gemm: %zu bytes of shared memory > %d", smem_bytes, max_dyn);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (pl.two_cta) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = pl.cluster; attr[n_attr].val.clusterDim.y = 1; attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    if (pdl_enabled()) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue + weight prefetch overlap the
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;            // previous kernel; see griddepcontrol.wait
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    // weight operand maps: 128-row boxes, 256-row boxes for the stacked single-CTA tiles, and for the interleaved SiLU tiles
    // the gate / up halves (BM/2 rows per box of a single CTA, 64 rows per CTA of a pair)
    const CUtensorMap* tw = (!pl.two_cta && pl.BM == 256) ? w.tmap256 : w.tmap;
    if (pl.epi_kind == EPI_SILU_MUL) tw = (pl.two_cta || pl.BM == 128) ? w.tmap64 : w.tmap;
    if (pl.two_cta) {
        ATS_CHECK_ARG((pl.cluster == 2 || pl.cluster == 4) && pl.grid % pl.cluster == 0, "gemm (CTA pairs): grid %d, cluster %d", pl.grid, pl.cluster);
        ATS_CHECK_ARG(xm.cluster == pl.cluster, "gemm: activation map built for clusters of %d, plan uses %d", xm.cluster, pl.cluster);
        if (pl.cluster == 4)
            ATS_CUDA(cudaLaunchKernelEx(&cfg, gemm_wx_tcgen05_2cta<4>, tw[0], tw[w.n > 1 ? 1 : 0], tw[w.n > 2 ? 2 : 0], xm.tm0, p));
        else
            ATS_CUDA(cudaLaunchKernelEx(&cfg, gemm_wx_tcgen05_2cta<2>, tw[0], tw[w.n > 1 ? 1 : 0], tw[w.n > 2 ? 2 : 0], xm.tm0, p));
        return ATS_OK;
    }
    ATS_CUDA(cudaLaunchKernelEx(&cfg, gemm_wx_tcgen05, tw[0], tw[w.n > 1 ? 1 : 0], tw[w.n > 2 ? 2 : 0], xm.tm0, xm.tm1, p));
    return ATS_OK;
}

int gemm_wx(const GemmWeights& w, const XMap& xm, const GemmPlan& pl, const OMap& om, cudaStream_t stream) {
    return gemm_launch(w, xm, pl, &om, nullptr, stream);
}

int gemm_wx_fused(const GemmWeights& w, const XMap& xm, const GemmPlan& pl, const FusedEpi& epi, cudaStream_t stream) {
    return gemm_launch(w, xm, pl, nullptr, &epi, stream);
}

bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("ATSPEED_PDL"); return !(e && atoi(e) == 0); }();   // thread-safe static init
    return on;
}

}  // namespace atspeed
