// Shared device/host helpers for the atspeed_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "atspeed_b200 kernels are written for sm_100a only"
#endif

namespace atspeed {

// ---------------------------------------------------------------------------------------------
// error reporting: every C-ABI entry returns 0 or a negative code; text via atspeed_last_error()
// ---------------------------------------------------------------------------------------------
enum : int {
    ATS_OK = 0,
    ATS_ERR_ARG = -1,      // bad argument (shape, alignment, limit)
    ATS_ERR_CUDA = -2,     // CUDA runtime / driver error
    ATS_ERR_STATE = -3,    // call out of order for the session state
    ATS_ERR_LIMIT = -4,    // exceeds a compiled limit
};

void set_error(const char* fmt, ...);
const char* last_error();

#define ATS_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            ::atspeed::set_error(__VA_ARGS__);        \
            return ::atspeed::ATS_ERR_ARG;            \
        }                                             \
    } while (0)

#define ATS_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::atspeed::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                                 __FILE__, __LINE__);                                         \
            return ::atspeed::ATS_ERR_CUDA;                                                   \
        }                                                                                     \
    } while (0)

#define ATS_LAUNCH_CHECK()  ATS_CUDA(cudaGetLastError())

#define ATS_TRY(expr)                 \
    do {                              \
        int _r = (expr);              \
        if (_r != 0) return _r;       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// limits of the beam-tree state (see DESIGN.md "Data layout")
// ---------------------------------------------------------------------------------------------
constexpr int MAX_BEAMS = 64;      // K (target beams) <= 32 is required; N (draft beams) <= 64
constexpr int MAX_K = 32;
constexpr int MAX_LEVELS = 5;      // roots + up to 4 draft levels per round
constexpr int MAX_NEW = 6;         // max_new_tokens
constexpr int VIS_WORDS = 16;      // 512 tree/accepted KV slots addressable by a beam's visibility mask
constexpr int MAX_TREE_SLOTS = VIS_WORDS * 32;

bool pdl_enabled();   // gemm.cu; ATSPEED_PDL=0 turns programmatic dependent launch off

// ---------------------------------------------------------------------------------------------
// Bounded mbarrier waits.  Every mbarrier wait of the library (GEMM pipelines, KV gather) gives up after
// spin_limit_ns() (ATSPEED_SPIN_LIMIT_MS, default 4000; 0 = wait forever): the first thread whose wait expires
// writes WHO was waiting for WHAT into a HangDiag record in mapped pinned host memory (readable after the
// context is gone) and traps, so a protocol bug or a lost arrival ends the launch with a CUDA error and a
// message instead of a GPU that spins until somebody kills the process.  engine.cu appends the decoded record
// to atspeed_last_error() whenever a CUDA call fails.
// ---------------------------------------------------------------------------------------------
struct HangDiag {
    unsigned int flag;        // 0 clear, 1 being written, 2 valid
    unsigned int kernel;      // HANG_K_*
    unsigned int block, thread;
    unsigned int role;        // HANG_R_*
    unsigned int barrier;     // HANG_B_*
    unsigned int index;       // pipeline stage / accumulator buffer
    unsigned int parity;
    unsigned int unit, u_begin, u_end;   // GEMM: work unit being waited for and the CTA's unit range
    unsigned int T;
    unsigned long long waited_ns;
};
enum : unsigned { HANG_K_GEMM = 1, HANG_K_GEMM_PAIR = 2, HANG_K_KVGATHER = 3, HANG_K_ROWWISE = 4 };
enum : unsigned { HANG_R_PRODUCER = 1, HANG_R_MMA = 2, HANG_R_EPILOGUE = 3, HANG_R_COPY = 4 };
enum : unsigned { HANG_B_EMPTY = 1, HANG_B_FULL = 2, HANG_B_ACCUM_FULL = 3, HANG_B_ACCUM_EMPTY = 4, HANG_B_ROW = 5, HANG_B_FLAG = 6 };
struct SpinGuard {            // passed by value to the kernels that wait on mbarriers
    HangDiag* diag;           // device-visible address of the mapped record (nullptr: trap without a record)
    unsigned long long limit_ns;
};
// Optional per-CTA progress trace of the GEMM kernels (ATSPEED_GEMM_TRACE=1; diagnostics only, off by default): 16 words
// per CTA in mapped pinned host memory, readable WHILE a launch is stuck (atspeed_debug_gemm_trace) -- for the waits the
// guard above cannot bound (tcgen05.alloc, cluster barriers, griddepcontrol.wait).
struct GemmTrace {
    unsigned int* buf;        // nullptr: tracing off
    unsigned int seq;         // launch sequence number
};
constexpr int TRACE_WORDS = 16, TRACE_CTAS = 256;
SpinGuard spin_guard();                       // engine.cu: mapped record (allocated once per process) + the limit
GemmTrace gemm_trace();                       // engine.cu: {nullptr, 0} unless ATSPEED_GEMM_TRACE=1
void hang_diag_describe(char* buf, size_t n); // engine.cu: "" when no record has been written

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
static __device__ unsigned int g_hang_elect = 0;    // per translation unit: the first reporter of a launch writes the record
static __device__ __noinline__ void hang_report(const SpinGuard g, unsigned kernel, unsigned role, unsigned barrier, unsigned index,
                                         unsigned parity, unsigned unit, unsigned u_begin, unsigned u_end, unsigned T,
                                         unsigned long long waited_ns) {
    volatile HangDiag* d = g.diag;
    if (d != nullptr && atomicCAS(&g_hang_elect, 0u, 1u) == 0u) {     // plain stores only: the record is host memory
        d->flag = 1u;
        d->kernel = kernel; d->block = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); d->thread = threadIdx.x;
        d->role = role; d->barrier = barrier; d->index = index; d->parity = parity;
        d->unit = unit; d->u_begin = u_begin; d->u_end = u_end; d->T = T; d->waited_ns = waited_ns;
        __threadfence_system();
        d->flag = 2u;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait for the phase with parity `parity` of the mbarrier at shared address bar_addr, at most g.limit_ns
__device__ __forceinline__ void mbar_wait_guarded(uint32_t bar_addr, uint32_t parity, const SpinGuard& g, unsigned kernel,
                                                  unsigned role, unsigned barrier, unsigned index, unsigned unit,
                                                  unsigned u_begin, unsigned u_end, unsigned T) {
    if (mbar_try_wait_u32(bar_addr, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_wait_u32(bar_addr, parity)) return;
        if ((spins & 63u) == 0u && g.limit_ns != 0ull) {
            const unsigned long long dt = global_timer_ns() - t0;
            if (dt > g.limit_ns) hang_report(g, kernel, role, barrier, index, parity, unit, u_begin, u_end, T, dt);
        }
    }
}

// Programmatic dependent launch (PDL): every kernel of the forward is launched with the
// programmatic-stream-serialization attribute, lets its successor start early (launch_dependents) and waits for its
// predecessor's memory (wait) before touching anything the predecessor wrote or still reads.  Launch latency and
// prologues overlap the previous kernel's tail; the GEMM additionally prefetches weights before its wait.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// bf16(silu(g)) * u with g, u rounded to bf16 first: the rounding points of an HF bf16 LlamaMLP.  silu runs in fp32 on the SFU
// (ex2.approx, rcp.approx: relative error < 1e-6, three orders of magnitude below the bf16 rounding that follows).  The IEEE
// division + expf this replaces cost ~135 SASS instructions per element and made silu_mul issue-bound (20 us of issue slots at
// T = 480 against 16 us of HBM time, tools/rowwise_bench.py).
__device__ __forceinline__ float silu_mul_bf16(float g, float u) {
    g = bf16_round(g); u = bf16_round(u);
    return bf16_round(__fdividef(g, 1.0f + __expf(-g))) * u;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// total order used for every top-k in the library: higher score first, ties -> lower index first.
// key = (orderable score bits << 32) | (0xffffffff - idx): bigger key wins.
__device__ __forceinline__ unsigned long long rank_key(float score, uint32_t idx) {
    uint32_t b = __float_as_uint(score);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return (static_cast<unsigned long long>(b) << 32) | static_cast<unsigned long long>(0xffffffffu - idx);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}
#endif

}  // namespace atspeed
