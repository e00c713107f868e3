// Kernel (c): KV-cache row gather / beam reorder with TMA bulk copies.
//
// Replaces (reference file:line relative to /root/reference/code):
//   * the KV "truncation" after verify, beamSD.py:418-429, whose sliced views are re-copied in full by
//     the next forward's torch.cat: here only the ancestor rows of the K surviving
//     beams are moved from the round's tree region into the dense accepted region;
//   * HF's per-layer `index_select(0, beam_idx)` cache reorder used by the `TF_target` baseline
//     (code/inference.py:178): same kernel, out-of-place, rows = (beam, position).
//
// The cache is token-major ([slot][heads*head_dim] per layer and per K/V plane), so one token's K (or
// V) for all heads is ONE contiguous row (8 KiB for the 7B shape) -- a single cp.async.bulk
// global->shared followed by a bulk shared->global store, issued by one thread per CTA.  HBM-bound:
// algorithmic bytes = 2 * rows * planes * row_bytes.  grid = (rows, planes) keeps thousands of bulk
// copies in flight.  Source and destination row sets must be disjoint (the engine copies from the
// tree region into the accepted region), so no ordering between CTAs is needed.
#include "common.cuh"
#include "kernels.h"

namespace atspeed {

static constexpr int GATHER_CHUNK = 16 * 1024;

// one thread: row `src` -> shared memory -> row `dst`, GATHER_CHUNK bytes at a time (bounded wait: common.cuh)
__device__ __forceinline__ void copy_row_bulk(const uint8_t* src, uint8_t* dst, int row_bytes, uint8_t* buf, uint64_t* bar,
                                              const SpinGuard& guard) {
    const uint32_t bar_a = smem_u32(bar), buf_a = smem_u32(buf);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t phase = 0;
    for (int off = 0; off < row_bytes; off += GATHER_CHUNK) {
        const uint32_t n = static_cast<uint32_t>(row_bytes - off < GATHER_CHUNK ? row_bytes - off : GATHER_CHUNK);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf_a),
                     "l"(src + off), "r"(n), "r"(bar_a)
                     : "memory");
        mbar_wait_guarded(bar_a, phase, guard, HANG_K_KVGATHER, HANG_R_COPY, HANG_B_ROW, 0, static_cast<unsigned>(off), 0,
                          static_cast<unsigned>(row_bytes), 0);
        phase ^= 1;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(buf_a), "r"(n)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // buffer reusable / safe to exit
    }
}

__global__ void __launch_bounds__(32)
kv_gather_bulk_kernel(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base, long long src_plane_stride,
                      long long dst_plane_stride, int row_bytes, const int* __restrict__ src_rows,
                      const int* __restrict__ dst_rows, const int* __restrict__ n_rows_dev, SpinGuard guard) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    const int i = blockIdx.x, plane = blockIdx.y;
    if (n_rows_dev != nullptr && i >= *n_rows_dev) return;
    if (threadIdx.x != 0) return;
    copy_row_bulk(src_base + plane * src_plane_stride + static_cast<long long>(src_rows[i]) * row_bytes,
                  dst_base + plane * dst_plane_stride + static_cast<long long>(dst_rows[i]) * row_bytes, row_bytes, buf, &bar, guard);
}

int kv_gather_rows_oop(const void* src_base, void* dst_base, long long src_plane_stride, long long dst_plane_stride,
                       int n_planes, int row_bytes, const int* src, const int* dst, const int* n_rows_dev,
                       int max_rows, cudaStream_t st) {
    ATS_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "kv_gather: row_bytes=%d must be a multiple of 16", row_bytes);
    ATS_CHECK_ARG((reinterpret_cast<uintptr_t>(src_base) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_base) & 15) == 0 &&
                      src_plane_stride % 16 == 0 && dst_plane_stride % 16 == 0,
                  "kv_gather: bases and plane strides must be 16-byte aligned");
    ATS_CHECK_ARG(n_planes >= 1 && n_planes <= 65535 && max_rows >= 1, "kv_gather: planes=%d rows=%d", n_planes, max_rows);
    const int smem = row_bytes < GATHER_CHUNK ? row_bytes : GATHER_CHUNK;
    dim3 grid(max_rows, n_planes);
    kv_gather_bulk_kernel<<<grid, 32, smem, st>>>(static_cast<const uint8_t*>(src_base), static_cast<uint8_t*>(dst_base),
                                                  src_plane_stride, dst_plane_stride, row_bytes, src, dst, n_rows_dev, spin_guard());
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// cohort variant: grid.z = user; user u moves rows inside its own cache (base + byte_off[u]) using its own lists
__global__ void __launch_bounds__(32)
kv_gather_cohort_kernel(uint8_t* __restrict__ base, long long plane_stride, int row_bytes, GatherCohort gc, SpinGuard guard) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    const int i = blockIdx.x, plane = blockIdx.y, u = blockIdx.z;
    if (i >= *gc.n_rows[u]) return;
    if (threadIdx.x != 0) return;
    uint8_t* ub = base + gc.byte_off[u] + plane * plane_stride;
    copy_row_bulk(ub + static_cast<long long>(gc.src[u][i]) * row_bytes, ub + static_cast<long long>(gc.dst[u][i]) * row_bytes,
                  row_bytes, buf, &bar, guard);
}

int kv_gather_rows_cohort(void* base, long long plane_stride, int n_planes, int row_bytes, const GatherCohort& gc,
                          int max_rows, cudaStream_t st) {
    ATS_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "kv_gather: row_bytes=%d must be a multiple of 16", row_bytes);
    ATS_CHECK_ARG(gc.n >= 1 && gc.n <= MAX_USERS && max_rows >= 1 && n_planes >= 1, "kv_gather: users=%d rows=%d", gc.n, max_rows);
    const int smem = row_bytes < GATHER_CHUNK ? row_bytes : GATHER_CHUNK;
    dim3 grid(max_rows, n_planes, gc.n);
    kv_gather_cohort_kernel<<<grid, 32, smem, st>>>(static_cast<uint8_t*>(base), plane_stride, row_bytes, gc, spin_guard());
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

int kv_gather_rows(void* base, long long plane_stride, int n_planes, int row_bytes, const int* src, const int* dst,
                   const int* n_rows_dev, int max_rows, cudaStream_t st) {
    return kv_gather_rows_oop(base, base, plane_stride, plane_stride, n_planes, row_bytes, src, dst, n_rows_dev,
                              max_rows, st);
}

}  // namespace atspeed
