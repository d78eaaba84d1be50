// row_pairs.cu -- dense int32 rows in HBM -> the (bin, count) pairs of their non-zero bins.
//
// Used by the host-buffer operator (api.cu, cfrk_count_dense_host = the reference's kmer_main contract):
// dense rows are 4^k * 4 bytes per read (256 KiB at k = 8, of which <= 150 words are not zero for a 150-bp
// read) and PCIe moves 55 GB/s, so for part of the reads the finished rows -- counted on the GPU like all
// the others, compat spill included -- are compacted here, only the pairs cross the bus (<= 1.2 KB per
// read) and host threads expand them into the caller's buffer with streaming stores, while the DMA engine
// carries the dense rows of the other reads.  The host never counts anything: it writes zeros and copies
// counts.
//
// One warp per row; 128 bins per step (one 16-byte load per lane), all-zero steps skipped after one
// ballot.  Row r's pairs go to keys/counts[off[r] ..], off[] given by the caller (capacity per row =
// min(4^k, visited windows + 1): every non-zero bin holds at least one window or the spill).
#include "kernels.h"

namespace cfrk {

extern void count_launch();

__global__ void __launch_bounds__(256) rows_to_pairs_kernel(const int32_t* __restrict__ rows, int64_t nrows, int bins,
                                                            const int64_t* __restrict__ off, uint32_t* __restrict__ keys,
                                                            uint32_t* __restrict__ counts, int32_t* __restrict__ row_count)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const int4* row = reinterpret_cast<const int4*>(rows + r * bins);
        const int64_t base = off[r];
        int n = 0;
        for (int i0 = 0; i0 < bins; i0 += 128) {
            int4 v = make_int4(0, 0, 0, 0);
            if (i0 + 4 * lane < bins) v = row[(i0 >> 2) + lane];
            if (!__any_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0)) continue;
            const uint32_t m0 = __ballot_sync(0xffffffffu, v.x != 0), m1 = __ballot_sync(0xffffffffu, v.y != 0);
            const uint32_t m2 = __ballot_sync(0xffffffffu, v.z != 0), m3 = __ballot_sync(0xffffffffu, v.w != 0);
            int p = n + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
            const uint32_t b = (uint32_t)(i0 + 4 * lane);
            if (v.x) { keys[base + p] = b;      counts[base + p] = (uint32_t)v.x; p++; }
            if (v.y) { keys[base + p] = b + 1u; counts[base + p] = (uint32_t)v.y; p++; }
            if (v.z) { keys[base + p] = b + 2u; counts[base + p] = (uint32_t)v.z; p++; }
            if (v.w) { keys[base + p] = b + 3u; counts[base + p] = (uint32_t)v.w; p++; }
            n += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        }
        if (lane == 0) row_count[r] = n;
    }
}

// Three small device -> pinned-host copies done by SMs (stores to mapped host memory) instead of the copy
// engine: the engine is busy for milliseconds at a time with the dense rows of the DMA share, and the
// pairs of the host share would queue behind them.
__global__ void __launch_bounds__(256) pairs_to_host_kernel(const uint32_t* __restrict__ s0, uint32_t* __restrict__ d0, int64_t n0,
                                                            const uint32_t* __restrict__ s1, uint32_t* __restrict__ d1, int64_t n1,
                                                            const uint32_t* __restrict__ s2, uint32_t* __restrict__ d2, int64_t n2)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t; i < n0; i += stride) d0[i] = s0[i];
    for (int64_t i = t; i < n1; i += stride) d1[i] = s1[i];
    for (int64_t i = t; i < n2; i += stride) d2[i] = s2[i];
}

cudaError_t launch_pairs_to_host(const uint32_t* s0, uint32_t* d0, int64_t n0, const uint32_t* s1, uint32_t* d1, int64_t n1,
                                 const uint32_t* s2, uint32_t* d2, int64_t n2, cudaStream_t st)
{
    const int64_t n = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    if (n <= 0) return cudaSuccess;
    const int64_t ctas = (n + 255) / 256;
    pairs_to_host_kernel<<<(unsigned)(ctas < 148 * 4 ? ctas : 148 * 4), 256, 0, st>>>(s0, d0, n0, s1, d1, n1, s2, d2, n2);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rows_to_pairs(const int32_t* rows, int64_t nrows, int bins, const int64_t* off, uint32_t* keys,
                                 uint32_t* counts, int32_t* row_count, cudaStream_t st)
{
    if (nrows <= 0) return cudaSuccess;
    const int64_t ctas = (nrows + 7) / 8;
    rows_to_pairs_kernel<<<(unsigned)(ctas < 148 * 8 ? ctas : 148 * 8), 256, 0, st>>>(rows, nrows, bins, off, keys, counts, row_count);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cfrk
