// dense_index.cu -- the k-mer index of every visited window of every read, as a flat uint32 list.
//
// Used by the host-buffer operator (api.cu, cfrk_count_dense_host = the reference's kmer_main contract):
// dense int32 rows are 4^k * 4 bytes per read and PCIe moves 55 GB/s, so for part of the reads only
// this list (4 bytes per window) crosses the bus and host threads expand it into the caller's rows
// with streaming stores, while the DMA engine carries the dense rows of the other reads.  What is
// computed is exactly ComputeIndex (reference src/kmer_kernel.cu:21-49): Index[start+t] for the
// positions t < vis that ComputeFreqNew visits (src/kmer_kernel.cu:83-88), -1 (0xFFFFFFFF) where the
// window holds a non-ACGT byte or the terminator -- without the float32 arithmetic and without the
// other nN - vis entries.
#include "kernels.h"
#include "kmer_device.cuh"
#include "stream_device.cuh"

namespace cfrk {

extern void count_launch();

constexpr int kIdxChunk = 512;   // windows per warp step
constexpr int kIdxBlocks = (15 + kIdxChunk + 30 + 15) / 16 + 3;
constexpr int kIdxWarps = 8;

template <int FMT>
__global__ void __launch_bounds__(kIdxWarps * 32) dense_index_kernel(const uint8_t* __restrict__ bases,
                                                                   const int64_t* __restrict__ start,
                                                                   const int32_t* __restrict__ length,
                                                                   const int64_t* __restrict__ ibeg,   // [r - r_begin]
                                                                   int64_t r_begin, int64_t r_end, int k, int mode,
                                                                   uint32_t* __restrict__ idx_out)
{
    __shared__ uint32_t s_cw[kIdxWarps][kIdxBlocks];
    __shared__ __align__(4) uint16_t s_vh[kIdxWarps][2 * ((kIdxBlocks + 1) / 2) + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStream st{s_cw[warp], s_vh[warp]};
    const int64_t nwarps = (int64_t)gridDim.x * kIdxWarps;
    for (int64_t r = r_begin + (int64_t)blockIdx.x * kIdxWarps + warp; r < r_end; r += nwarps) {
        const int64_t s = start[r];
        const int len = length[r];
        const int vis = mode == MODE_COMPAT ? min(len - 1, kRefBlockThreads) : len - k + 1;
        uint32_t* out = idx_out + ibeg[r - r_begin];
        for (int w0 = 0; w0 < vis; w0 += kIdxChunk) {
            const int64_t p0 = s + w0;                    // first base of the step
            const int a = (int)(p0 & 15);
            const int64_t blk0 = p0 >> 4;
            __syncwarp();
            for (int b = lane; b < kIdxBlocks; b += 32) {
                uint32_t c = 0, v = 0;
                const int64_t pos = (blk0 + b) * 16;      // buffer position of the block's first byte
                if (pos < s + len) {
                    encode16<FMT>(ld_block(bases + pos), c, v);
                    const int lo = (int)max((int64_t)0, min((int64_t)16, s - pos));
                    const int hi = (int)max((int64_t)0, min((int64_t)16, s + len - pos));
                    v &= from_pos(lo) & ~from_pos(hi);    // bases of this read only
                }
                st.cw[b] = c;
                st.vh[b ^ 1] = (uint16_t)v;
            }
            __syncwarp();
            uint32_t key[16];
            extract_windows<uint32_t, 16>(st, a + lane * 16, k, key);   // invalid windows: 0xFFFFFFFF
            const int t0 = w0 + lane * 16;
            if (t0 + 16 <= vis) {
                uint4* o4 = reinterpret_cast<uint4*>(out + t0);
                if ((reinterpret_cast<uintptr_t>(o4) & 15) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; q++) o4[q] = make_uint4(key[4 * q], key[4 * q + 1], key[4 * q + 2], key[4 * q + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e++) out[t0 + e] = key[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++)
                    if (t0 + e < vis) out[t0 + e] = key[e];
            }
        }
    }
}

// reads [r_begin, r_end): idx_out[ibeg[r - r_begin] + t] = index of the window starting at base t of read r,
// t < vis(r); ibeg is the caller's exclusive prefix sum of vis over the range.  k <= 12.
cudaError_t launch_dense_index(const void* bases, int fmt, const int64_t* start, const int32_t* length, const int64_t* ibeg,
                               int64_t r_begin, int64_t r_end, int k, int mode, uint32_t* idx_out, cudaStream_t st)
{
    const int64_t n = r_end - r_begin;
    if (n <= 0) return cudaSuccess;
    const int64_t ctas = (n + kIdxWarps - 1) / kIdxWarps;
    const unsigned grid = (unsigned)(ctas < 148 * 8 ? ctas : 148 * 8);
    const uint8_t* b8 = static_cast<const uint8_t*>(bases);
    if (fmt == FMT_ASCII)
        dense_index_kernel<FMT_ASCII><<<grid, kIdxWarps * 32, 0, st>>>(b8, start, length, ibeg, r_begin, r_end, k, mode, idx_out);
    else
        dense_index_kernel<FMT_CODES><<<grid, kIdxWarps * 32, 0, st>>>(b8, start, length, ibeg, r_begin, r_end, k, mode, idx_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cfrk
