// hist_split.cu -- whole-dataset k-mer histogram for 4^k bins that do not fit shared memory
// (k = 9..13), without one L2 reduction per window.  Default for k = 13 (the 256 MiB table no longer
// fits L2: 25 -> 147 Gbases/s); CFRK_HIST_SPLIT=1 forces it for k = 9..13, =0 disables it.  At k = 12
// (BASELINE.json config 5) the direct kernel stays the default: 199 vs 146 Gbases/s -- ncu shows this
// kernel waiting at its six barriers per tile (barrier stalls 9.7 per issue), not on memory.
//
// The direct kernel (global_hist_kernel, kernels.cu) issues one red.global per valid window into the
// 4^k-bin histogram; at k = 12 the 64 MiB table is L2-resident and the L2 reduction rate (~187 G/s)
// is the bound: 202 Gbases/s.  Here the k-mers are first SPLIT by their top 2k-15 bits into
// partitions of 32768 bins, each of which is then counted in shared memory:
//   0. extent_mask_kernel   (start, length) -> one validity bit per byte of the bases buffer, so that
//                           the buffer can be walked as ONE sequence: bytes between reads (FASTA
//                           headers, separators) are invalid, hence so is every window touching them.
//   1. split_scatter_kernel tiles of 8192 window starts: bases encoded once into a shared-memory
//                           2-bit stream, 16 consecutive windows per thread from registers
//                           (stream_device.cuh), tile-local counting sort by partition in shared
//                           memory (rank = returning shared atomic), ONE global atomic per partition
//                           and tile to reserve a run in the partition's buffer, runs written with
//                           consecutive threads on consecutive addresses: 15-bit suffixes (uint16).
//                           Partitions have a fixed capacity (1.25 x the mean + slack); a suffix that
//                           does not fit goes straight to the histogram with a red.global, so skewed
//                           data degrade towards the direct kernel instead of failing.
//   2. split_count_kernel   per partition: 32768 shared-memory counters (128 KiB), shared atomics
//                           over the partition's suffixes, added to the histogram at the end.
#include "kernels.h"
#include "kmer_device.cuh"
#include "stream_device.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace cfrk {

extern void count_launch();

constexpr int kSuffixBits = 15;
constexpr int kPartBins = 1 << kSuffixBits;       // bins per partition
constexpr int kSplitThreads = 512;
constexpr int kSplitW = 16;                        // consecutive windows per thread
constexpr int kSplitStep = kSplitThreads * kSplitW;
constexpr int kSplitBlocks = (15 + kSplitStep + 30 + 15) / 16 + 3;
constexpr int kMaxParts = 2048;                    // k <= 13
constexpr int kCountThreads = 1024;

// the bytes of every read become 1-bits (bit 15-j of mask[b] = byte j of block b).  One THREAD per read
// sets the first kMaskHead blocks (all of a short read: many reads in flight hide the two dependent
// loads); extent_mask_long_kernel, one warp per read, fills the rest of longer reads.
constexpr int kMaskHead = 64;

__device__ __forceinline__ void set_block_mask(uint16_t* __restrict__ mask, int64_t b, int64_t s, int len, int64_t b0, int64_t b1)
{
    const int lo = (int)max((int64_t)0, s - 16 * b);
    const int hi = (int)min((int64_t)16, s + len - 16 * b);
    const uint32_t m = from_pos(lo) & ~from_pos(hi);
    if (b == b0 || b == b1) {
        // a block at the end of a read may be shared with the neighbouring read
        atomicOr(reinterpret_cast<uint32_t*>(mask) + (b >> 1), m << ((b & 1) * 16));
    } else {
        mask[b] = (uint16_t)m;
    }
}

__global__ void __launch_bounds__(256) extent_mask_kernel(const int64_t* __restrict__ start, const int32_t* __restrict__ length,
                                                          int64_t nS, uint16_t* __restrict__ mask)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nS; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = start[r];
        const int len = length[r];
        if (len <= 0) continue;
        const int64_t b0 = s >> 4, b1 = (s + len - 1) >> 4;
        const int64_t bend = min(b1, b0 + kMaskHead - 1);
        for (int64_t b = b0; b <= bend; b++) set_block_mask(mask, b, s, len, b0, b1);
    }
}

__global__ void __launch_bounds__(256) extent_mask_long_kernel(const int64_t* __restrict__ start, const int32_t* __restrict__ length,
                                                               int64_t nS, uint16_t* __restrict__ mask)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < nS; r += nwarps) {
        const int len = length[r];
        if (len <= 16 * (kMaskHead - 1)) continue;   // at most kMaskHead blocks: done by its thread
        const int64_t s = start[r];
        const int64_t b0 = s >> 4, b1 = (s + len - 1) >> 4;
        for (int64_t b = b0 + kMaskHead + lane; b <= b1; b += 32) set_block_mask(mask, b, s, len, b0, b1);
    }
}

template <int FMT>
__global__ void __launch_bounds__(kSplitThreads, 2) split_scatter_kernel(const uint8_t* __restrict__ bases,
                                                                      const uint16_t* __restrict__ mask, int64_t nN, int k,
                                                                      int part_bits, uint32_t cap,
                                                                      uint32_t* __restrict__ cursor, uint16_t* __restrict__ buf,
                                                                      uint32_t* __restrict__ hist)
{
    constexpr int T = kSplitThreads, W = kSplitW, STEP = kSplitStep, NBLK = kSplitBlocks;
    constexpr int NB = (NBLK + T - 1) / T;       // raw blocks per thread and tile
    constexpr int PPT = kMaxParts / T;           // partitions per thread in the scan
    __shared__ uint32_t s_cw[NBLK];
    __shared__ __align__(4) uint16_t s_vh[2 * ((NBLK + 1) / 2) + 2];
    __shared__ uint32_t s_cnt[kMaxParts + 1], s_lstart[kMaxParts + 1], s_room[kMaxParts];
    __shared__ unsigned long long s_gaddr[kMaxParts];
    __shared__ uint32_t s_wsum[T / 32];
    __shared__ uint32_t s_total;
    extern __shared__ uint32_t s_staged[];       // STEP entries: the k-mer indices of the tile in partition order
    WarpStream st{s_cw, s_vh};
    const int np = 1 << part_bits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwin = nN - k + 1;
    const int64_t ntiles = (nwin + STEP - 1) / STEP;
    const int64_t blk_end = (nN + 15) >> 4;
    for (int p = threadIdx.x; p <= kMaxParts; p += T) s_cnt[p] = 0u;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t blk0 = t * (STEP / 16);   // tiles start at multiples of 16 bases: stream offset 0
        __syncthreads();   // the previous tile is done with the stream and the staging buffer
#pragma unroll
        for (int q = 0; q < NB; q++) {
            const int b = threadIdx.x + q * T;
            if (b < NBLK) {
                uint32_t c = 0, v = 0;
                if (blk0 + b < blk_end) {
                    encode16<FMT>(ld_block(bases + (blk0 + b) * 16), c, v);
                    v &= mask[blk0 + b];
                }
                st.cw[b] = c;
                st.vh[b ^ 1] = (uint16_t)v;
            }
        }
        __syncthreads();
        // invalid windows go to the dummy partition np: no branches around the shared atomics, and they
        // end up behind the valid entries of the staging buffer
        uint32_t key[W], rank[W];
        const uint32_t valid = extract_windows<uint32_t, W>(st, (int)threadIdx.x * W, k, key);
#pragma unroll
        for (int e = 0; e < W; e++) {
            if (!(valid >> e & 1u)) key[e] = (uint32_t)np << kSuffixBits;
            rank[e] = atomicAdd(&s_cnt[key[e] >> kSuffixBits], 1u);
        }
        __syncthreads();
        // exclusive scan of the tile's partition counts; one reservation per non-empty partition
        {
            uint32_t c[PPT], sum = 0;
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                const int p = threadIdx.x * PPT + i;
                c[i] = p < np ? s_cnt[p] : 0u;
                sum += c[i];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < T / 32; w++) before += w < warp ? s_wsum[w] : 0u;
            uint32_t off = before + inc - sum;
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                const int p = threadIdx.x * PPT + i;
                if (p < np) {
                    s_lstart[p] = off;
                    if (c[i]) {
                        // entry i of the staging buffer (off <= i < off + c) -> buf[p * cap + g + i - off], while
                        // g + i - off < cap, i.e. i < off + cap - g
                        const uint32_t g = atomicAdd(&cursor[p], c[i]);
                        s_gaddr[p] = (unsigned long long)p * cap + g - off;
                        s_room[p] = g < cap ? off + (cap - g) : off;
                    }
                    s_cnt[p] = 0u;   // for the next tile
                    off += c[i];
                }
            }
            if (threadIdx.x == T - 1) { s_total = before + inc; s_lstart[np] = before + inc; s_cnt[np] = 0u; }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < W; e++) s_staged[s_lstart[key[e] >> kSuffixBits] + rank[e]] = key[e];
        __syncthreads();
        const uint32_t total = s_total;
        for (uint32_t i = threadIdx.x; i < total; i += T) {
            const uint32_t v = s_staged[i];
            const uint32_t d = v >> kSuffixBits;
            if (i < s_room[d]) buf[s_gaddr[d] + i] = (uint16_t)(v & (kPartBins - 1));
            else atomicAdd(&hist[v], 1u);   // partition full: count it directly
        }
    }
}

// partition p is counted by `cpp` CTAs (shares of its suffixes), each in its own shared-memory counters
__global__ void __launch_bounds__(kCountThreads) split_count_kernel(const uint32_t* __restrict__ cursor, uint32_t cap,
                                                                    const uint16_t* __restrict__ buf, int cpp,
                                                                    uint32_t* __restrict__ hist)
{
    extern __shared__ uint32_t s_hist[];   // kPartBins counters
    const int p = blockIdx.x / cpp, sub = blockIdx.x % cpp;
    const uint32_t n = min(cursor[p], cap);
    // shares in units of 8 suffixes (one 16-byte load); cap is a multiple of 8
    const uint32_t units = (n + 7) / 8;
    const uint32_t u0 = (uint32_t)((uint64_t)units * sub / cpp), u1 = (uint32_t)((uint64_t)units * (sub + 1) / cpp);
    if (u0 >= u1) return;
    for (int i = threadIdx.x; i < kPartBins; i += kCountThreads) s_hist[i] = 0u;
    __syncthreads();
    const uint4* in = reinterpret_cast<const uint4*>(buf + (size_t)p * cap);
    for (uint32_t u = u0 + threadIdx.x; u < u1; u += kCountThreads) {
        const uint4 v = in[u];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t live = min(8u, n - u * 8);   // suffixes of this unit that exist
#pragma unroll
        for (int j = 0; j < 8; j++)
            if ((uint32_t)j < live) atomicAdd(&s_hist[(w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu], 1u);
    }
    __syncthreads();
    uint32_t* out = hist + ((size_t)p << kSuffixBits);
    for (int i = threadIdx.x; i < kPartBins; i += kCountThreads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&out[i], c);
    }
}

bool hist_split_applies(int k, int64_t nN)
{
    const char* ev = getenv("CFRK_HIST_SPLIT");   // 0 off, 1 always, default by size
    const int mode = ev ? atoi(ev) : -1;
    if (k < 9 || k > 13 || nN < k) return false;
    if (mode == 0) return false;
    if (mode == 1) return true;
    // measured on 1 Gbase (tools/bench_hist.py): k = 13 direct 25 Gbases/s (the 256 MiB table misses L2) vs
    // 147 here; k = 12 direct 199 vs 146 here, k = 10 207 vs 214: the split pays only once the table
    // leaves L2
    return k == 13 && nN >= ((int64_t)32 << 20);
}

cudaError_t launch_global_hist_split(const void* bases, int fmt, const int64_t* start, const int32_t* length, int64_t nN,
                                     int64_t nS, int k, uint32_t* hist, cudaStream_t st)
{
    cudaError_t e;
    int dev = 0, num_sms = 148;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    keep_pool_memory(dev);
    const int part_bits = 2 * k - kSuffixBits;
    const int np = 1 << part_bits;
    const int64_t nwin = nN - k + 1;
    const int64_t nblocks = (nN + 15) / 16;
    // capacity per partition: 1.25 x the mean + slack, a multiple of 8 (16-byte loads); what does not fit
    // is counted directly
    int64_t cap64 = (nwin / np) + (nwin / np) / 4 + 8192;
    cap64 = (cap64 + 7) & ~(int64_t)7;
    if (cap64 > 0xFFFFFFF0ll) return cudaErrorInvalidValue;
    const uint32_t cap = (uint32_t)cap64;

    uint16_t *mask = nullptr, *buf = nullptr;
    uint32_t* cursor = nullptr;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&mask), (size_t)(nblocks + 2) * 2, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&cursor), (size_t)np * 4, st)) != cudaSuccess) { cudaFreeAsync(mask, st); return e; }
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&buf), (size_t)np * cap * 2, st)) != cudaSuccess) {
        cudaFreeAsync(mask, st); cudaFreeAsync(cursor, st);
        return e;
    }
    const bool trace = getenv("CFRK_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cfrk hist] %9.3f ms  %s\n", std::chrono::duration<double, std::milli>(now - t_last).count(), what);
        t_last = now;
    };
    mark("scratch");
    cudaMemsetAsync(mask, 0, (size_t)(nblocks + 2) * 2, st);
    cudaMemsetAsync(cursor, 0, (size_t)np * 4, st);
    {
        const int64_t ctas = (nS + 255) / 256, wctas = (nS + 7) / 8;
        extent_mask_kernel<<<(unsigned)(ctas < (int64_t)num_sms * 8 ? ctas : (int64_t)num_sms * 8), 256, 0, st>>>(start, length, nS, mask);
        extent_mask_long_kernel<<<(unsigned)(wctas < (int64_t)num_sms * 8 ? wctas : (int64_t)num_sms * 8), 256, 0, st>>>(start, length, nS, mask);
        count_launch(); count_launch();
    }
    mark("extent mask");
    {
        const size_t dyn = (size_t)kSplitStep * 4;
        const int64_t ntiles = (nwin + kSplitStep - 1) / kSplitStep;
        const unsigned grid = (unsigned)(ntiles < (int64_t)num_sms * 2 ? ntiles : (int64_t)num_sms * 2);
        const uint8_t* b8 = static_cast<const uint8_t*>(bases);
        if (fmt == FMT_ASCII) {
            cudaFuncSetAttribute(split_scatter_kernel<FMT_ASCII>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            split_scatter_kernel<FMT_ASCII><<<grid, kSplitThreads, dyn, st>>>(b8, mask, nN, k, part_bits, cap, cursor, buf, hist);
        } else {
            cudaFuncSetAttribute(split_scatter_kernel<FMT_CODES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            split_scatter_kernel<FMT_CODES><<<grid, kSplitThreads, dyn, st>>>(b8, mask, nN, k, part_bits, cap, cursor, buf, hist);
        }
        count_launch();
    }
    mark("split + scatter");
    {
        int cpp = 1;
        while (np * cpp < num_sms * 4) cpp *= 2;
        const size_t dyn = (size_t)kPartBins * 4;
        cudaFuncSetAttribute(split_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        split_count_kernel<<<(unsigned)(np * cpp), kCountThreads, dyn, st>>>(cursor, cap, buf, cpp, hist);
        count_launch();
    }
    mark("count partitions");
    e = cudaGetLastError();
    cudaFreeAsync(buf, st);
    cudaFreeAsync(cursor, st);
    cudaFreeAsync(mask, st);
    return e;
}

}  // namespace cfrk
