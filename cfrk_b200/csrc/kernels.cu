// kernels.cu -- sm_100a kernels of the CFRK hot path and their launchers.
//
// Per-read dense int32 rows (replace SetMatrix x2 + ComputeIndex + ComputeFreqNew, reference
// src/kmer_kernel.cu:6-90 as launched by src/kmer_main.cu:107-111), one kernel family per row size:
//   dense_count_kernel   k = 1..3  CTA-cooperative 16 KiB tiles, TMA bulk store
//   dense_warp_kernel    k = 4..6  warp-autonomous tiles, read-clear-store, spill hand-off
//   dense_bigrow_kernel  k = 7, 8  TMA zero stream + L2 reductions on L2-resident tiles
//   dense_row_kernel     k = 7     measured alternative (one shared-memory row per CTA)
//   spill_fixup_kernel             data-dependent part of the cross-tile spill (compat mode)
// Other stages:
//   global_hist_kernel   whole-dataset histogram, k <= 15    (no reference counterpart; config C5)
//   encode_2bit_kernel   bases -> packed 2-bit + validity    (replaces src/fastaIO.h:123-139)
// Which kernel serves which k, and every tile size, is a measured choice: profiles/r1_notes.md.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
#include "kernels.h"
#include "kmer_device.cuh"
#include "dense_args.h"

#include "internal.h"

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace cfrk {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Small device scratch that survives between launches: one buffer per (device, stream), grown on
// demand, process-wide.  Launches of one stream are ordered, so they can share it; different
// streams get their own.  (cudaMallocAsync per launch stalled the host for about a millisecond per
// launch -- visible at k=4, where a whole sweep takes 2.6 ms.)  cfrk_release() frees them.
namespace {
struct ScratchSlot { int dev; cudaStream_t st; void* p; size_t cap; };
std::mutex g_scratch_mu;
std::vector<ScratchSlot> g_scratch;
}

cudaError_t stream_scratch(cudaStream_t st, size_t bytes, void** out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    for (ScratchSlot& s : g_scratch) {
        if (s.dev == dev && s.st == st) {
            if (s.cap < bytes) {
                if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;   // last user may still run
                cudaFree(s.p);
                s.p = nullptr; s.cap = 0;
                if ((e = cudaMalloc(&s.p, bytes + bytes / 4)) != cudaSuccess) return e;
                s.cap = bytes + bytes / 4;
            }
            *out = s.p;
            return cudaSuccess;
        }
    }
    ScratchSlot s{dev, st, nullptr, bytes + bytes / 4 + 256};
    if ((e = cudaMalloc(&s.p, s.cap)) != cudaSuccess) return e;
    g_scratch.push_back(s);
    *out = s.p;
    return cudaSuccess;
}

void release_stream_scratch()
{
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    for (ScratchSlot& s : g_scratch) {
        cudaSetDevice(s.dev);
        cudaFree(s.p);   // implicit device synchronisation: no launch still uses it
    }
    g_scratch.clear();
    cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Tile geometry of the CTA-cooperative kernel.  A tile is TILE_BINS consecutive int32 of the
// output, i.e. a contiguous TILE_BINS*4-byte span of HBM that one CTA builds in shared memory and
// ships with one TMA bulk store: RPT whole rows (reads) per tile.  (SUB > 1, rows larger than a
// tile, is the big-row kernel's business.)
template <int K, int TILE_BINS_T>
struct Geo {
    static constexpr int BINS = 1 << (2 * K);
    static constexpr int SUB = BINS > TILE_BINS_T ? BINS / TILE_BINS_T : 1;
    static constexpr int RPT_RAW = BINS >= TILE_BINS_T ? 1 : TILE_BINS_T / BINS;
    static constexpr int RPT = RPT_RAW > kMaxGroupReads ? kMaxGroupReads : RPT_RAW;
    static constexpr int TILE_BINS = SUB > 1 ? TILE_BINS_T : RPT * BINS;
    static constexpr int TILE_BYTES = TILE_BINS * 4;
    static constexpr int TABLE_READS = RPT + 1;  // + halo read (compat spill)
    static constexpr int ROW_ALIGN = BINS * 4 < 16 ? 16 : BINS * 4;  // rows aligned to their size
    // shared memory: alignment slack, NBUF tile buffers, then the read table
    static constexpr int table_bytes() { return TABLE_READS * 16 + (TABLE_READS + 1) * 4 + 16; }
};

template <int K, int TILE_BINS_T>
struct DenseSink {
    using G = Geo<K, TILE_BINS_T>;
    static constexpr bool kCtaUniform = false;
    static constexpr bool kSharedRows = true;
    static constexpr bool kRowsAligned = true;
    uint32_t hist_saddr;  // shared-window address of the tile buffer (aligned to the row size)
    int qb, period;       // table-local reads qb, qb+period, ... open a reference chunk (period 0: none)
    __device__ __forceinline__ uint32_t row_saddr(int q) const { return hist_saddr + (uint32_t)q * (G::BINS * 4); }
    // the reference adds at Freq[4^k*i + (-1)]: last bin of read i-1 (src/kmer_kernel.cu:84-87);
    // q == 0 belongs to the previous tile (counted there through its halo read); for the first
    // read of a reference chunk (a separate kmer_main call) it is the lost Freq[-1] store.
    __device__ __forceinline__ void invalid(int q, int in_read, int extra)
    {
        if (period > 0 && q >= qb && (q - qb) % period == 0) return;
        if (q >= 1)
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(row_saddr(q) - 4u), "r"((uint32_t)(in_read + extra)) : "memory");
    }
};


template <int K, int FMT, int TILE_BINS_T, int NTHREADS, int NBUF>
__global__ void __launch_bounds__(NTHREADS) dense_count_kernel(const DenseArgs a)
{
    using G = Geo<K, TILE_BINS_T>;
    static_assert(G::SUB == 1, "rows larger than a tile go through dense_bigrow_kernel");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // align the tile buffers to the row size in the shared window (emit_item ORs bin offsets in)
    const uint32_t raw_saddr = (uint32_t)__cvta_generic_to_shared(smem_raw);
    unsigned char* smem = smem_raw + ((G::ROW_ALIGN - (raw_saddr & (G::ROW_ALIGN - 1))) & (G::ROW_ALIGN - 1));
    uint32_t* bufs = reinterpret_cast<uint32_t*>(smem);
    unsigned char* tp = smem + (size_t)NBUF * G::TILE_BYTES;
    ReadTable tb;
    tb.start = reinterpret_cast<int64_t*>(tp);
    tb.tend = reinterpret_cast<int32_t*>(tp + G::TABLE_READS * 8);
    tb.extra = reinterpret_cast<int32_t*>(tp + G::TABLE_READS * 12);
    tb.cum = reinterpret_cast<uint32_t*>(tp + G::TABLE_READS * 16);

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const int64_t group = tile / G::SUB;
        const int sub = (int)(tile - group * G::SUB);
        const int64_t r0 = a.read_begin + group * G::RPT;
        const int nrows = (int)min((int64_t)G::RPT, a.read_end - r0);
        const bool has_last = (sub == G::SUB - 1);
        // table-local reads qb, qb+period, ... open a reference chunk: their spill is dropped
        int qb = 0, period = 0;
        if (a.chunk_size > 0) {
            const int64_t phase = (a.index_base + r0) % a.chunk_size;
            const int64_t first = phase == 0 ? 0 : a.chunk_size - phase;
            if (first <= G::TABLE_READS) {
                qb = (int)first;
                period = (int)min(a.chunk_size, (int64_t)(4 * kMaxGroupReads));
            }
        } else if (r0 == 0) {
            period = 4 * kMaxGroupReads;  // only read 0
        }
        const bool next_opens_chunk = period > 0 && nrows >= qb && (nrows - qb) % period == 0;
        const bool halo = a.mode == MODE_COMPAT && has_last && (r0 + nrows < a.nS) && !next_opens_chunk;
        const int nreads = nrows + (halo ? 1 : 0);
        uint32_t* hist = bufs + (size_t)(it % NBUF) * G::TILE_BINS;

        // the TMA store that last used this buffer (NBUF tiles ago) must have read it out
        if (threadIdx.x == 0) bulk_wait_read<NBUF - 1>();
        fill_read_table<K>(tb, a.start, a.length, r0, nreads, a.mode, a.nN, a.nS, a.chunk_size, a.index_base);
        __syncthreads();

        {   // zero the tile (replaces SetMatrix(d_Freq, 0), src/kmer_main.cu:108) ...
            uint4* h4 = reinterpret_cast<uint4*>(hist);
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
            for (int i = threadIdx.x; i < G::TILE_BYTES / 16; i += NTHREADS) h4[i] = z;
            // ... while warp 0 turns block counts into item offsets
            if (threadIdx.x < 32) scan_read_table(tb, nreads);
        }
        __syncthreads();

        DenseSink<K, TILE_BINS_T> sink{(uint32_t)__cvta_generic_to_shared(hist), qb, period};
        for_each_window<K, FMT, G::TABLE_READS>(BasesRef{a.bases, a.valid}, tb, nreads, nrows, a.mode, sink);

        fence_async_proxy_shared();  // make the shared-memory counts visible to the TMA engine
        __syncthreads();
        if (threadIdx.x == 0) {
            const int64_t row0 = (r0 - a.read_begin);
            uint32_t* dst = a.out + row0 * G::BINS + (int64_t)sub * G::TILE_BINS;
            const uint32_t bytes = G::SUB > 1 ? (uint32_t)G::TILE_BYTES : (uint32_t)nrows * G::BINS * 4u;
            bulk_store_tile(dst, hist, bytes);
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
// Small rows (k <= 5), warp-autonomous: every warp builds its own tiles of RW rows in its own two
// shared-memory buffers and ships them with its own TMA bulk stores.  No CTA barrier, no
// shared-memory read table (lane q holds read q, lookups are shuffles), the next tile's offsets
// and the next chunk's bases are already in flight while the current ones are processed.  This
// replaces the CTA-cooperative tile loop for small k, which ncu showed to be one serial chain per
// 16 KiB (43-77 % of warp samples waiting at barriers, profiles/r1_notes.md).
template <int K>
struct WarpSink {
    static constexpr bool kCtaUniform = false;
    static constexpr bool kSharedRows = true;
    static constexpr bool kRowsAligned = true;
    static constexpr int BINS = 1 << (2 * K);
    uint32_t hist_saddr;
    int qb, period;
    int carry0;   // per lane: data-dependent invalid windows of the tile's FIRST read (owed to the previous tile)
    __device__ __forceinline__ uint32_t row_saddr(int q) const { return hist_saddr + (uint32_t)q * (BINS * 4); }
    __device__ __forceinline__ void invalid(int q, int in_read, int extra)
    {
        if (period > 0 && q >= qb && (q - qb) % period == 0) return;
        if (q >= 1)
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(row_saddr(q) - 4u), "r"((uint32_t)(in_read + extra)) : "memory");
        else
            carry0 += in_read;   // its `extra` was added by the previous tile from the read length alone
    }
};

// DIRECT = false: two buffers per warp, rows leave through the warp's own TMA bulk stores.
// DIRECT = true : one buffer per warp; the warp reads its finished rows out of shared memory,
//                 clears them in the same pass and writes them with coalesced 16-byte streaming
//                 stores.  Half the shared memory per warp -> twice the resident warps, and no wait
//                 for the TMA engine before the buffer can be reused.
template <int K, int FMT, int RW, int WARPS, bool DIRECT, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB) dense_warp_kernel(const DenseArgs a)
{
    constexpr int BINS = 1 << (2 * K);
    constexpr int TILE_BINS = RW * BINS;
    constexpr int TILE_BYTES = TILE_BINS * 4;
    constexpr int NBUF = DIRECT ? 1 : 2;
    static_assert(RW + 1 <= 32, "one lane per read of the tile (+ halo)");
    constexpr int ROW_ALIGN = BINS * 4 < 16 ? 16 : BINS * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t raw_saddr = (uint32_t)__cvta_generic_to_shared(smem_raw);
    unsigned char* smem = smem_raw + ((ROW_ALIGN - (raw_saddr & (ROW_ALIGN - 1))) & (ROW_ALIGN - 1));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* bufs = reinterpret_cast<uint32_t*>(smem) + (size_t)warp * NBUF * TILE_BINS;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    int64_t tile = (int64_t)blockIdx.x * WARPS + warp;
    if (tile >= a.num_tiles) return;
    if (DIRECT) {
        uint4* h4 = reinterpret_cast<uint4*>(bufs);
#pragma unroll
        for (int i = lane; i < TILE_BYTES / 16; i += 32) h4[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
    }

    // offsets of the first tile; afterwards always one tile ahead
    int64_t s = 0; int len = 0;
    {
        const int64_t r = a.read_begin + tile * RW + lane;
        if (lane <= RW && r < a.nS) { s = a.start[r]; len = a.length[r]; }
    }
    for (int it = 0; tile < a.num_tiles; ++it) {
        const int64_t r0 = a.read_begin + tile * RW;
        const int nrows = (int)min((int64_t)RW, a.read_end - r0);
        int qb = 0, period = 0;
        if (a.chunk_size > 0) {
            const int64_t phase = (a.index_base + r0) % a.chunk_size;
            const int64_t first = phase == 0 ? 0 : a.chunk_size - phase;
            if (first <= RW + 1) {
                qb = (int)first;
                period = (int)min(a.chunk_size, (int64_t)(4 * kMaxGroupReads));
            }
        } else if (r0 == 0) {
            period = 4 * kMaxGroupReads;
        }
        const bool next_opens_chunk = period > 0 && nrows >= qb && (nrows - qb) % period == 0;
        const bool halo = a.mode == MODE_COMPAT && (r0 + nrows < a.nS) && !next_opens_chunk;
        // The spill of the read after the tile (read r0+nrows) into the tile's last row:
        //   * its length-only part (`extra`) is added here from length[] alone;
        //   * its data-dependent part (windows holding a non-ACGT byte; zero for clean reads) is
        //     recorded by the tile that owns that read in a.handoff[this tile] and added by
        //     spill_fixup_kernel after this kernel -- the kernel boundary is the only ordering
        //     needed (a fence + atomic handshake inside the kernel cost 3x the run time at k=4).
        // Without a.handoff, and for the last tile of a launch (the owner is another launch), the
        // read is scanned here as a halo read instead.
        const bool scan_halo = halo && (a.handoff == nullptr || tile == a.num_tiles - 1);
        const int nreads = nrows + (scan_halo ? 1 : 0);
        const ChunkScope cs{a.start, a.nS, a.nN, a.chunk_size, a.index_base};
        const LaneRead lr = make_lane_read<K>(lane < nrows + (halo ? 1 : 0), lane < nreads, s, len, a.mode,
                                              compat_avail(cs, r0 + lane, s, len, a.mode));

        const int64_t next_tile = tile + nwarps;
        if (next_tile < a.num_tiles) {  // prefetch the next tile's offsets
            const int64_t r = a.read_begin + next_tile * RW + lane;
            if (lane <= RW && r < a.nS) { s = a.start[r]; len = a.length[r]; }
        }

        uint32_t* hist = bufs + (DIRECT ? 0 : (size_t)(it & 1) * TILE_BINS);
        if (!DIRECT) {
            if (lane == 0) bulk_wait_read<1>();  // the store issued two tiles ago has read this buffer
            __syncwarp();
            uint4* h4 = reinterpret_cast<uint4*>(hist);
#pragma unroll
            for (int i = lane; i < TILE_BYTES / 16; i += 32) h4[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
        }
        WarpSink<K> sink{(uint32_t)__cvta_generic_to_shared(hist), qb, period, 0};
        warp_for_each_window<K, FMT, RW + 1>(BasesRef{a.bases, a.valid}, lr, nreads, nrows, a.mode, sink);
        uint32_t* dst = a.out + (r0 - a.read_begin) * BINS;
        if (halo && !scan_halo) {   // length-only part of the next read's spill
            const int ex = __shfl_sync(0xffffffffu, lr.extra, nrows);
            __syncwarp();
            if (lane == 0 && ex > 0) hist[nrows * BINS - 1] += (uint32_t)ex;
        }
        if (DIRECT) {
            __syncwarp();
            uint4* h4 = reinterpret_cast<uint4*>(hist);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            const int n16 = nrows * (BINS / 4);
#pragma unroll 4
            for (int i = lane; i < n16; i += 32) {
                const uint4 v = h4[i];
                h4[i] = make_uint4(0u, 0u, 0u, 0u);
                asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
                             :: "l"(d4 + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            __syncwarp();
        } else {
            fence_async_proxy_shared();
            __syncwarp();
            if (lane == 0) bulk_store_tile(dst, hist, (uint32_t)nrows * BINS * 4u);
        }
        if (DIRECT && a.handoff != nullptr && a.mode == MODE_COMPAT) {
            int c0 = sink.carry0;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) c0 += __shfl_xor_sync(0xffffffffu, c0, d);
            const bool q0_opens_chunk = period > 0 && qb == 0;
            if (lane == 0 && c0 > 0 && tile != 0 && !q0_opens_chunk) a.handoff[tile - 1] = (uint32_t)c0;
        }
        tile = next_tile;
    }
    if (!DIRECT && lane == 0) bulk_wait_all();
}

// handoff[t] != 0: the first read of tile t+1 had that many invalid windows -> last bin of the
// last row of tile t (rows_per_tile * bins int32 per tile)
__global__ void spill_fixup_kernel(const uint32_t* __restrict__ handoff, int64_t ntiles, uint32_t* __restrict__ out,
                                   int64_t tile_bins)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles - 1; t += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = handoff[t];
        if (v) out[(t + 1) * tile_bins - 1] += v;
    }
}

cudaError_t launch_spill_fixup(const uint32_t* handoff, int64_t ntiles, uint32_t* out, int64_t tile_bins, cudaStream_t st)
{
    const int64_t blocks = (ntiles + 255) / 256;
    spill_fixup_kernel<<<(unsigned)(blocks < 2368 ? blocks : 2368), 256, 0, st>>>(handoff, ntiles, out, tile_bins);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <int K, int FMT, int RW, int WARPS, bool DIRECT = false, int MINB = 1>
static cudaError_t launch_warp_k(const DenseArgs& a0, cudaStream_t st)
{
    auto kern = dense_warp_kernel<K, FMT, RW, WARPS, DIRECT, MINB>;
    constexpr int smem = WARPS * (DIRECT ? 1 : 2) * RW * (1 << (2 * K)) * 4 + ((1 << (2 * K)) * 4 < 16 ? 16 : (1 << (2 * K)) * 4);
    static thread_local int configured_dev = -1;
    static thread_local int ctas_per_sm = 0, num_sms = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured_dev = dev;
    }
    DenseArgs a = a0;
    a.num_tiles = (a.read_end - a.read_begin + RW - 1) / RW;
    if (a.num_tiles <= 0) return cudaSuccess;
    const int64_t ctas_needed = (a.num_tiles + WARPS - 1) / WARPS;
    const int64_t resident = (int64_t)num_sms * ctas_per_sm;
    const unsigned grid = (unsigned)(ctas_needed < resident ? ctas_needed : resident);
    static const int use_handoff = env_int("CFRK_HANDOFF", 1);
    if (DIRECT && use_handoff && a.mode == MODE_COMPAT && a.num_tiles > 1) {
        // one word per tile boundary, zeroed for this launch
        if ((e = stream_scratch(st, (size_t)a.num_tiles * 4, reinterpret_cast<void**>(&a.handoff))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(a.handoff, 0, (size_t)a.num_tiles * 4, st)) != cudaSuccess) return e;
    }
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (a.handoff) {
        const int64_t blocks = (a.num_tiles + 255) / 256;
        spill_fixup_kernel<<<(unsigned)(blocks < 2368 ? blocks : 2368), 256, 0, st>>>(
            a.handoff, a.num_tiles, a.out, (int64_t)RW * (1 << (2 * K)));
        g_launches.fetch_add(1);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    return e;
}

// ------------------------------------------------------------------------------------------
// One row per CTA (k = 7: 64 KiB rows): the CTA counts one read into a shared-memory row, then all
// 256 threads read the row out, clear it in the same pass and write it with 16-byte streaming
// stores.  Same idea as the DIRECT warp tiles, at CTA granularity: no TMA, no L2 reductions, no
// dependence on the row staying L2-resident.  The spill of the next read is handled like there:
// length-only part from length[], data part through the hand-off words + spill_fixup_kernel.
template <int K>
struct RowSink {
    static constexpr bool kCtaUniform = false;
    static constexpr bool kSharedRows = true;
    static constexpr bool kRowsAligned = false;
    static constexpr int BINS = 1 << (2 * K);
    uint32_t hist_saddr;
    bool drop0;     // read 0 of the table opens a chunk / the launch: its spill is dropped
    bool scan;      // read 1 of the table is a halo read scanned here (last row of a launch)
    int carry0;
    __device__ __forceinline__ uint32_t row_saddr(int) const { return hist_saddr; }
    __device__ __forceinline__ void invalid(int q, int in_read, int extra)
    {
        if (q == 0) { carry0 += in_read; return; }
        // q == 1: halo read scanned here -> last bin of this row
        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hist_saddr + (uint32_t)(BINS * 4 - 4)), "r"((uint32_t)(in_read + extra)) : "memory");
    }
};

constexpr int kRowThreads = 256;

template <int K, int FMT>
__global__ void __launch_bounds__(kRowThreads) dense_row_kernel(const DenseArgs a)
{
    constexpr int BINS = 1 << (2 * K);
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem);
    unsigned char* tp = smem + (size_t)BINS * 4;
    ReadTable tb;
    tb.start = reinterpret_cast<int64_t*>(tp);
    tb.tend = reinterpret_cast<int32_t*>(tp + 2 * 8);
    tb.extra = reinterpret_cast<int32_t*>(tp + 2 * 12);
    tb.cum = reinterpret_cast<uint32_t*>(tp + 2 * 16);
    __shared__ int s_carry;

    for (int i = threadIdx.x; i < BINS / 4; i += kRowThreads) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0u, 0u, 0u, 0u);
    const int64_t nrows_total = a.read_end - a.read_begin;
    for (int64_t row = blockIdx.x; row < nrows_total; row += gridDim.x) {
        const int64_t r = a.read_begin + row;
        const bool compat = a.mode == MODE_COMPAT;
        const bool opens = a.chunk_size > 0 ? ((a.index_base + r) % a.chunk_size == 0) : (r == 0);
        const bool next_opens = a.chunk_size > 0 ? ((a.index_base + r + 1) % a.chunk_size == 0) : false;
        const bool halo = compat && (r + 1 < a.nS) && !next_opens;
        const bool scan_halo = halo && (a.handoff == nullptr || row == nrows_total - 1);
        const int nreads = 1 + (scan_halo ? 1 : 0);
        if (threadIdx.x == 0) s_carry = 0;
        fill_read_table<K>(tb, a.start, a.length, r, 1 + (halo ? 1 : 0), a.mode, a.nN, a.nS, a.chunk_size, a.index_base);
        __syncthreads();
        if (threadIdx.x == 0) {   // 2-entry "scan"
            const uint32_t n0 = tb.cum[0], n1 = scan_halo ? tb.cum[1] : 0u;
            tb.cum[0] = 0u; tb.cum[1] = n0; tb.cum[2] = n0 + n1;
        }
        __syncthreads();
        RowSink<K> sink{(uint32_t)__cvta_generic_to_shared(hist), opens, scan_halo, 0};
        for_each_window<K, FMT, 2>(BasesRef{a.bases, a.valid}, tb, nreads, 1, a.mode, sink);
        if (sink.carry0) atomicAdd(&s_carry, sink.carry0);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (halo && !scan_halo && tb.extra[1] > 0) hist[BINS - 1] += (uint32_t)tb.extra[1];
            if (compat && a.handoff != nullptr && s_carry > 0 && row != 0 && !opens) a.handoff[row - 1] = (uint32_t)s_carry;
        }
        __syncthreads();
        uint4* h4 = reinterpret_cast<uint4*>(hist);
        uint4* d4 = reinterpret_cast<uint4*>(a.out + row * BINS);
#pragma unroll 4
        for (int i = threadIdx.x; i < BINS / 4; i += kRowThreads) {
            const uint4 v = h4[i];
            h4[i] = make_uint4(0u, 0u, 0u, 0u);
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(d4 + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Big rows (k >= 6: 16 / 64 / 256 KiB per read, of which a read touches <= 1024+1 bins).
// The row write IS the roofline, so the kernel is a streaming zero-writer with sparse patches:
//   1. one thread streams the rows of a group of reads as zeros with TMA bulk stores whose source
//      is a CONSTANT zero buffer in shared memory (never modified: any number of stores in flight,
//      nothing to re-zero, no buffer hand-over),
//   2. meanwhile all warps load + encode the group's bases,
//   3. once the zero stores have completed, the counts are scattered with red.global.add: the
//      rows were written microseconds ago and are still L2-resident, so the reductions merge in
//      L2 and DRAM sees each row exactly once.
// L2 residency is what makes this work, so the bytes in flight (resident CTAs x tile) are bounded
// far below the 126 MB L2: measured on B200 (profiles/r1_notes.md) 888 CTAs x 256 KiB = 227 MB gave
// 3 % L2 hits for the reductions and 0.65-0.79 of the HBM roofline; 444 x 64 KiB = 28 MB gives 0.91
// (k=7) and 1.06 (k=8) of the measured copy bandwidth.
template <int K, int FMT>
static cudaError_t launch_row_k(const DenseArgs& a0, cudaStream_t st)
{
    auto kern = dense_row_kernel<K, FMT>;
    constexpr int smem = (1 << (2 * K)) * 4 + 2 * 16 + 3 * 4 + 16;
    static thread_local int configured_dev = -1;
    static thread_local int ctas_per_sm = 0, num_sms = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kRowThreads, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured_dev = dev;
    }
    DenseArgs a = a0;
    a.num_tiles = a.read_end - a.read_begin;
    if (a.num_tiles <= 0) return cudaSuccess;
    const int64_t resident = (int64_t)num_sms * ctas_per_sm;
    const unsigned grid = (unsigned)(a.num_tiles < resident ? a.num_tiles : resident);
    if (a.mode == MODE_COMPAT && a.num_tiles > 1) {
        if ((e = stream_scratch(st, (size_t)a.num_tiles * 4, reinterpret_cast<void**>(&a.handoff))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(a.handoff, 0, (size_t)a.num_tiles * 4, st)) != cudaSuccess) return e;
    }
    kern<<<grid, kRowThreads, smem, st>>>(a);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (a.handoff) {
        const int64_t blocks = (a.num_tiles + 255) / 256;
        spill_fixup_kernel<<<(unsigned)(blocks < 2368 ? blocks : 2368), 256, 0, st>>>(a.handoff, a.num_tiles, a.out,
                                                                                        (int64_t)(1 << (2 * K)));
        g_launches.fetch_add(1);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    return e;
}

template <int K, int TILE_BYTES>
struct BigGeo {
    static constexpr int BINS = 1 << (2 * K);
    static constexpr int ROW_BYTES = BINS * 4;
    static constexpr int SUB = ROW_BYTES > TILE_BYTES ? ROW_BYTES / TILE_BYTES : 1;  // tiles per row
    static constexpr int GROUP = ROW_BYTES >= TILE_BYTES ? 1 : TILE_BYTES / ROW_BYTES;  // rows per tile
    static constexpr int TILE_BINS = SUB > 1 ? TILE_BYTES / 4 : BINS;
    static constexpr int TABLE_READS = GROUP + 1;
};

template <int K, int TILE_BYTES>
struct BigRowSink {
    using G = BigGeo<K, TILE_BYTES>;
    static constexpr bool kCtaUniform = true;
    static constexpr bool kSharedRows = false;
    bool tma;         // zeros were streamed by TMA (wait for the bulk group) or by plain stores
    uint32_t* rows;   // row of table-local read 0
    int sub;          // which slice of the row this tile covers (SUB > 1)
    bool has_last;
    int qb, period;   // chunk openers (see DenseSink)
    __device__ __forceinline__ void before_first_emit()
    {
        if (tma && threadIdx.x == 0) { bulk_wait_all(); fence_async_proxy_global(); }
        __syncthreads();   // plain stores: the barrier orders them before the CTA's reductions
    }
    __device__ __forceinline__ void kmer(int q, uint32_t idx)
    {
        if (G::SUB > 1 && (int)(idx / G::TILE_BINS) != sub) return;
        atomicAdd(rows + (int64_t)q * G::BINS + idx, 1u);
    }
    __device__ __forceinline__ void invalid(int q, int in_read, int extra)
    {
        if (period > 0 && q >= qb && (q - qb) % period == 0) return;
        if (q >= 1 && has_last) atomicAdd(rows + (int64_t)q * G::BINS - 1, (uint32_t)(in_read + extra));
    }
};

constexpr int kBigZeroBytes = 32 << 10;    // constant zero source
constexpr int kBigThreads = 256;

template <int K, int FMT, int TILE_BYTES>
__global__ void __launch_bounds__(kBigThreads) dense_bigrow_kernel(const DenseArgs a)
{
    using G = BigGeo<K, TILE_BYTES>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint4* zero = reinterpret_cast<uint4*>(smem);
    unsigned char* tp = smem + kBigZeroBytes;
    ReadTable tb;
    tb.start = reinterpret_cast<int64_t*>(tp);
    tb.tend = reinterpret_cast<int32_t*>(tp + G::TABLE_READS * 8);
    tb.extra = reinterpret_cast<int32_t*>(tp + G::TABLE_READS * 12);
    tb.cum = reinterpret_cast<uint32_t*>(tp + G::TABLE_READS * 16);

    for (int i = threadIdx.x; i < kBigZeroBytes / 16; i += kBigThreads) zero[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_proxy_shared();
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int64_t group = tile / G::SUB;
        const int sub = (int)(tile - group * G::SUB);
        const int64_t r0 = a.read_begin + group * G::GROUP;
        const int nrows = (int)min((int64_t)G::GROUP, a.read_end - r0);
        const bool has_last = (sub == G::SUB - 1);
        int qb = 0, period = 0;
        if (a.chunk_size > 0) {
            const int64_t phase = (a.index_base + r0) % a.chunk_size;
            const int64_t first = phase == 0 ? 0 : a.chunk_size - phase;
            if (first <= G::TABLE_READS) {
                qb = (int)first;
                period = (int)min(a.chunk_size, (int64_t)(4 * kMaxGroupReads));
            }
        } else if (r0 == 0) {
            period = 4 * kMaxGroupReads;
        }
        const bool next_opens_chunk = period > 0 && nrows >= qb && (nrows - qb) % period == 0;
        const bool halo = a.mode == MODE_COMPAT && has_last && (r0 + nrows < a.nS) && !next_opens_chunk;
        const int nreads = nrows + (halo ? 1 : 0);
        uint32_t* rows = a.out + (r0 - a.read_begin) * G::BINS;

        if (a.flags & 1) {       // 1'. zero stream by plain 16-byte stores from every thread
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(rows) + (int64_t)sub * TILE_BYTES);
            const int n16 = (int)((G::SUB > 1 ? (int64_t)TILE_BYTES : (int64_t)nrows * G::ROW_BYTES) / 16);
#pragma unroll 4
            for (int i = threadIdx.x; i < n16; i += kBigThreads)
                asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(dst + i), "r"(0u) : "memory");
        } else if (threadIdx.x == 0) {  // 1. zero stream by TMA
            unsigned char* dst = reinterpret_cast<unsigned char*>(rows) + (int64_t)sub * TILE_BYTES;
            int64_t left = G::SUB > 1 ? (int64_t)TILE_BYTES : (int64_t)nrows * G::ROW_BYTES;
            while (left > 0) {
                const uint32_t n = (uint32_t)min((int64_t)kBigZeroBytes, left);
                bulk_store_issue(dst, zero, n);
                dst += n; left -= n;
            }
            bulk_commit();
        }
        fill_read_table<K>(tb, a.start, a.length, r0, nreads, a.mode, a.nN, a.nS, a.chunk_size, a.index_base);
        __syncthreads();
        if (threadIdx.x < 32) scan_read_table(tb, nreads);
        __syncthreads();
        BigRowSink<K, TILE_BYTES> sink{(a.flags & 1) == 0, rows, sub, has_last, qb, period};
        for_each_window<K, FMT, G::TABLE_READS>(BasesRef{a.bases, a.valid}, tb, nreads, nrows, a.mode, sink);  // 2. + 3.
        __syncthreads();  // table is reused by the next tile
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

template <int K, int FMT, int TILE_BYTES>
static cudaError_t launch_bigrow_t(const DenseArgs& a0, cudaStream_t st, int ctas_cap)
{
    using G = BigGeo<K, TILE_BYTES>;
    auto kern = dense_bigrow_kernel<K, FMT, TILE_BYTES>;
    const int smem = kBigZeroBytes + G::TABLE_READS * 16 + (G::TABLE_READS + 1) * 4 + 16;
    static thread_local int configured_dev = -1;
    static thread_local int ctas_per_sm = 0, num_sms = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kBigThreads, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured_dev = dev;
    }
    DenseArgs a = a0;
    a.num_tiles = (a.read_end - a.read_begin + G::GROUP - 1) / G::GROUP * G::SUB;
    if (a.num_tiles <= 0) return cudaSuccess;
    // Rows must still be L2-resident when their reductions arrive: bound the bytes in flight
    // (CTAs x tile) well below the 126 MB L2 (measured: profiles/r1_notes.md).
    const int per_sm = ctas_cap > 0 && ctas_cap < ctas_per_sm ? ctas_cap : ctas_per_sm;
    const int64_t resident = (int64_t)num_sms * per_sm;
    const unsigned grid = (unsigned)(a.num_tiles < resident ? a.num_tiles : resident);
    kern<<<grid, kBigThreads, smem, st>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <int K, int FMT>
static cudaError_t launch_bigrow_k(const DenseArgs& a, cudaStream_t st)
{
    // measured optimum (profiles/r1_notes.md): k=7 32 KiB x 5 CTAs/SM (23.7 MB in flight) 0.965,
    // k=8 64 KiB x 3 (28.4 MB) 1.06; the curve is sharp (k=7: 32x4 0.81, 32x6 0.92, 64x3 0.93)
    static const int tile_kb = env_int("CFRK_BIG_TILE_KB", K == 7 ? 32 : 64);
    static const int ctas = env_int("CFRK_BIG_CTAS", K == 7 ? 5 : 3);
    switch (tile_kb) {
    case 16: return launch_bigrow_t<K, FMT, (16 << 10)>(a, st, ctas);
    case 32: return launch_bigrow_t<K, FMT, (32 << 10)>(a, st, ctas);
    case 128: return launch_bigrow_t<K, FMT, (128 << 10)>(a, st, ctas);
    case 256: return launch_bigrow_t<K, FMT, (256 << 10)>(a, st, ctas);
    default: return launch_bigrow_t<K, FMT, (64 << 10)>(a, st, ctas);
    }
}

// ------------------------------------------------------------------------------------------
constexpr int kTileBins = 4096;  // 16 KiB tiles
constexpr int kDenseThreads = 256;
constexpr int kDenseBufs = 2;

template <int K, int FMT, int TILE_BINS_T = kTileBins, int NTHREADS = kDenseThreads>
static cudaError_t launch_dense_k(const DenseArgs& a0, cudaStream_t st)
{
    using G = Geo<K, TILE_BINS_T>;
    auto kern = dense_count_kernel<K, FMT, TILE_BINS_T, NTHREADS, kDenseBufs>;
    const int smem = kDenseBufs * G::TILE_BYTES + G::table_bytes() + G::ROW_ALIGN;
    static thread_local int configured_dev = -1;
    static thread_local int ctas_per_sm = 0, num_sms = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NTHREADS, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured_dev = dev;
    }
    DenseArgs a = a0;
    const int64_t groups = (a.read_end - a.read_begin + G::RPT - 1) / G::RPT;
    a.num_tiles = groups * G::SUB;
    if (a.num_tiles <= 0) return cudaSuccess;
    const int64_t resident = (int64_t)num_sms * ctas_per_sm;  // persistent: one wave
    const unsigned grid = (unsigned)(a.num_tiles < resident ? a.num_tiles : resident);
    kern<<<grid, NTHREADS, smem, st>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <int FMT>
static cudaError_t launch_dense_fmt(int k, const DenseArgs& a, cudaStream_t st)
{
    // Which kernel serves which k is a measured choice (profiles/r1_notes.md).  The environment
    // switches below exist to re-run those A/B measurements; the defaults are the winners.
    //   launch_warp_k<K, FMT, RW, WARPS, DIRECT>: warp tiles of RW rows, WARPS warps per CTA,
    //   DIRECT = read-clear-store (one buffer) instead of TMA stores (two buffers).
    static const int k4_variant = env_int("CFRK_K4", 0);
    static const int k5_variant = env_int("CFRK_K5", 0);
    // k <= 4: lane-per-read tiles with TMA-staged input (dense_lane.cu); CFRK_DENSE_LANE = 0 selects the
    // round-1 kernels below for A/B runs, a bit mask selects per k (bit k-1)
    static const int lane_mask = env_int("CFRK_DENSE_LANE", 15);
    if (k <= 4 && (lane_mask >> (k - 1) & 1)) return launch_dense_lane(k, FMT, a, st);
    if (k == 4) {
        switch (k4_variant) {
        case 1: return launch_dense_k<4, FMT, 4096, 256>(a, st);      // CTA tiles 16 KiB            0.64
        case 2: return launch_dense_k<4, FMT, 4096, 192>(a, st);      // CTA tiles, 6 warps          0.66
        case 3: return launch_warp_k<4, FMT, 4, 4, false>(a, st);     // warp tiles + TMA, RW=4      0.60
        case 4: return launch_warp_k<4, FMT, 8, 4, true>(a, st);      // direct, RW=8 (93 items)     0.69-0.75
        case 5: return launch_warp_k<4, FMT, 16, 4, true>(a, st);     // direct, RW=16               0.53
        case 6: return launch_warp_k<4, FMT, 6, 4, true>(a, st);      // direct, RW=6 (62 items = 2 chunks)  0.61
        case 7: return launch_warp_k<4, FMT, 3, 4, true>(a, st);      // direct, RW=3 (31 items = 1 chunk)   0.46
        default: return launch_warp_k<4, FMT, 9, 4, true>(a, st);     // direct, RW=9 (93 items, no halo scan) +2.5 %
        }
    }
    if (k == 5) {
        switch (k5_variant) {
        case 1: return launch_dense_k<5, FMT>(a, st);                 // CTA tiles                   0.70
        case 2: return launch_warp_k<5, FMT, 1, 4, false>(a, st);     // warp tiles + TMA, RW=1      0.79
        case 3: return launch_warp_k<5, FMT, 1, 4, true>(a, st);      // direct, RW=1                0.77-0.87
        case 4: return launch_warp_k<5, FMT, 2, 4, true>(a, st);      // direct, RW=2                0.89
        case 5: return launch_warp_k<5, FMT, 3, 4, true>(a, st);      // direct, RW=3, 4 warps       0.85
        default: return launch_warp_k<5, FMT, 3, 3, true>(a, st);     // direct, RW=3 (31 items = one full chunk), 3 warps  0.91
        }
    }
    switch (k) {
    case 1: return launch_dense_k<1, FMT>(a, st);
    case 2: return launch_dense_k<2, FMT>(a, st);
    case 3: return launch_dense_k<3, FMT>(a, st);
    case 6: {
        static const int k6_variant = env_int("CFRK_K6", 2);  // 0: big-row path, 1/2: warp tiles (direct), 3: CTA tiles
        if (k6_variant == 1) return launch_warp_k<6, FMT, 1, 4, true>(a, st);
        if (k6_variant == 2) return launch_warp_k<6, FMT, 1, 6, true>(a, st);
        if (k6_variant == 3) return launch_dense_k<6, FMT>(a, st);
        return launch_bigrow_k<6, FMT>(a, st);
    }
    case 7: {
        static const int k7_variant = env_int("CFRK_K7", 0);  // 0: big-row path (TMA zeros + L2 reductions) 0.965, 1: one row per CTA 0.92
        return k7_variant == 1 ? launch_row_k<7, FMT>(a, st) : launch_bigrow_k<7, FMT>(a, st);
    }
    case 8: return launch_bigrow_k<8, FMT>(a, st);
    // k = 9..12: 1 MiB .. 64 MiB per read.  Only the reference's operator can ask for these (its
    // own driver overflows beyond k = 8); same path, 256 KiB tiles, 2 CTAs/SM.
    case 9: return launch_bigrow_t<9, FMT, (256 << 10)>(a, st, 2);
    case 10: return launch_bigrow_t<10, FMT, (256 << 10)>(a, st, 2);
    case 11: return launch_bigrow_t<11, FMT, (256 << 10)>(a, st, 2);
    case 12: return launch_bigrow_t<12, FMT, (256 << 10)>(a, st, 2);
    default: return cudaErrorInvalidValue;
    }
}

int dense_reads_per_tile(int k)
{
    static const int lane_mask = env_int("CFRK_DENSE_LANE", 15);
    if (k >= 1 && k <= 4 && (lane_mask >> (k - 1) & 1)) return dense_lane_reads_per_tile(k);
    switch (k) {
    case 1: return Geo<1, kTileBins>::RPT;
    case 2: return Geo<2, kTileBins>::RPT;
    case 3: return Geo<3, kTileBins>::RPT;
    case 4: return Geo<4, kTileBins>::RPT;
    case 5: return Geo<5, kTileBins>::RPT;
    default: return 1;
    }
}

cudaError_t launch_dense(const void* bases, int fmt, const int64_t* start, const int32_t* length,
                         int64_t nN, int64_t nS, int64_t read_begin, int64_t read_end, int k, int mode,
                         int64_t chunk_size, int64_t index_base, int32_t* out, cudaStream_t st,
                         const uint16_t* packed_valid)
{
    DenseArgs a;
    a.bases = static_cast<const uint8_t*>(bases);
    a.valid = packed_valid;
    a.start = start; a.length = length; a.nS = nS; a.nN = nN;
    a.read_begin = read_begin; a.read_end = read_end;
    a.out = reinterpret_cast<uint32_t*>(out);
    a.mode = mode; a.num_tiles = 0;
    a.chunk_size = chunk_size; a.index_base = index_base;
    static const int big_plain = env_int("CFRK_BIG_PLAIN", 0);
    a.flags = big_plain ? 1 : 0;
    a.handoff = nullptr;
    if (fmt == FMT_PACKED) return launch_dense_fmt<FMT_PACKED>(k, a, st);
    return fmt == FMT_ASCII ? launch_dense_fmt<FMT_ASCII>(k, a, st) : launch_dense_fmt<FMT_CODES>(k, a, st);
}

// ------------------------------------------------------------------------------------------
// Whole-dataset histogram: same item loop.  k <= 7: the CTA counts into a private shared-memory
// histogram (red.shared) and adds it to the global one once at the end -- with 4^k <= 16384 bins a
// direct red.global per window serialises in L2 (measured 4.4 Gbases/s at k=4).  k >= 8: one
// red.global per valid window into the L2-resident histogram (186 G reductions/s at k=12).
struct HistSink {
    static constexpr bool kCtaUniform = false;
    static constexpr bool kSharedRows = false;
    uint32_t* hist;
    __device__ __forceinline__ void kmer(int, uint32_t idx) { atomicAdd(&hist[idx], 1u); }
    __device__ __forceinline__ void invalid(int, int, int) {}
};
struct HistSinkShared {
    static constexpr bool kCtaUniform = false;
    static constexpr bool kSharedRows = true;
    static constexpr bool kRowsAligned = false;
    uint32_t saddr;
    __device__ __forceinline__ uint32_t row_saddr(int) const { return saddr; }
    __device__ __forceinline__ void invalid(int, int, int) {}
};

constexpr int kHistGroup = 128;  // reads per work group
constexpr int kHistThreads = 256;
constexpr int kHistSharedMaxK = 7;

template <int K, int FMT>
__global__ void __launch_bounds__(kHistThreads) global_hist_kernel(const uint8_t* __restrict__ bases,
                                                                 const int64_t* __restrict__ start,
                                                                 const int32_t* __restrict__ length,
                                                                 int64_t nS, uint32_t* hist)
{
    constexpr bool SHARED = K <= kHistSharedMaxK;
    constexpr int BINS = 1 << (2 * K);
    __shared__ int64_t s_start[kHistGroup];
    __shared__ int32_t s_tend[kHistGroup], s_extra[kHistGroup];
    __shared__ uint32_t s_cum[kHistGroup + 1];
    extern __shared__ __align__(16) uint32_t s_hist[];
    ReadTable tb{s_start, s_tend, s_extra, s_cum};
    if (SHARED) {
        for (int i = threadIdx.x; i < BINS; i += kHistThreads) s_hist[i] = 0u;
    }
    HistSink gsink{hist};
    HistSinkShared ssink{(uint32_t)__cvta_generic_to_shared(s_hist)};
    const int64_t groups = (nS + kHistGroup - 1) / kHistGroup;
    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const int64_t r0 = g * kHistGroup;
        const int n = (int)min((int64_t)kHistGroup, nS - r0);
        __syncthreads();  // previous group's item loop is done with the table (and s_hist is zeroed)
        fill_read_table<K>(tb, start, length, r0, n, MODE_EXACT, INT64_MAX);
        __syncthreads();
        if (threadIdx.x < 32) scan_read_table(tb, n);
        __syncthreads();
        if constexpr (SHARED) for_each_window<K, FMT, kHistGroup>(BasesRef{bases, nullptr}, tb, n, n, MODE_EXACT, ssink);
        else for_each_window<K, FMT, kHistGroup>(BasesRef{bases, nullptr}, tb, n, n, MODE_EXACT, gsink);
    }
    if (SHARED) {
        __syncthreads();
        for (int i = threadIdx.x; i < BINS; i += kHistThreads) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(&hist[i], v);
        }
    }
}

template <int K, int FMT>
static cudaError_t launch_hist_k(const void* bases, const int64_t* start, const int32_t* length,
                                 int64_t nS, uint32_t* hist, cudaStream_t st)
{
    auto kern = global_hist_kernel<K, FMT>;
    int dev = 0, num_sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int smem = K <= kHistSharedMaxK ? (1 << (2 * K)) * 4 : 0;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kHistThreads, smem);
    if (e != cudaSuccess) return e;
    const int64_t groups = (nS + kHistGroup - 1) / kHistGroup;
    if (groups <= 0) return cudaSuccess;
    const int64_t resident = (int64_t)num_sms * (per_sm < 1 ? 1 : per_sm);
    const unsigned grid = (unsigned)(groups < resident ? groups : resident);
    kern<<<grid, kHistThreads, smem, st>>>(static_cast<const uint8_t*>(bases), start, length, nS, hist);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <int FMT>
static cudaError_t launch_hist_fmt(int k, const void* b, const int64_t* s, const int32_t* l, int64_t nS,
                                   uint32_t* h, cudaStream_t st)
{
    switch (k) {
#define CASE(KK) case KK: return launch_hist_k<KK, FMT>(b, s, l, nS, h, st);
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
    CASE(11) CASE(12) CASE(13) CASE(14) CASE(15)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_global_hist(const void* bases, int fmt, const int64_t* start, const int32_t* length,
                               int64_t nN, int64_t nS, int k, uint32_t* hist, cudaStream_t st)
{
    if (hist_split_applies(k, nN)) return launch_global_hist_split(bases, fmt, start, length, nN, nS, k, hist, st);
    return fmt == FMT_ASCII ? launch_hist_fmt<FMT_ASCII>(k, bases, start, length, nS, hist, st)
                            : launch_hist_fmt<FMT_CODES>(k, bases, start, length, nS, hist, st);
}

// ------------------------------------------------------------------------------------------
// bases -> packed 2-bit words + validity masks; one 16-byte block per thread.
template <int FMT>
__global__ void __launch_bounds__(256) encode_2bit_kernel(const uint8_t* __restrict__ bases, int64_t nblocks,
                                                          int64_t n, uint32_t* __restrict__ codes,
                                                          uint16_t* __restrict__ valid)
{
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks;
         b += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c, v;
        encode16<FMT>(ld_block(bases + b * 16), c, v);
        const int64_t rem = n - b * 16;  // bytes of this block inside the buffer
        if (rem < 16) v &= ~from_pos((int)rem);
        codes[b] = c;
        valid[b] = (uint16_t)v;
    }
}

cudaError_t launch_encode_2bit(const void* bases, int fmt, int64_t n, uint32_t* codes, uint16_t* valid,
                               cudaStream_t st)
{
    const int64_t nblocks = (n + 15) / 16;
    if (nblocks <= 0) return cudaSuccess;
    int64_t grid = (nblocks + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    if (fmt == FMT_ASCII)
        encode_2bit_kernel<FMT_ASCII><<<(unsigned)grid, 256, 0, st>>>(static_cast<const uint8_t*>(bases), nblocks, n, codes, valid);
    else
        encode_2bit_kernel<FMT_CODES><<<(unsigned)grid, 256, 0, st>>>(static_cast<const uint8_t*>(bases), nblocks, n, codes, valid);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

}  // namespace cfrk
