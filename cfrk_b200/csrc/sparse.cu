// sparse.cu -- sparse per-read k-mer counts for k = 1..31 (exact semantics): for each read the
// sorted distinct k-mers and their multiplicities.  No reference counterpart: the reference's dense
// rows stop being usable at k = 9 (4^k int32 per read; int overflow); this is the
// path BASELINE.json configs 3 and 4 name.
//
//   rows     row r owns keys/counts[row_begin[r] .. row_begin[r] + row_count[r]), where
//            row_begin[r] = sum_{j<r} max(0, len_j - k + 1) (every window distinct: worst case), so
//            rows are written independently; callers compact if they want tight CSR.
//   short reads (<= 512 windows): ONE WARP per read.  The read is encoded once into a per-warp
//            bit stream in shared memory (2-bit codes + validity); every lane pulls the bases and
//            validity bits behind its E consecutive windows into registers and forms each window
//            with two funnel shifts (stream_device.cuh); the 32*E keys are sorted by a bitonic
//            network that lives entirely in registers (blocked layout: strides < E are register
//            min/max, the others shfl.xor + compare-xor-select) and run-length encoded with two
//            warp scans (all-distinct reads skip them).
//   medium reads (513..4096 windows): ONE CTA per read, keys grouped, sorted by warps and run-length
//            encoded entirely in shared memory (details at "medium rows" below).
//   long reads: MSD bucket partition over the bases (counting pass with slab-private shared-memory
//            counters, scatter pass with one L2 cursor per bucket), one warp sort + RLE per bucket
//            with the same register network, compaction of the runs into the row; 32-bit suffixes
//            for rows whose key bits below the bucket digit fit 32 bits; segmented LSD radix sort
//            (8-bit digits, one warp per 2048-key tile, stable ranks from match.any) for buckets
//            above 512 keys.  Details at "long reads" below.
//   Everything is hand-written; cub::DeviceScan (exclusive sums of offsets) is the only library
//   call (plumbing).
#include "kernels.h"
#include "kmer_device.cuh"
#include "stream_device.cuh"

#include <cub/device/device_scan.cuh>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include <algorithm>

namespace cfrk {

extern void count_launch();

// Scratch comes from the stream-ordered pool.  By default the pool gives freed memory back to the
// OS at the next synchronisation, so every call would map its gigabytes of scratch again (measured:
// 1.4 s per call for 3.3 GB); keep it cached instead.
void keep_pool_memory(int dev)
{
    static thread_local int done_for = -1;
    if (done_for == dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_for = dev;
}

// CFRK_TRACE=1: per-phase wall times of the sparse path on stderr (synchronises the stream)
struct SparseTrace {
    bool on = getenv("CFRK_TRACE") != nullptr;
    cudaStream_t st;
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    explicit SparseTrace(cudaStream_t s) : st(s) {}
    void mark(const char* what)
    {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cfrk sparse] %9.3f ms  %s\n", std::chrono::duration<double, std::milli>(now - last).count(), what);
        last = now;
    }
};

constexpr int kShortMaxWindows = 512;
constexpr int kStreamBlocks = (15 + kShortMaxWindows + 30 + 15) / 16 + 3;  // 16-base blocks per warp stream
constexpr int kSparseWarps = 8;

// ------------------------------------------------------------------------------------------
// window counts -> row_begin (exclusive scan done by cub::DeviceScan: plumbing)
// also counts the reads per size class, so that only the kernels that have work are launched:
// cls[0] reads of 1..144 windows, cls[1] 145..256, cls[2] 257..512, cls[3] longer
__global__ void nwin_kernel(const int32_t* __restrict__ length, int64_t nS, int k, int64_t* __restrict__ out,
                            unsigned long long* __restrict__ cls)
{
    unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nS; r += (int64_t)gridDim.x * blockDim.x) {
        const int n = r < nS ? max(0, length[r] - k + 1) : 0;
        out[r] = n;
        c0 += n >= 1 && n <= 144;
        c1 += n > 144 && n <= 256;
        c2 += n > 256 && n <= 512;
        c3 += n > 512;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, d);
        c1 += __shfl_xor_sync(0xffffffffu, c1, d);
        c2 += __shfl_xor_sync(0xffffffffu, c2, d);
        c3 += __shfl_xor_sync(0xffffffffu, c3, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(cls + 0, c0);
        if (c1) atomicAdd(cls + 1, c1);
        if (c2) atomicAdd(cls + 2, c2);
        if (c3) atomicAdd(cls + 3, c3);
    }
}

// ------------------------------------------------------------------------------------------
// the lower lane of a pair keeps the minimum, the upper one the maximum: one compare whose result
// is XOR-ed with the lane's side (ISETP.LT.XOR) and one select per 32-bit word
template <typename KeyT>
__device__ __forceinline__ KeyT keep_minmax(KeyT mine, KeyT other, bool upper)
{
    return ((mine < other) != upper) ? mine : other;
}

// Bitonic network over 32*E keys in blocked layout (lane L holds elements L*E .. L*E+E-1), in the
// direction-free form: every merge starts with a MIRRORED compare (i with i ^ (size-1)) and goes on
// with the usual strides, and every comparator leaves the minimum at the lower index.  In-lane
// comparators are then pure min/max on compile-time registers; cross-lane ones need one per-lane
// predicate per stage.
// MAXSIZE < 32*E sorts every aligned block of MAXSIZE slots on its own (the grouped path below).
template <typename KeyT, int E, int MAXSIZE = 32 * E>
__device__ __forceinline__ void bitonic_sort_blocked(KeyT (&key)[E])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int size = 2; size <= MAXSIZE; size <<= 1) {
        // mirrored step
        if (size <= E) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int p = e ^ (size - 1);
                if (p > e) {
                    const KeyT a = key[e], b = key[p];
                    key[e] = a < b ? a : b;
                    key[p] = a < b ? b : a;
                }
            }
        } else {
            const int lm = size / E - 1;
            const bool upper = (lane & (size / (2 * E))) != 0;
            KeyT other[E];
#pragma unroll
            for (int e = 0; e < E; e++) other[e] = __shfl_xor_sync(0xffffffffu, key[E - 1 - e], lm);
#pragma unroll
            for (int e = 0; e < E; e++) key[e] = keep_minmax<KeyT>(key[e], other[e], upper);
        }
        // remaining strides
#pragma unroll
        for (int stride = size >> 2; stride >= 1; stride >>= 1) {
            if (stride >= E) {
                const int ls = stride / E;
                const bool upper = (lane & ls) != 0;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const KeyT other = __shfl_xor_sync(0xffffffffu, key[e], ls);
                    key[e] = keep_minmax<KeyT>(key[e], other, upper);
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const int p = e ^ stride;
                    if (p > e) {
                        const KeyT a = key[e], b = key[p];
                        key[e] = a < b ? a : b;
                        key[p] = a < b ? b : a;
                    }
                }
            }
        }
    }
}

// Sorted keys in blocked layout -> (key, count) pairs.  The lane's E slots are consecutive in the
// sorted sequence of REAL keys: slot e is real iff bit e of `live` (a low-bit mask), and then it is
// element number vpos0 + e of that sequence; nvalid = its length.  first0: the lane's slot 0 starts
// a sorted group, so it is a head whatever the key before it (plain case: lane 0 only).
template <typename KeyT, int E, bool BLOCKED = false>
__device__ __forceinline__ int warp_rle_store(const KeyT (&key)[E], uint32_t live, int vpos0, bool first0, int nvalid,
                                              KeyT* __restrict__ keys_out, uint32_t* __restrict__ counts_out,
                                              KeyT* __restrict__ stage_k, uint32_t* __restrict__ stage_c)
{
    const int lane = threadIdx.x & 31;
    KeyT prev = __shfl_up_sync(0xffffffffu, key[E - 1], 1);
    uint32_t heads = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const bool head = (live >> e & 1u) && ((e == 0 && first0) || key[e] != prev);
        heads |= head ? (1u << e) : 0u;
        prev = key[e];
    }
    // common case: every key is distinct -> element v goes to slot v with count 1
    if (__all_sync(0xffffffffu, heads == live)) {
        if (BLOCKED) {   // vpos0 == lane * E: the lane's slots are one aligned block, pads included
            constexpr int V = 16 / (int)sizeof(KeyT);   // keys per 16-byte shared-memory store
            static_assert(E % V == 0, "blocked keys are staged with 16-byte stores");
#pragma unroll
            for (int q = 0; q < E / V; q++) {
                if (sizeof(KeyT) == 4)
                    reinterpret_cast<uint4*>(stage_k + lane * E)[q] = make_uint4((uint32_t)key[4 * q], (uint32_t)key[4 * q + 1], (uint32_t)key[4 * q + 2], (uint32_t)key[4 * q + 3]);
                else
                    reinterpret_cast<ulonglong2*>(stage_k + lane * E)[q] = make_ulonglong2((unsigned long long)key[2 * q], (unsigned long long)key[2 * q + 1]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; e++)
                if (live >> e & 1u) stage_k[vpos0 + e] = key[e];
        }
        __syncwarp();
        for (int i = lane; i < nvalid; i += 32) {
            keys_out[i] = stage_k[i];
            counts_out[i] = 1u;
        }
        __syncwarp();
        return nvalid;
    }
    // exclusive scan of head counts -> first output slot of this lane
    const int hc = __popc(heads);
    int inc = hc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    int off = inc - hc;
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    // sequence position of the next head after this lane: suffix-min of each lane's first head
    int first = heads ? vpos0 + (__ffs(heads) - 1) : nvalid;
    int nxt = __shfl_down_sync(0xffffffffu, first, 1);
    if (lane == 31) nxt = nvalid;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_down_sync(0xffffffffu, nxt, d);
        if (lane + d < 32) nxt = min(nxt, o);
    }
    nxt = min(nxt, nvalid);
    // walk the lane's elements backwards: a head's run ends at the next head
    uint32_t cnt[E];
#pragma unroll
    for (int e = E - 1; e >= 0; e--) {
        cnt[e] = (uint32_t)(nxt - (vpos0 + e));
        if (heads & (1u << e)) nxt = vpos0 + e;
    }
    // stage the pairs in shared memory, then write the row with coalesced stores
#pragma unroll
    for (int e = 0; e < E; e++) {
        if (heads & (1u << e)) {
            stage_k[off] = key[e];
            stage_c[off] = cnt[e];
            off++;
        }
    }
    __syncwarp();
    for (int i = lane; i < total; i += 32) {
        keys_out[i] = stage_k[i];
        counts_out[i] = stage_c[i];
    }
    __syncwarp();
    return total;
}

// plain case: the first nvalid slots of the warp are the real keys
template <typename KeyT, int E>
__device__ __forceinline__ int warp_rle_store(const KeyT (&key)[E], int nvalid, KeyT* __restrict__ keys_out,
                                              uint32_t* __restrict__ counts_out, KeyT* __restrict__ stage_k,
                                              uint32_t* __restrict__ stage_c)
{
    const int lane = threadIdx.x & 31;
    const int nlive = min(E, max(0, nvalid - lane * E));
    return warp_rle_store<KeyT, E, true>(key, (1u << nlive) - 1u, lane * E, lane == 0, nvalid, keys_out, counts_out, stage_k, stage_c);
}

// One read, E keys per lane: extract, sort, run-length encode.
//
// GROUPED (E = 8, reads of <= 160 windows, k >= 2): the keys are first split by their top 3 bits into
// 8 groups of <= 32 slots (4 lanes each) through shared memory, ranks from a warp scan of byte-packed
// per-lane group counts (no shared counters, no match.any), and every group is sorted on its own: a
// 32-slot network has 3 cross-lane stages instead of the 15 of the 256-slot one.  Groups are disjoint
// key ranges in ascending order, so the concatenation is sorted.  A group that overflows
// (low-complexity reads) sends the read through the full network instead.
constexpr int kGroupSlots = 32;
constexpr int kGroupedMaxWindows = 160;   // mean <= 20 keys per group on random reads: overflow is rare

// key[]: E keys per lane in any order, bit e of `valid` set for the real ones (the others hold the
// maximum key).  gshift >= 0 (E = 8, at most 160 real keys): grouped, the group of a key is bits
// gshift..gshift+2.  Returns the number of pairs written to keys_out / counts_out.
template <typename KeyT, int E>
__device__ __forceinline__ int warp_sort_rle(KeyT (&key)[E], uint32_t valid, int gshift,
                                             KeyT* __restrict__ keys_out, uint32_t* __restrict__ counts_out,
                                             KeyT* __restrict__ stage_k, uint32_t* __restrict__ stage_c)
{
    const int lane = threadIdx.x & 31;
    if (E == 8 && gshift >= 0) {
        // Split without shared counters: every lane packs its keys-per-group into 8 bytes of a 64-bit
        // word; a warp scan of that word gives, byte by byte, the keys of each group in the lanes
        // before it; the rank of a key = that + the lane's own earlier keys of the group.  (At most
        // 160 keys in all, so no byte overflows.)
        constexpr int V = 16 / (int)sizeof(KeyT);
        uint64_t h = 0;
        uint32_t lrank = 0;   // 4 bits per slot: keys of the same group in earlier slots of this lane
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (valid >> e & 1u) {
                const uint32_t sh = ((uint32_t)(key[e] >> gshift) & 7u) * 8u;
                lrank |= ((uint32_t)(h >> sh) & 0xFu) << (4 * e);
                h += 1ull << sh;
            }
        }
        uint64_t inc = h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        const uint64_t excl = inc - h;
        const uint64_t totals = __shfl_sync(0xffffffffu, inc, 31);
        // a group above 32 keys (byte + 95 >= 128): the full network below
        if (((totals + 0x5F5F5F5F5F5F5F5Full) & 0x8080808080808080ull) == 0) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                if (valid >> e & 1u) {
                    const uint32_t d = (uint32_t)(key[e] >> gshift) & 7u;
                    const uint32_t slot = ((uint32_t)(excl >> (8 * d)) & 0xFFu) + ((lrank >> (4 * e)) & 0xFu);
                    stage_k[d * kGroupSlots + slot] = key[e];
                }
            }
            __syncwarp();
            const int g = lane >> 2;
            const int ng = (int)((totals >> (8 * g)) & 0xFFu);
            const int gstart = (int)(((totals * 0x0101010101010100ull) >> (8 * g)) & 0xFFu);   // keys of the groups before g
            const int nvalid = (int)((totals * 0x0101010101010101ull) >> 56);
#pragma unroll
            for (int q = 0; q < E / V; q++) {
                if (sizeof(KeyT) == 4) {
                    const uint4 v = reinterpret_cast<const uint4*>(stage_k + lane * E)[q];
                    key[4 * q] = (KeyT)v.x; key[4 * q + 1] = (KeyT)v.y; key[4 * q + 2] = (KeyT)v.z; key[4 * q + 3] = (KeyT)v.w;
                } else {
                    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(stage_k + lane * E)[q];
                    key[2 * q] = (KeyT)v.x; key[2 * q + 1] = (KeyT)v.y;
                }
            }
            const int sig0 = (lane & 3) * E;                 // the lane's first slot inside its group
            const int nlive = min(E, max(0, ng - sig0));
#pragma unroll
            for (int e = 0; e < E; e++)
                if (e >= nlive) key[e] = KeyMax<KeyT>::value;   // slots behind the group's keys hold leftovers: pad
            __syncwarp();   // the staging buffer is free again
            bitonic_sort_blocked<KeyT, E, kGroupSlots>(key);
            return warp_rle_store<KeyT, E>(key, (1u << nlive) - 1u, gstart + sig0, (lane & 3) == 0, nvalid, keys_out, counts_out,
                                           stage_k, stage_c);
        }
        // overflow: the full network on the keys as extracted
    }
    int nvalid = __popc(valid);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, d);
    bitonic_sort_blocked<KeyT, E>(key);
    return warp_rle_store<KeyT, E>(key, nvalid, keys_out, counts_out, stage_k, stage_c);
}

template <typename KeyT, int E>
__device__ __forceinline__ int warp_count_read(const WarpStream& st, int a, int k, bool grouped,
                                               KeyT* __restrict__ keys_out, uint32_t* __restrict__ counts_out,
                                               KeyT* __restrict__ stage_k, uint32_t* __restrict__ stage_c)
{
    const int lane = threadIdx.x & 31;
    KeyT key[E];
    const uint32_t valid = extract_windows<KeyT, E>(st, a + lane * E, k, key);
    return warp_sort_rle<KeyT, E>(key, valid, grouped ? 2 * k - 3 : -1, keys_out, counts_out, stage_k, stage_c);
}

// E = keys per lane: the kernel handles the reads whose window count falls in its class
// ((E == 4: 1..128, E == 8: 129..256, E == 16: 257..512) and leaves the others to its siblings, so
// that the common short-read case is not compiled with the register budget of the largest network.
template <int E> struct SparseCta { static constexpr int WARPS = E == 16 ? 4 : kSparseWarps; };

template <typename KeyT, int FMT, int E>
__global__ void __launch_bounds__(SparseCta<E>::WARPS * 32) sparse_short_kernel(
    const BasesRef src, const int64_t* __restrict__ start, const int32_t* __restrict__ length,
    int64_t nS, int k, const int64_t* __restrict__ row_begin, int32_t* __restrict__ row_count,
    KeyT* __restrict__ keys, uint32_t* __restrict__ counts, bool group_split, int min_windows)
{
    constexpr int WARPS = SparseCta<E>::WARPS;
    __shared__ uint32_t s_cw[WARPS][kStreamBlocks];
    __shared__ __align__(4) uint16_t s_vh[WARPS][2 * ((kStreamBlocks + 1) / 2) + 2];
    __shared__ __align__(16) KeyT s_stage_k[WARPS][32 * E];
    __shared__ uint32_t s_stage_c[WARPS][32 * E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStream st{s_cw[warp], s_vh[warp]};
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < nS; r += nwarps) {
        const int len = length[r];
        const int nwin = len - k + 1;
        if (nwin <= 0) { if (min_windows <= 1 && lane == 0) row_count[r] = 0; continue; }
        if (nwin > 32 * E || nwin < min_windows) continue;  // another class, or the long path
        const int64_t s = start[r];
        const int a = (int)(s & 15);
        const int nblocks = (a + len + 15) >> 4;
        __syncwarp();
        for (int b = lane; b < kStreamBlocks; b += 32) {
            uint32_t c = 0, v = 0;
            if (b < nblocks) {
                decode_block<FMT>(load_block<FMT>(src, (s >> 4) + b), c, v);
                v &= from_pos(max(0, a - 16 * b)) & ~from_pos(min(16, max(0, a + len - 16 * b)));
            }
            st.cw[b] = c;
            st.vh[b ^ 1] = (uint16_t)v;
        }
        __syncwarp();
        KeyT* ko = keys + row_begin[r];
        uint32_t* co = counts + row_begin[r];
        // measured (10 M x 150 bp): uint32 keys k=12 100.1 -> 107.8 Gbases/s, uint64 keys k=20 49.4 -> 68.0
        const bool grouped = E == 8 && group_split && nwin <= kGroupedMaxWindows && k >= 2;
        const int nd = warp_count_read<KeyT, E>(st, a, k, grouped, ko, co, s_stage_k[warp], s_stage_c[warp]);
        if (lane == 0) row_count[r] = nd;
    }
}

// ------------------------------------------------------------------------------------------
// Reads of <= 144 windows (150-bp reads at k >= 7: the short-read configs): TWO reads per warp, 16 lanes
// x 9 keys each.  The 256-slot network above sorts 139 windows with 46 % padding and 15 cross-lane
// stages; 16 x 9 = 144 slots fit them with 3 %: in-lane 9-input network (25 comparators), then four
// merge levels over 2, 4, 8, 16 lanes -- a mirrored compare, the remaining lane strides, and the lane's 9
// keys (a bitonic sequence by then) through the same 25 comparators -- 10 cross-lane stages in all, and
// both halves of the warp run the same instructions on their own read.  ncu, round 1: the grouped
// 256-slot path issued 976 warp-instructions per 150-bp read on the integer pipe at 80 %; this one
// issues about 420 (profiles/r2_notes.md).
constexpr int kHalfE = 9;
constexpr int kHalfSlots = 16 * kHalfE;              // 144
constexpr int kHalfBlocks = 16;                      // 16-base blocks per half-warp stream

// optimal comparator networks for 9 inputs (25 comparators) and 11 inputs (35), checked with the 0-1 principle
template <typename KeyT, int E>
__device__ __forceinline__ void sort_lane(KeyT (&k)[E])
{
    static_assert(E == 9 || E == 11, "networks for 9 and 11 keys per lane");
#define CFRK_CE(i, j) { const KeyT a_ = k[i], b_ = k[j]; k[i] = a_ < b_ ? a_ : b_; k[j] = a_ < b_ ? b_ : a_; }
    if constexpr (E == 9) {
        CFRK_CE(0, 3) CFRK_CE(1, 7) CFRK_CE(2, 5) CFRK_CE(4, 8)
        CFRK_CE(0, 7) CFRK_CE(2, 4) CFRK_CE(3, 8) CFRK_CE(5, 6)
        CFRK_CE(0, 2) CFRK_CE(1, 3) CFRK_CE(4, 5) CFRK_CE(7, 8)
        CFRK_CE(1, 4) CFRK_CE(3, 6) CFRK_CE(5, 7)
        CFRK_CE(0, 1) CFRK_CE(2, 4) CFRK_CE(3, 5) CFRK_CE(6, 8)
        CFRK_CE(2, 3) CFRK_CE(4, 5) CFRK_CE(6, 7)
        CFRK_CE(1, 2) CFRK_CE(3, 4) CFRK_CE(5, 6)
    } else {
        CFRK_CE(0, 9) CFRK_CE(1, 6) CFRK_CE(2, 4) CFRK_CE(3, 7) CFRK_CE(5, 8)
        CFRK_CE(0, 1) CFRK_CE(3, 5) CFRK_CE(4, 10) CFRK_CE(6, 9) CFRK_CE(7, 8)
        CFRK_CE(1, 3) CFRK_CE(2, 5) CFRK_CE(4, 7) CFRK_CE(8, 10)
        CFRK_CE(0, 4) CFRK_CE(1, 2) CFRK_CE(3, 7) CFRK_CE(5, 9) CFRK_CE(6, 8)
        CFRK_CE(0, 1) CFRK_CE(2, 6) CFRK_CE(4, 5) CFRK_CE(7, 8) CFRK_CE(9, 10)
        CFRK_CE(2, 4) CFRK_CE(3, 6) CFRK_CE(5, 7) CFRK_CE(8, 9)
        CFRK_CE(1, 2) CFRK_CE(3, 4) CFRK_CE(5, 6) CFRK_CE(7, 8)
        CFRK_CE(2, 3) CFRK_CE(4, 5) CFRK_CE(6, 7)
    }
#undef CFRK_CE
}

// After the cross-lane stages of a merge level a lane holds the right E keys as a BITONIC sequence (any rotation of
// an ascending one: the 74 / 112 reachable 0-1 patterns, tools/host/lane_merge_search.py), so a merger does instead of
// the full network: 18 comparators for 9 keys (sort the columns, then the rows, of a 3 x 3 grid), 25 for 11.
template <typename KeyT, int E>
__device__ __forceinline__ void merge_lane(KeyT (&k)[E])
{
    static_assert(E == 9 || E == 11, "mergers for 9 and 11 keys per lane");
#define CFRK_CE(i, j) { const KeyT a_ = k[i], b_ = k[j]; k[i] = a_ < b_ ? a_ : b_; k[j] = a_ < b_ ? b_ : a_; }
    if constexpr (E == 9) {
        CFRK_CE(0, 3) CFRK_CE(1, 4) CFRK_CE(5, 8)
        CFRK_CE(0, 6) CFRK_CE(1, 7) CFRK_CE(2, 8)
        CFRK_CE(3, 6) CFRK_CE(2, 5) CFRK_CE(4, 7)
        CFRK_CE(0, 1) CFRK_CE(3, 4) CFRK_CE(6, 7)
        CFRK_CE(0, 2) CFRK_CE(3, 5) CFRK_CE(6, 8)
        CFRK_CE(1, 2) CFRK_CE(4, 5) CFRK_CE(7, 8)
    } else {
        CFRK_CE(0, 4) CFRK_CE(5, 10) CFRK_CE(1, 6) CFRK_CE(3, 9) CFRK_CE(2, 7)
        CFRK_CE(1, 3) CFRK_CE(6, 9) CFRK_CE(0, 8) CFRK_CE(2, 5) CFRK_CE(7, 10)
        CFRK_CE(4, 8) CFRK_CE(3, 5) CFRK_CE(6, 7) CFRK_CE(0, 1)
        CFRK_CE(8, 9) CFRK_CE(0, 2)
        CFRK_CE(1, 2) CFRK_CE(2, 3) CFRK_CE(3, 4) CFRK_CE(4, 5) CFRK_CE(5, 6) CFRK_CE(6, 7) CFRK_CE(7, 8) CFRK_CE(8, 10) CFRK_CE(9, 10)
    }
#undef CFRK_CE
}

// 16 lanes x E keys, blocked layout (lane hl holds elements E*hl .. E*hl+E-1), ascending; both halves of
// the warp at once (lane masks < 16 never cross the halves)
template <typename KeyT, int E = kHalfE>
__device__ __forceinline__ void half_sort(KeyT (&key)[E])
{
    const int hl = threadIdx.x & 15;
    sort_lane<KeyT, E>(key);
#pragma unroll
    for (int m = 2; m <= 16; m <<= 1) {
        {   // mirrored compare: element e with element E-1-e of lane hl ^ (m-1)
            const bool upper = (hl & (m >> 1)) != 0;
            KeyT other[E];
#pragma unroll
            for (int e = 0; e < E; e++) other[e] = __shfl_xor_sync(0xffffffffu, key[E - 1 - e], m - 1);
#pragma unroll
            for (int e = 0; e < E; e++) key[e] = keep_minmax<KeyT>(key[e], other[e], upper);
        }
#pragma unroll
        for (int st = m >> 2; st >= 1; st >>= 1) {
            const bool upper = (hl & st) != 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
                const KeyT other = __shfl_xor_sync(0xffffffffu, key[e], st);
                key[e] = keep_minmax<KeyT>(key[e], other, upper);
            }
        }
        merge_lane<KeyT, E>(key);
    }
}

// sorted keys of one half (first nvalid slots real) -> (key, count) pairs; both halves run in lockstep
// (every shuffle is a full-mask shuffle of width 16).  Returns the number of pairs of this half.
template <typename KeyT, int E = kHalfE>
__device__ __forceinline__ int half_rle_store(const KeyT (&key)[E], int nvalid, KeyT* __restrict__ keys_out,
                                              uint32_t* __restrict__ counts_out, KeyT* __restrict__ stage_k,
                                              uint32_t* __restrict__ stage_c)
{
    const int hl = threadIdx.x & 15;
    const int vpos0 = hl * E;
    const int nlive = min(E, max(0, nvalid - vpos0));
    const uint32_t live = (1u << nlive) - 1u;
    KeyT prev = __shfl_up_sync(0xffffffffu, key[E - 1], 1, 16);
    uint32_t heads = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const bool head = (live >> e & 1u) && ((e == 0 && hl == 0) || key[e] != prev);
        heads |= head ? (1u << e) : 0u;
        prev = key[e];
    }
    // common case: every key of BOTH reads is distinct -> element v goes to slot v with count 1
    if (__all_sync(0xffffffffu, heads == live)) {
#pragma unroll
        for (int e = 0; e < E; e++) stage_k[vpos0 + e] = key[e];      // odd stride in words: conflict-free
        __syncwarp();
        for (int i = hl; i < nvalid; i += 16) {
            keys_out[i] = stage_k[i];
            counts_out[i] = 1u;
        }
        __syncwarp();
        return nvalid;
    }
    const int hc = __popc(heads);
    int inc = hc;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d, 16);
        if (hl >= d) inc += o;
    }
    int off = inc - hc;
    const int total = __shfl_sync(0xffffffffu, inc, 15, 16);
    // position of the next head after this lane: suffix-min of every lane's first head
    int first = heads ? vpos0 + (__ffs(heads) - 1) : nvalid;
    int nxt = __shfl_down_sync(0xffffffffu, first, 1, 16);
    if (hl == 15) nxt = nvalid;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const int o = __shfl_down_sync(0xffffffffu, nxt, d, 16);
        if (hl + d < 16) nxt = min(nxt, o);
    }
    nxt = min(nxt, nvalid);
    uint32_t cnt[E];
#pragma unroll
    for (int e = E - 1; e >= 0; e--) {
        cnt[e] = (uint32_t)(nxt - (vpos0 + e));
        if (heads & (1u << e)) nxt = vpos0 + e;
    }
#pragma unroll
    for (int e = 0; e < E; e++) {
        if (heads & (1u << e)) {
            stage_k[off] = key[e];
            stage_c[off] = cnt[e];
            off++;
        }
    }
    __syncwarp();
    for (int i = hl; i < total; i += 16) {
        keys_out[i] = stage_k[i];
        counts_out[i] = stage_c[i];
    }
    __syncwarp();
    return total;
}

template <typename KeyT, int FMT>
__global__ void __launch_bounds__(kSparseWarps * 32) sparse_half_kernel(
    const BasesRef src, const int64_t* __restrict__ start,
    const int32_t* __restrict__ length, int64_t nS, int k, const int64_t* __restrict__ row_begin,
    int32_t* __restrict__ row_count, KeyT* __restrict__ keys, uint32_t* __restrict__ counts)
{
    constexpr int WARPS = kSparseWarps;
    __shared__ uint32_t s_cw[WARPS][2][kHalfBlocks];
    __shared__ __align__(4) uint16_t s_vh[WARPS][2][kHalfBlocks + 4];
    __shared__ __align__(16) KeyT s_stage_k[WARPS][2][kHalfSlots];
    __shared__ uint32_t s_stage_c[WARPS][2][kHalfSlots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, hl = lane & 15;
    WarpStream st{s_cw[warp][h], s_vh[warp][h]};
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    const int64_t npairs = (nS + 1) >> 1;
    for (int64_t pair = (int64_t)blockIdx.x * WARPS + warp; pair < npairs; pair += nwarps) {
        const int64_t r = 2 * pair + h;
        const bool in = r < nS;
        const int len = in ? length[r] : 0;
        const int nwin = len - k + 1;
        if (in && nwin <= 0 && hl == 0) row_count[r] = 0;
        const bool mine = in && nwin > 0 && nwin <= kHalfSlots;     // longer reads: the other classes
        if (!__any_sync(0xffffffffu, mine)) continue;
        const int64_t s = mine ? start[r] : 0;
        const int a = (int)(s & 15);
        const int nblocks = mine ? (a + len + 15) >> 4 : 0;          // <= 12 for len <= 143 + 31
        __syncwarp();
        {
            uint32_t c = 0, v = 0;
            if (hl < nblocks) {
                const uint4 raw = load_block<FMT>(src, (s >> 4) + hl);
                decode_block<FMT>(raw, c, v);
                v &= from_pos(max(0, a - 16 * hl)) & ~from_pos(min(16, max(0, a + len - 16 * hl)));
            }
            st.cw[hl] = c;
            st.vh[hl ^ 1] = (uint16_t)v;
            if (hl < 4) st.vh[kHalfBlocks + hl] = 0;
        }
        __syncwarp();
        KeyT key[kHalfE];
        uint32_t valid = extract_windows<KeyT, kHalfE>(st, a + hl * kHalfE, k, key);
        if (!mine) {
            valid = 0;
#pragma unroll
            for (int e = 0; e < kHalfE; e++) key[e] = KeyMax<KeyT>::value;
        }
        int nvalid = __popc(valid);
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, d);
        half_sort<KeyT, kHalfE>(key);
        const int64_t rb = mine ? row_begin[r] : 0;
        const int nd = half_rle_store<KeyT, kHalfE>(key, nvalid, keys + rb, counts + rb, s_stage_k[warp][h], s_stage_c[warp][h]);
        if (mine && hl == 0) row_count[r] = nd;
    }
}

// ------------------------------------------------------------------------------------------
// long reads (> 512 windows): MSD partition + one warp sort per bucket
//
//   A row of n windows is split by the top B bits of its keys into 2^B buckets (B chosen per row so
//   that uniform keys give <= 192 per bucket).  Two passes over the BASES do the split (keys are
//   never stored unsorted more than once): pass 1 counts the buckets, pass 2 scatters the keys
//   behind the scanned counts.  Every bucket of <= 512 keys is then sorted by ONE WARP with the
//   same register network as the short reads and run-length encoded in place; buckets are disjoint
//   key ranges, so their (key, count) runs are final and only have to be moved behind each other
//   (scan of the distinct counts + copy).  Buckets above 512 keys (skewed or low-complexity rows)
//   take the segmented LSD radix sort below.  Rows are processed in batches of <= 128 Mi windows,
//   which bounds the scratch to 12 bytes per window of one batch.
constexpr int kBucketTarget = 192;       // mean keys per bucket at most this (uniform keys): with 256 the buckets of a
                                         // row whose mean lands just below 256 straddle the 256-slot network size and
                                         // half of them pay for the 512-slot one
constexpr int kBucketCap = 512;          // what one warp sorts in registers (E = 16)
constexpr int kMaxBucketBits = 22;
constexpr int kPartTile = 4096;          // windows per tile: the unit in which partition work is cut
constexpr int64_t kDefaultBatchKeys = (int64_t)128 << 20;
constexpr int kSortTile = 2048;          // keys per warp tile of the fallback radix sort
constexpr int kSortWarps = 8;

__host__ __device__ __forceinline__ int bucket_bits(int64_t nwin, int k)
{
    int b = 0;
    while (b < kMaxBucketBits && ((int64_t)kBucketTarget << b) < nwin) b++;
    return b < 2 * k ? b : 2 * k;
}

// A row is NARROW when the key bits below its bucket digit fit 32 bits: all keys of a bucket share
// the digit, so only these 32-bit suffixes are stored, sorted and run-length encoded, and the digit
// is put back when the pairs move to the output (uint64 keys with k <= ~23 on Mbp rows: half the
// scratch traffic and a 32-bit sorting network).
__host__ __device__ __forceinline__ bool narrow_row(int64_t nwin, int k) { return 2 * k - bucket_bits(nwin, k) <= 32; }

// ------------------------------------------------------------------------------------------
// medium rows (513..4096 windows): ONE CTA per row, everything in shared memory.
//   The row's bases are encoded into a CTA-wide stream, every thread extracts 16 consecutive windows,
//   the keys are split by their top 3..5 bits into groups of <= 128 keys on average (ranks from
//   shared atomics, tight packing after a scan of the group counts), every group is sorted and
//   run-length encoded by one warp with the short-read machinery (including its 3-bit sub-groups), the
//   pairs stay in shared memory until the distinct counts of all groups are known and then leave with
//   coalesced stores.  HBM traffic = the bases in, the pairs out; no scratch, no global atomics.
//   A row with a group above 256 keys (skew) is appended to the long-row lists instead.
constexpr int kMedMaxWindows = 4096;
constexpr int kMedThreads = 256;
constexpr int kMedW = kMedMaxWindows / kMedThreads;      // 16 consecutive windows per thread
constexpr int kMedBlocks = (15 + kMedMaxWindows + 30 + 15) / 16 + 3;
constexpr int kMedGroupMean = 128;     // keys per group at most, on average
constexpr int kMedMaxGroups = kMedMaxWindows / kMedGroupMean;   // 32

__host__ __device__ __forceinline__ int medium_group_bits(int nwin)
{
    int b = 0;
    while ((kMedGroupMean << b) < nwin) b++;
    return b;
}
__host__ __device__ __forceinline__ bool medium_row(int64_t nwin, int k)
{
    return nwin > kShortMaxWindows && nwin <= kMedMaxWindows && 2 * k >= medium_group_bits((int)nwin);
}

// rows above 512 windows: medium rows into medium_rows[]; long rows into long_rows[cap] from both ends
// (narrow rows from the front, wide ones from the back).  n[0] narrow, n[1] wide, n[2] medium.
__global__ void collect_long_kernel(const int32_t* __restrict__ length, int64_t nS, int k, bool split, bool use_medium,
                                    int64_t* __restrict__ long_rows, int64_t* __restrict__ medium_rows,
                                    unsigned long long* __restrict__ n, int64_t cap)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nS; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nwin = (int64_t)length[r] - k + 1;
        if (nwin > kShortMaxWindows) {
            if (use_medium && medium_row(nwin, k)) {
                const unsigned long long slot = atomicAdd(n + 2, 1ull);
                if ((int64_t)slot < cap) medium_rows[slot] = r;
            } else {
                const bool wide = split && !narrow_row(nwin, k);
                const unsigned long long slot = atomicAdd(n + (wide ? 1 : 0), 1ull);
                if ((int64_t)slot < cap) long_rows[wide ? cap - 1 - (int64_t)slot : (int64_t)slot] = r;
            }
        }
    }
}

template <typename KeyT, int FMT>
__global__ void __launch_bounds__(kMedThreads, sizeof(KeyT) == 8 ? 3 : 4) sparse_medium_kernel(
    const BasesRef src, const int64_t* __restrict__ start, const int32_t* __restrict__ length, int k,
    const int64_t* __restrict__ medium_rows, int64_t n_medium, const int64_t* __restrict__ row_begin,
    int32_t* __restrict__ row_count, KeyT* __restrict__ keys, uint32_t* __restrict__ counts, bool split,
    int64_t* __restrict__ long_rows, unsigned long long* __restrict__ n_long, int64_t cap)
{
    constexpr int T = kMedThreads, W = kMedW, WARPS = T / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    KeyT* s_keys = reinterpret_cast<KeyT*>(smem);                                   // [4096] keys by group, then pairs
    KeyT* s_stage_k = s_keys + kMedMaxWindows;                                      // [WARPS][256]
    uint32_t* s_oc = reinterpret_cast<uint32_t*>(s_stage_k + WARPS * 256);          // [4096] counts of the pairs
    uint32_t* s_stage_c = s_oc + kMedMaxWindows;                                    // [WARPS][256]
    __shared__ uint32_t s_cw[kMedBlocks];
    __shared__ __align__(4) uint16_t s_vh[2 * ((kMedBlocks + 1) / 2) + 2];
    __shared__ uint32_t s_gcnt[kMedMaxGroups], s_gstart[kMedMaxGroups + 1], s_gdist[kMedMaxGroups], s_gdst[kMedMaxGroups + 1];
    __shared__ int s_overflow;
    WarpStream st{s_cw, s_vh};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t j = blockIdx.x; j < n_medium; j += gridDim.x) {
        const int64_t r = medium_rows[j];
        const int64_t s = start[r];
        const int len = length[r];
        const int nwin = len - k + 1;
        const int gbits = medium_group_bits(nwin), ngroups = 1 << gbits;
        const int a = (int)(s & 15);
        const int nblocks = (a + len + 15) >> 4;
        __syncthreads();   // the previous row is out of shared memory
        for (int b = threadIdx.x; b < kMedBlocks; b += T) {
            uint32_t c = 0, v = 0;
            if (b < nblocks) {
                decode_block<FMT>(load_block<FMT>(src, (s >> 4) + b), c, v);
                v &= from_pos(max(0, a - 16 * b)) & ~from_pos(min(16, max(0, a + len - 16 * b)));
            }
            st.cw[b] = c;
            st.vh[b ^ 1] = (uint16_t)v;
        }
        if (threadIdx.x < kMedMaxGroups) s_gcnt[threadIdx.x] = 0u;
        __syncthreads();
        // 1. keys and their rank inside their group
        KeyT key[W];
        uint32_t rank[W];
        const uint32_t valid = extract_windows<KeyT, W>(st, a + (int)threadIdx.x * W, k, key);
        const int gsh = 2 * k - gbits;
#pragma unroll
        for (int e = 0; e < W; e++) {
            rank[e] = 0;
            if (valid >> e & 1u) rank[e] = atomicAdd(&s_gcnt[(uint32_t)(key[e] >> gsh)], 1u);
        }
        __syncthreads();
        // 2. group starts; a group above 256 keys does not fit a warp's network
        if (warp == 0) {
            const uint32_t c = lane < ngroups ? s_gcnt[lane] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            s_gstart[lane] = inc - c;
            const bool ovf = __any_sync(0xffffffffu, c > 256u);
            if (lane == 0) s_overflow = ovf ? 1 : 0;
        }
        __syncthreads();
        if (s_overflow) {   // skewed row: the long-row path takes it
            if (threadIdx.x == 0) {
                const bool wide = split && !narrow_row(nwin, k);
                const unsigned long long slot = atomicAdd(n_long + (wide ? 1 : 0), 1ull);
                if ((int64_t)slot < cap) long_rows[wide ? cap - 1 - (int64_t)slot : (int64_t)slot] = r;
            }
            continue;
        }
#pragma unroll
        for (int e = 0; e < W; e++)
            if (valid >> e & 1u) s_keys[s_gstart[(uint32_t)(key[e] >> gsh)] + rank[e]] = key[e];
        __syncthreads();
        // 3. one warp per group: sort + RLE, pairs back to the group's place in shared memory
        for (int g = warp; g < ngroups; g += WARPS) {
            const int ng = (int)s_gcnt[g];
            const int g0 = (int)s_gstart[g];
            KeyT gk[8];
            const int nlive = min(8, max(0, ng - lane * 8));
#pragma unroll
            for (int e = 0; e < 8; e++) gk[e] = e < nlive ? s_keys[g0 + lane * 8 + e] : KeyMax<KeyT>::value;
            __syncwarp();
            const int sub = gsh - 3;     // sub-groups by the 3 bits below the group digit
            const int nd = warp_sort_rle<KeyT, 8>(gk, (1u << nlive) - 1u, ng <= kGroupedMaxWindows && sub >= 0 ? sub : -1,
                                                  s_keys + g0, s_oc + g0, s_stage_k + warp * 256, s_stage_c + warp * 256);
            if (lane == 0) s_gdist[g] = (uint32_t)nd;
        }
        __syncthreads();
        // 4. distinct counts -> place of every group's pairs in the row
        if (warp == 0) {
            const uint32_t c = lane < ngroups ? s_gdist[lane] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            s_gdst[lane] = inc - c;
            if (lane == 31) { s_gdst[32] = inc; row_count[r] = (int32_t)inc; }
        }
        __syncthreads();
        KeyT* ko = keys + row_begin[r];
        uint32_t* co = counts + row_begin[r];
        for (int g = warp; g < ngroups; g += WARPS) {
            const int nd = (int)s_gdist[g], g0 = (int)s_gstart[g], d0 = (int)s_gdst[g];
            for (int i = lane; i < nd; i += 32) {
                ko[d0 + i] = s_keys[g0 + i];
                co[d0 + i] = s_oc[g0 + i];
            }
        }
    }
}

// A row's partition work is cut into SLABS of up to 64 consecutive tiles; a CTA owns one slab.  In
// the counting pass it keeps the row's bucket counters PRIVATE in shared memory, so the per-key
// atomics never leave the SM, and adds them to the row's counters in HBM once per slab.  Rows with
// more buckets than fit shared memory (> 8 Mi windows) are one slab and count with global atomics.
// The scatter pass keeps ONE cursor per bucket in HBM (L2): a bucket then fills front to back, its
// write frontier is one 32-byte sector, and the frontiers of all rows in flight (26 rows x 32768
// buckets x 32 B = 27 MB) stay L2-resident so that partial sector writes merge there.  Private
// per-slab cursors were measured: 20x as many frontiers (545 MB), scatter 1.7 -> 3.1 ms.
constexpr int kSlabTiles = 64;
constexpr int kSmemBuckets = 32768;      // 128 KiB of uint32 counters
__host__ __device__ __forceinline__ int64_t row_slabs(int64_t ntiles, int64_t nbuckets)
{
    return nbuckets > kSmemBuckets ? 1 : (ntiles + kSlabTiles - 1) / kSlabTiles;
}

// per long row j: windows, buckets, partition tiles and slabs (inputs of four exclusive scans)
__global__ void long_sizes_kernel(const int64_t* __restrict__ long_rows, int64_t n_long, const int32_t* __restrict__ length,
                                  int k, int64_t* __restrict__ loff, int64_t* __restrict__ boff, int64_t* __restrict__ ptile,
                                  int64_t* __restrict__ uoff)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= n_long) {
        const int64_t nwin = j < n_long ? length[long_rows[j]] - k + 1 : 0;
        const int64_t nb = j < n_long ? (int64_t)1 << bucket_bits(nwin, k) : 0;
        const int64_t nt = (nwin + kPartTile - 1) / kPartTile;
        const int64_t ns = j < n_long ? row_slabs(nt, nb) : 0;
        loff[j] = nwin;
        boff[j] = nb;
        ptile[j] = nt;
        uoff[j] = ns;
    }
}

// last j in [lo, hi) with scan[j] <= t (scan strictly increasing, scan[lo] <= t < scan[hi])
__device__ __forceinline__ int64_t locate(const int64_t* __restrict__ scan, int64_t lo, int64_t hi, int64_t t)
{
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (scan[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

// Partition pass over the bases of the batch rows [j0, j1).  Work unit = one slab (see above) of one
// row, walked in steps of 16*T windows: the step's bases are encoded once into a shared-memory bit
// stream (as for the short reads; the raw blocks of the next step are already in flight), every
// thread extracts 16 consecutive windows from registers.  T = 512 for the counting pass when the
// counters leave room for one CTA per SM only, T = 256 otherwise.
//   SCATTER = false: count the buckets of the row into bucket[] (through shared memory).
//   SCATTER = true : bucket[] holds the scanned counts; a key takes the next slot of its bucket
//                    (afterwards bucket[b] = end of bucket b in the scratch).
template <typename KeyT, typename SortT, int FMT, bool SCATTER, int T, int ILP = 1>
__global__ void __launch_bounds__(T, SCATTER ? (ILP > 1 ? 4 : 5) : 1) partition_kernel(
    const BasesRef src, const int64_t* __restrict__ start, const int32_t* __restrict__ length, int k,
    const int64_t* __restrict__ long_rows, const int64_t* __restrict__ boff, const int64_t* __restrict__ ptile,
    const int64_t* __restrict__ uoff, int64_t j0, int64_t j1, unsigned long long* __restrict__ bucket,
    SortT* __restrict__ scratch)
{
    extern __shared__ uint32_t s_cnt[];   // counting pass: private counters of the slab's row
    constexpr int W = 16;                 // consecutive windows per thread
    constexpr int STEP = W * T;           // windows per step
    constexpr int kPartBlocks = (15 + STEP + 30 + 15) / 16 + 3;
    constexpr int NB = (kPartBlocks + T - 1) / T;   // raw blocks per thread and step
    __shared__ uint32_t s_cw[kPartBlocks];
    __shared__ __align__(4) uint16_t s_vh[2 * ((kPartBlocks + 1) / 2) + 2];
    WarpStream st{s_cw, s_vh};
    const int64_t u0 = uoff[j0], nunits = uoff[j1] - u0, b0 = boff[j0];
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int64_t j = locate(uoff, j0, j1, u0 + u);
        const int64_t slab = u0 + u - uoff[j], S = uoff[j + 1] - uoff[j];
        const int nb = (int)(boff[j + 1] - boff[j]);
        const int64_t ntiles_row = ptile[j + 1] - ptile[j];
        const int64_t per = (ntiles_row + S - 1) / S;
        const int64_t r = long_rows[j];
        const int64_t s = start[r];
        const int len = length[r];
        const int nwin = len - k + 1;
        const int64_t wbeg = slab * per * kPartTile, wend = min((int64_t)nwin, wbeg + per * kPartTile);   // the slab's windows
        const int shift = 2 * k - bucket_bits(nwin, k);
        unsigned long long* bk = bucket + (boff[j] - b0);
        const bool priv = !SCATTER && nb <= kSmemBuckets;
        if (priv) {
            __syncthreads();   // the previous unit has flushed its counters
            for (int b = threadIdx.x; b < nb; b += T) s_cnt[b] = 0u;
        }
        // a step covers windows [w0, w0 + STEP): stream position 0 = first byte of the 16-byte block that
        // holds the step's first base; consecutive steps are STEP/16 blocks apart
        const int a = (int)((s + wbeg) & 15);
        int64_t blk0 = (s + wbeg) >> 4;
        const int64_t blk_end = (s + len + 15) >> 4;       // blocks of this read end here
        uint4 raw[NB];
#pragma unroll
        for (int q = 0; q < NB; q++) {
            const int64_t b = blk0 + threadIdx.x + q * T;
            raw[q] = threadIdx.x + q * T < kPartBlocks && b < blk_end ? load_block<FMT>(src, b) : make_uint4(0, 0, 0, 0);
        }
        for (int64_t w0 = wbeg; w0 < wend; w0 += STEP) {
            const int64_t rs = s - blk0 * 16, re = s + len - blk0 * 16;   // the read in stream coordinates
            __syncthreads();   // the previous tile has been read (and the counters are cleared)
#pragma unroll
            for (int q = 0; q < NB; q++) {
                const int b = threadIdx.x + q * T;
                if (b < kPartBlocks) {
                    uint32_t cc, v;
                    decode_block<FMT>(raw[q], cc, v);
                    const int lo = (int)max((int64_t)0, min((int64_t)16, rs - 16 * b));
                    const int hi = (int)max((int64_t)0, min((int64_t)16, re - 16 * b));
                    v &= from_pos(lo) & ~from_pos(hi);
                    st.cw[b] = cc;
                    st.vh[b ^ 1] = (uint16_t)v;
                }
            }
            __syncthreads();
            blk0 += STEP / 16;
            if (w0 + STEP < wend) {
#pragma unroll
                for (int q = 0; q < NB; q++) {
                    const int64_t b = blk0 + threadIdx.x + q * T;
                    raw[q] = threadIdx.x + q * T < kPartBlocks && b < blk_end ? load_block<FMT>(src, b) : make_uint4(0, 0, 0, 0);
                }
            }
            KeyT key[W];
            uint32_t valid = extract_windows<KeyT, W>(st, a + (int)threadIdx.x * W, k, key);
            if (w0 + (int64_t)threadIdx.x * W >= wend) valid = 0;   // the next slab's windows (slabs end at multiples of 16)
            if (!SCATTER) {
#pragma unroll
                for (int e = 0; e < W; e++) {
                    if (valid >> e & 1u) {
                        const uint32_t d = (uint32_t)((uint64_t)key[e] >> shift);
                        if (priv) atomicAdd(&s_cnt[d], 1u); else atomicAdd(&bk[d], 1ull);
                    }
                }
            } else {
                // slots of ILP keys first, then their stores.  Measured: ILP = 8 (64 registers, 4 CTAs/SM) 1.87 ms
                // per 128 Mi keys, ILP = 1 (48 registers, 5 CTAs/SM) 1.71 ms: the pass is bound by L2 operations
                // (one atomic + one sector write per key, ~150 G/s), not by latency
#pragma unroll
                for (int h = 0; h < W; h += ILP) {
                    unsigned long long pos[ILP];
#pragma unroll
                    for (int e = 0; e < ILP; e++) {
                        pos[e] = 0;
                        if (valid >> (h + e) & 1u) pos[e] = atomicAdd(&bk[(uint32_t)((uint64_t)key[h + e] >> shift)], 1ull);
                    }
#pragma unroll
                    for (int e = 0; e < ILP; e++)
                        if (valid >> (h + e) & 1u) scratch[pos[e]] = (SortT)key[h + e];   // narrow rows: the suffix (the digit is cut off)
                }
            }
        }
        if (priv) {
            __syncthreads();
            for (int b = threadIdx.x; b < nb; b += T) {
                const uint32_t n = s_cnt[b];
                if (n) atomicAdd(&bk[b], (unsigned long long)n);
            }
        }
    }
}

// one warp per bucket of <= 32*E keys: load, sort in registers, run-length encode IN PLACE (the
// pairs of bucket b start where its keys started).  The E = 4 instance also lists the buckets that
// are too large for a warp.
template <typename KeyT, int E>
__global__ void __launch_bounds__(SparseCta<E>::WARPS * 32) bucket_sort_kernel(
    const unsigned long long* __restrict__ bend, int64_t nb, KeyT* __restrict__ scratch, uint32_t* __restrict__ pc,
    unsigned long long* __restrict__ distinct, int64_t* __restrict__ fb_list, unsigned long long* __restrict__ n_fb,
    const int64_t* __restrict__ boff, int64_t j0, int64_t j1, const int64_t* __restrict__ long_rows,
    const int32_t* __restrict__ length, int k, int min_keys, bool list_oversized)
{
    constexpr int WARPS = SparseCta<E>::WARPS;
    __shared__ __align__(16) KeyT s_stage_k[WARPS][32 * E];
    __shared__ uint32_t s_stage_c[WARPS][32 * E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = (int64_t)blockIdx.x * WARPS + warp; b < nb; b += (int64_t)gridDim.x * WARPS) {
        const int64_t beg = b ? (int64_t)bend[b - 1] : 0;
        const int64_t n64 = (int64_t)bend[b] - beg;
        if (list_oversized && n64 > kBucketCap && lane == 0) fb_list[atomicAdd(n_fb, 1ull)] = b;
        if (n64 == 0 || n64 > 32 * E || n64 < min_keys) continue;
        const int n = (int)n64;
        KeyT key[E];
        uint32_t valid = 0;
#pragma unroll
        for (int e = 0; e < E; e++) {      // any placement will do: the network sorts it
            const int g = e * 32 + lane;
            key[e] = g < n ? scratch[beg + g] : KeyMax<KeyT>::value;
            valid |= g < n ? 1u << e : 0u;
        }
        // 64-bit keys, buckets of <= 160 keys: sub-groups by the 3 key bits below the bucket digit (the row's
        // shift).  Measured on 5 Mbp rows: k=31 (uint64 sort keys) 24.6-25.5 -> 26.1 Gbases/s; with 32-bit sort
        // keys the split and the row lookup cost more than the smaller networks save (sort phase 1.70 -> 1.87 ms)
        int gshift = -1;
        if (E == 8 && sizeof(KeyT) == 8 && n <= kGroupedMaxWindows) {
            const int64_t j = locate(boff, j0, j1, boff[j0] + b);
            gshift = 2 * k - bucket_bits((int64_t)length[long_rows[j]] - k + 1, k) - 3;
        }
        __syncwarp();   // every lane has its keys: the bucket's place is rewritten below
        const int nd = warp_sort_rle<KeyT, E>(key, valid, gshift, scratch + beg, pc + beg, s_stage_k[warp], s_stage_c[warp]);
        if (lane == 0) distinct[b] = (unsigned long long)nd;
    }
}

// Buckets of <= 176 keys (97 % of them on uniform keys: the mean is ~150): TWO buckets per warp, 16 lanes x 11
// keys each, the network of the short reads with an 11-input lane network (35 comparators).  Also lists the
// buckets that are too large for a warp.
constexpr int kHalfBucketE = 11;
constexpr int kHalfBucketSlots = 16 * kHalfBucketE;   // 176

template <typename KeyT>
__global__ void __launch_bounds__(kSparseWarps * 32) bucket_sort_half_kernel(
    const unsigned long long* __restrict__ bend, int64_t nb, KeyT* __restrict__ scratch, uint32_t* __restrict__ pc,
    unsigned long long* __restrict__ distinct, int64_t* __restrict__ fb_list, unsigned long long* __restrict__ n_fb)
{
    constexpr int WARPS = kSparseWarps, E = kHalfBucketE;
    __shared__ __align__(16) KeyT s_stage_k[WARPS][2][kHalfBucketSlots];
    __shared__ uint32_t s_stage_c[WARPS][2][kHalfBucketSlots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, hl = lane & 15;
    const int64_t npairs = (nb + 1) >> 1;
    for (int64_t pair = (int64_t)blockIdx.x * WARPS + warp; pair < npairs; pair += (int64_t)gridDim.x * WARPS) {
        const int64_t b = 2 * pair + h;
        int64_t beg = 0, n64 = 0;
        if (b < nb) {
            beg = b ? (int64_t)bend[b - 1] : 0;
            n64 = (int64_t)bend[b] - beg;
            if (n64 > kBucketCap && hl == 0) fb_list[atomicAdd(n_fb, 1ull)] = b;
        }
        const bool mine = n64 > 0 && n64 <= kHalfBucketSlots;
        if (!__any_sync(0xffffffffu, mine)) continue;
        const int n = mine ? (int)n64 : 0;
        KeyT key[E];
#pragma unroll
        for (int e = 0; e < E; e++) {      // any placement will do: the network sorts it
            const int g = e * 16 + hl;
            key[e] = g < n ? scratch[beg + g] : KeyMax<KeyT>::value;
        }
        __syncwarp();   // every lane has its keys: the bucket's place is rewritten below
        half_sort<KeyT, E>(key);
        const int nd = half_rle_store<KeyT, E>(key, n, scratch + beg, pc + beg, s_stage_k[warp][h], s_stage_c[warp][h]);
        if (mine && hl == 0) distinct[b] = (unsigned long long)nd;
    }
}

// ---- fallback for buckets above 512 keys: segmented LSD radix sort (8-bit digits, one warp per
// 2048-key tile, stable ranks from match.any, per-segment offsets from ONE flat uint32 scan used
// modulo 2^32: segments are < 2^31 keys)
__global__ void fallback_segments_kernel(const int64_t* __restrict__ fb_list, int64_t nfb,
                                         const unsigned long long* __restrict__ bend, int64_t* __restrict__ seg_begin,
                                         int64_t* __restrict__ seg_n, int64_t* __restrict__ stile)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nfb) {
        const int64_t b = fb_list[i];
        const int64_t beg = b ? (int64_t)bend[b - 1] : 0;
        const int64_t n = (int64_t)bend[b] - beg;
        seg_begin[i] = beg;
        seg_n[i] = n;
        stile[i] = (n + kSortTile - 1) / kSortTile;
    } else if (i == nfb) {
        stile[i] = 0;
    }
}

// tile t of the flat tile list -> segment, tile inside the segment, keys of the tile
struct SortTile {
    int64_t j, tin, ntiles, key0;   // key0 = offset of the tile's first key in the scratch arrays
    int n;                          // keys in the tile
};
__device__ __forceinline__ SortTile locate_tile(int64_t t, const int64_t* __restrict__ stile,
                                                const int64_t* __restrict__ seg_begin,
                                                const int64_t* __restrict__ seg_n, int64_t nseg)
{
    SortTile st;
    st.j = locate(stile, 0, nseg, t);
    st.tin = t - stile[st.j];
    st.ntiles = stile[st.j + 1] - stile[st.j];
    st.key0 = seg_begin[st.j] + st.tin * kSortTile;
    st.n = (int)min((int64_t)kSortTile, seg_n[st.j] - st.tin * kSortTile);
    return st;
}

// digit histogram of every tile; counter layout = segment-major, then digit, then tile-in-segment,
// so that ONE flat exclusive scan gives, relative to the segment's first counter, the destination
// of every (digit, tile) group inside its segment
template <typename KeyT>
__global__ void __launch_bounds__(kSortWarps * 32) sort_hist_kernel(const KeyT* __restrict__ src, int shift,
                                                                    const int64_t* __restrict__ stile,
                                                                    const int64_t* __restrict__ seg_begin,
                                                                    const int64_t* __restrict__ seg_n, int64_t nseg,
                                                                    int64_t ntiles_total, uint32_t* __restrict__ counters)
{
    __shared__ uint32_t s_h[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* h = s_h[warp];
    for (int64_t t = (int64_t)blockIdx.x * kSortWarps + warp; t < ntiles_total; t += (int64_t)gridDim.x * kSortWarps) {
        const SortTile st = locate_tile(t, stile, seg_begin, seg_n, nseg);
        for (int i = lane; i < 256; i += 32) h[i] = 0u;
        __syncwarp();
        for (int i = lane; i < st.n; i += 32) atomicAdd(&h[(uint32_t)(src[st.key0 + i] >> shift) & 255u], 1u);
        __syncwarp();
        uint32_t* c = counters + stile[st.j] * 256;
        for (int d = lane; d < 256; d += 32) c[(int64_t)d * st.ntiles + st.tin] = h[d];
        __syncwarp();
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortWarps * 32) sort_scatter_kernel(const KeyT* __restrict__ src, KeyT* __restrict__ dst,
                                                                       int shift, const int64_t* __restrict__ stile,
                                                                       const int64_t* __restrict__ seg_begin,
                                                                       const int64_t* __restrict__ seg_n, int64_t nseg,
                                                                       int64_t ntiles_total,
                                                                       const uint32_t* __restrict__ scanned)
{
    __shared__ uint32_t s_b[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* base = s_b[warp];
    for (int64_t t = (int64_t)blockIdx.x * kSortWarps + warp; t < ntiles_total; t += (int64_t)gridDim.x * kSortWarps) {
        const SortTile st = locate_tile(t, stile, seg_begin, seg_n, nseg);
        const uint32_t* c = scanned + stile[st.j] * 256;
        const uint32_t seg0 = c[0];
        for (int d = lane; d < 256; d += 32) base[d] = c[(int64_t)d * st.ntiles + st.tin] - seg0;
        __syncwarp();
        KeyT* out = dst + seg_begin[st.j];
        for (int i0 = 0; i0 < st.n; i0 += 32) {        // keys in index order: stable
            const int i = i0 + lane;
            const bool live = i < st.n;
            const KeyT key = live ? src[st.key0 + i] : (KeyT)0;
            const uint32_t d = (uint32_t)(key >> shift) & 255u;
            const uint32_t active = __ballot_sync(0xffffffffu, live);
            if (live) {
                const uint32_t peers = __match_any_sync(active, d);
                const uint32_t b = base[d];
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                out[b + rank] = key;
                __syncwarp(active);
                if (rank == 0) base[d] = b + __popc(peers);
            }
            __syncwarp();
        }
    }
}

// sorted keys of one fallback segment -> (key,count) pairs at the start of the segment, one CTA
// per segment.  `sorted` may be the scratch itself: a head is written at or before its own
// position, after the chunk has been read.
template <typename KeyT>
__global__ void __launch_bounds__(256) segment_rle_kernel(const int64_t* __restrict__ fb_list,
                                                          const int64_t* __restrict__ seg_begin,
                                                          const int64_t* __restrict__ seg_n,
                                                          const KeyT* __restrict__ sorted, KeyT* __restrict__ scratch,
                                                          uint32_t* __restrict__ pc, unsigned long long* __restrict__ distinct)
{
    constexpr int T = 256;
    const int64_t base = seg_begin[blockIdx.x];
    const int nvalid = (int)seg_n[blockIdx.x];
    const KeyT* in = sorted + base;
    KeyT* ko = scratch + base;
    uint32_t* co = pc + base;      // pass 1: position of each head; pass 2: run lengths
    __shared__ int s_warp[T / 32];
    int nheads = 0;
    for (int c0 = 0; c0 < nvalid; c0 += T) {
        const int g = c0 + threadIdx.x;
        KeyT v = 0;
        bool head = false;
        if (g < nvalid) {
            v = in[g];
            head = g == 0 || in[g - 1] != v;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < T / 32; w++) {
            const int x = s_warp[w];
            before += w < warp ? x : 0;
            total += x;
        }
        const int pos = nheads + before + __popc(bal & ((1u << lane) - 1u));
        __syncthreads();   // all reads of `in` for this chunk are done (ko may alias it)
        if (head) { ko[pos] = v; co[pos] = (uint32_t)g; }
        nheads += total;
    }
    __syncthreads();
    for (int c0 = 0; c0 < nheads; c0 += T) {
        const int j = c0 + threadIdx.x;
        uint32_t cur = 0, nxt = 0;
        if (j < nheads) {
            cur = co[j];
            nxt = j + 1 < nheads ? co[j + 1] : (uint32_t)nvalid;
        }
        __syncthreads();
        if (j < nheads) co[j] = nxt - cur;
        __syncthreads();
    }
    if (threadIdx.x == 0) distinct[fb_list[blockIdx.x]] = (unsigned long long)nheads;
}

// the pairs of every bucket move behind those of the buckets before it in the same row (dscan =
// exclusive scan of the distinct counts over the batch's buckets); one warp per bucket.  Narrow rows
// get the bucket digit back here.
template <typename KeyT, typename SortT>
__global__ void __launch_bounds__(256) bucket_copy_kernel(const unsigned long long* __restrict__ bend,
                                                          const unsigned long long* __restrict__ dscan, int64_t nb,
                                                          const int64_t* __restrict__ boff, int64_t j0, int64_t j1,
                                                          const int64_t* __restrict__ long_rows,
                                                          const int32_t* __restrict__ length, int k,
                                                          const int64_t* __restrict__ row_begin,
                                                          const SortT* __restrict__ scratch, const uint32_t* __restrict__ pc,
                                                          KeyT* __restrict__ keys, uint32_t* __restrict__ counts,
                                                          int32_t* __restrict__ row_count)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * 8, b0 = boff[j0];
    for (int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < nb; b += nwarps) {
        const int64_t j = locate(boff, j0, j1, b0 + b);
        const int64_t bfirst = boff[j] - b0;
        const int64_t r = long_rows[j];
        if (b == bfirst && lane == 0) row_count[r] = (int32_t)(dscan[boff[j + 1] - b0] - dscan[bfirst]);
        const int64_t nd = (int64_t)(dscan[b + 1] - dscan[b]);
        if (nd == 0) continue;
        KeyT digit = 0;
        if (sizeof(SortT) < sizeof(KeyT)) {
            const int shift = 2 * k - bucket_bits((int64_t)length[r] - k + 1, k);
            digit = (KeyT)(b - bfirst) << shift;
        }
        const int64_t beg = b ? (int64_t)bend[b - 1] : 0;
        const int64_t dst = row_begin[r] + (int64_t)(dscan[b] - dscan[bfirst]);
        for (int64_t i = lane; i < nd; i += 32) {
            keys[dst + i] = (KeyT)scratch[beg + i] | digit;
            counts[dst + i] = pc[beg + i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// scratch of one call: everything is released (stream-ordered) when the object goes out of scope
struct PoolScratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit PoolScratch(cudaStream_t s) : st(s) {}
    ~PoolScratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    template <typename T> cudaError_t get(T** out, size_t n)
    {
        void* p = nullptr;
        const cudaError_t e = cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), st);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return e;
    }
};

template <typename T>
static cudaError_t scan_in_place(T* data, int64_t n, void* tmp, size_t tmp_bytes, cudaStream_t st)
{
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, data, data, n, st);
}

#define CFRK_TRY(x) do { if ((e = (x)) != cudaSuccess) return e; } while (0)

template <typename KeyT, typename SortT, int FMT>
static cudaError_t sparse_long_rows(const BasesRef bases, const int64_t* start, const int32_t* length, int k,
                                    const int64_t* row_begin, int32_t* row_count, KeyT* keys, uint32_t* counts,
                                    const int64_t* long_rows, int64_t nl, int num_sms, SparseTrace& tr, cudaStream_t st)
{
    cudaError_t e;
    PoolScratch pool(st);
    // per-row windows / buckets / partition tiles / slabs / counters, scanned
    int64_t *loff = nullptr, *boff = nullptr, *ptile = nullptr, *uoff = nullptr;
    CFRK_TRY(pool.get(&loff, (size_t)nl + 1));
    CFRK_TRY(pool.get(&boff, (size_t)nl + 1));
    CFRK_TRY(pool.get(&ptile, (size_t)nl + 1));
    CFRK_TRY(pool.get(&uoff, (size_t)nl + 1));
    long_sizes_kernel<<<(unsigned)((nl + 256) / 256), 256, 0, st>>>(long_rows, nl, length, k, loff, boff, ptile, uoff);
    count_launch();
    {
        size_t sb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, sb, loff, loff, nl + 1, st);
        void* stmp = nullptr;
        CFRK_TRY(pool.get(reinterpret_cast<char**>(&stmp), sb));
        CFRK_TRY(scan_in_place(loff, nl + 1, stmp, sb, st));
        CFRK_TRY(scan_in_place(boff, nl + 1, stmp, sb, st));
        CFRK_TRY(scan_in_place(ptile, nl + 1, stmp, sb, st));
        CFRK_TRY(scan_in_place(uoff, nl + 1, stmp, sb, st));
    }
    std::vector<int64_t> h_loff((size_t)nl + 1), h_boff((size_t)nl + 1), h_uoff((size_t)nl + 1);
    cudaMemcpyAsync(h_loff.data(), loff, ((size_t)nl + 1) * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h_boff.data(), boff, ((size_t)nl + 1) * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h_uoff.data(), uoff, ((size_t)nl + 1) * 8, cudaMemcpyDeviceToHost, st);
    CFRK_TRY(cudaStreamSynchronize(st));

    // batches of consecutive long rows, <= batch_keys windows each (a larger row is its own batch)
    std::vector<int64_t> cuts{0};
    int64_t max_keys = 0, max_nb = 0;
    int64_t batch_keys = kDefaultBatchKeys;
    if (const char* ev = getenv("CFRK_SPARSE_BATCH_KEYS")) batch_keys = std::max<int64_t>(1, atoll(ev));   // tests: force many batches
    for (int64_t j0 = 0; j0 < nl;) {
        int64_t j1 = j0 + 1;
        while (j1 < nl && h_loff[(size_t)j1 + 1] - h_loff[(size_t)j0] <= batch_keys) j1++;
        cuts.push_back(j1);
        max_keys = std::max(max_keys, h_loff[(size_t)j1] - h_loff[(size_t)j0]);
        max_nb = std::max(max_nb, h_boff[(size_t)j1] - h_boff[(size_t)j0]);
        j0 = j1;
    }
    SortT* scratch = nullptr;
    uint32_t* pc = nullptr;
    unsigned long long *bucket = nullptr, *distinct = nullptr, *n_fb = nullptr;
    int64_t* fb_list = nullptr;
    void* btmp = nullptr;
    size_t bb = 0;
    CFRK_TRY(pool.get(&scratch, (size_t)max_keys));
    CFRK_TRY(pool.get(&pc, (size_t)max_keys));
    CFRK_TRY(pool.get(&bucket, (size_t)max_nb + 1));
    CFRK_TRY(pool.get(&distinct, (size_t)max_nb + 1));
    CFRK_TRY(pool.get(&fb_list, (size_t)(max_keys / kBucketCap + 1)));
    CFRK_TRY(pool.get(&n_fb, 1));
    cub::DeviceScan::ExclusiveSum(nullptr, bb, bucket, bucket, max_nb + 1, st);
    CFRK_TRY(pool.get(reinterpret_cast<char**>(&btmp), bb));
    const size_t max_dyn = (size_t)kSmemBuckets * 4;
    CFRK_TRY(cudaFuncSetAttribute(partition_kernel<KeyT, SortT, FMT, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    tr.mark("long rows: collect + scratch");

    for (size_t c = 0; c + 1 < cuts.size(); c++) {
        const int64_t j0 = cuts[c], j1 = cuts[c + 1];
        const int64_t nb = h_boff[(size_t)j1] - h_boff[(size_t)j0];
        const int64_t nkeys = h_loff[(size_t)j1] - h_loff[(size_t)j0];
        const int64_t nunits = h_uoff[(size_t)j1] - h_uoff[(size_t)j0];
        // shared-memory counters: as many as the largest row of the batch that keeps them private
        int64_t priv_nb = 1;
        for (int64_t j = j0; j < j1; j++) {
            const int64_t rnb = h_boff[(size_t)j + 1] - h_boff[(size_t)j];
            if (rnb <= kSmemBuckets) priv_nb = std::max(priv_nb, rnb);
        }
        const size_t dyn = (size_t)priv_nb * 4;
        const bool big = dyn > 48 * 1024;   // counters leave room for few CTAs per SM: 512 threads each
        const int ht = big ? 512 : 256;
        const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / ht, (size_t)(220 * 1024) / (dyn + 4096)));
        const unsigned hgrid = (unsigned)std::min<int64_t>(nunits, (int64_t)num_sms * ctas_per_sm);
        const unsigned sgrid = (unsigned)std::min<int64_t>((nkeys + kPartTile - 1) / kPartTile + (j1 - j0), (int64_t)num_sms * 8);
        cudaMemsetAsync(bucket, 0, ((size_t)nb + 1) * 8, st);
        cudaMemsetAsync(distinct, 0, ((size_t)nb + 1) * 8, st);
        cudaMemsetAsync(n_fb, 0, 8, st);
        if (big)
            partition_kernel<KeyT, SortT, FMT, false, 512><<<hgrid, 512, dyn, st>>>(bases, start, length, k, long_rows, boff, ptile, uoff,
                                                                                   j0, j1, bucket, scratch);
        else
            partition_kernel<KeyT, SortT, FMT, false, 256><<<hgrid, 256, dyn, st>>>(bases, start, length, k, long_rows, boff, ptile, uoff,
                                                                                   j0, j1, bucket, scratch);
        count_launch();
        tr.mark("long rows: bucket histogram");
        CFRK_TRY(scan_in_place(bucket, nb + 1, btmp, bb, st));
        // scatter: units = single tiles in row order (ptile as the unit scan), so that few rows are in
        // flight and the write frontiers stay hot in L2
        partition_kernel<KeyT, SortT, FMT, true, 256><<<sgrid, 256, 0, st>>>(bases, start, length, k, long_rows, boff, ptile, ptile,
                                                                            j0, j1, bucket, scratch);
        count_launch();
        tr.mark("long rows: scatter");
        {
            const unsigned g4 = (unsigned)std::min<int64_t>((nb + SparseCta<4>::WARPS - 1) / SparseCta<4>::WARPS, (int64_t)num_sms * 8);
            const unsigned g16 = (unsigned)std::min<int64_t>((nb + SparseCta<16>::WARPS - 1) / SparseCta<16>::WARPS, (int64_t)num_sms * 8);
            static const bool half_buckets = !(getenv("CFRK_SPARSE_HALF") && atoi(getenv("CFRK_SPARSE_HALF")) == 0);
            if (half_buckets) {
                // <= 176 keys: two buckets per warp; 177..256 and 257..512: one warp each
                const unsigned gh = (unsigned)std::min<int64_t>(((nb + 1) / 2 + kSparseWarps - 1) / kSparseWarps, (int64_t)num_sms * 8);
                bucket_sort_half_kernel<SortT><<<gh, kSparseWarps * 32, 0, st>>>(bucket, nb, scratch, pc, distinct, fb_list, n_fb);
                bucket_sort_kernel<SortT, 8><<<g4, SparseCta<8>::WARPS * 32, 0, st>>>(bucket, nb, scratch, pc, distinct, fb_list, n_fb, boff, j0, j1, long_rows, length, k, kHalfBucketSlots + 1, false);
                bucket_sort_kernel<SortT, 16><<<g16, SparseCta<16>::WARPS * 32, 0, st>>>(bucket, nb, scratch, pc, distinct, fb_list, n_fb, boff, j0, j1, long_rows, length, k, 257, false);
            } else {
                bucket_sort_kernel<SortT, 4><<<g4, SparseCta<4>::WARPS * 32, 0, st>>>(bucket, nb, scratch, pc, distinct, fb_list, n_fb, boff, j0, j1, long_rows, length, k, 1, true);
                bucket_sort_kernel<SortT, 8><<<g4, SparseCta<8>::WARPS * 32, 0, st>>>(bucket, nb, scratch, pc, distinct, fb_list, n_fb, boff, j0, j1, long_rows, length, k, 129, false);
                bucket_sort_kernel<SortT, 16><<<g16, SparseCta<16>::WARPS * 32, 0, st>>>(bucket, nb, scratch, pc, distinct, fb_list, n_fb, boff, j0, j1, long_rows, length, k, 257, false);
            }
            count_launch(); count_launch(); count_launch();
        }
        unsigned long long nfb = 0;
        cudaMemcpyAsync(&nfb, n_fb, 8, cudaMemcpyDeviceToHost, st);
        CFRK_TRY(cudaStreamSynchronize(st));
        tr.mark("long rows: bucket sort");
        if (nfb > 0) {
            // buckets above 512 keys: segmented LSD radix sort over all 2k key bits
            PoolScratch fpool(st);
            const int64_t ns = (int64_t)nfb;
            int64_t *seg_begin = nullptr, *seg_n = nullptr, *stile = nullptr;
            SortT* other = nullptr;
            CFRK_TRY(fpool.get(&seg_begin, (size_t)ns));
            CFRK_TRY(fpool.get(&seg_n, (size_t)ns));
            CFRK_TRY(fpool.get(&stile, (size_t)ns + 1));
            CFRK_TRY(fpool.get(&other, (size_t)nkeys));
            fallback_segments_kernel<<<(unsigned)((ns + 256) / 256), 256, 0, st>>>(fb_list, ns, bucket, seg_begin, seg_n, stile);
            count_launch();
            size_t sb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, sb, stile, stile, ns + 1, st);
            void* stmp = nullptr;
            CFRK_TRY(fpool.get(reinterpret_cast<char**>(&stmp), sb));
            CFRK_TRY(scan_in_place(stile, ns + 1, stmp, sb, st));
            int64_t stiles = 0;
            cudaMemcpyAsync(&stiles, stile + ns, 8, cudaMemcpyDeviceToHost, st);
            CFRK_TRY(cudaStreamSynchronize(st));
            uint32_t* counters = nullptr;
            const size_t ncount = (size_t)stiles * 256;
            CFRK_TRY(fpool.get(&counters, ncount));
            size_t cb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, cb, counters, counters, (int64_t)ncount, st);
            void* ctmp = nullptr;
            CFRK_TRY(fpool.get(reinterpret_cast<char**>(&ctmp), cb));
            const unsigned sgrid = (unsigned)std::min<int64_t>((stiles + kSortWarps - 1) / kSortWarps, (int64_t)num_sms * 8);
            const int bits = std::min(2 * k, (int)sizeof(SortT) * 8);
            const int passes = (bits + 7) / 8;
            SortT *src = scratch, *dst = other;
            for (int p = 0; p < passes; p++) {
                sort_hist_kernel<SortT><<<sgrid, kSortWarps * 32, 0, st>>>(src, 8 * p, stile, seg_begin, seg_n, ns, stiles, counters);
                CFRK_TRY(scan_in_place(counters, (int64_t)ncount, ctmp, cb, st));
                sort_scatter_kernel<SortT><<<sgrid, kSortWarps * 32, 0, st>>>(src, dst, 8 * p, stile, seg_begin, seg_n, ns, stiles, counters);
                count_launch(); count_launch();
                std::swap(src, dst);
            }
            segment_rle_kernel<SortT><<<(unsigned)ns, 256, 0, st>>>(fb_list, seg_begin, seg_n, src, scratch, pc, distinct);
            count_launch();
            tr.mark("long rows: oversized buckets (radix sort)");
        }
        CFRK_TRY(scan_in_place(distinct, nb + 1, btmp, bb, st));
        {
            const unsigned cgrid = (unsigned)std::min<int64_t>((nb + 7) / 8, (int64_t)num_sms * 8);
            bucket_copy_kernel<KeyT, SortT><<<cgrid, 256, 0, st>>>(bucket, distinct, nb, boff, j0, j1, long_rows, length, k, row_begin,
                                                                  scratch, pc, keys, counts, row_count);
            count_launch();
        }
        tr.mark("long rows: compaction");
        CFRK_TRY(cudaGetLastError());
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
template <typename KeyT, int FMT>
static cudaError_t sparse_impl(const BasesRef bases, const int64_t* start, const int32_t* length, int64_t nS, int k,
                               int64_t* row_begin, int32_t* row_count, KeyT* keys, uint32_t* counts,
                               int64_t capacity, int64_t* total_windows, cudaStream_t st)
{
    cudaError_t e;
    int dev = 0, num_sms = 148;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);

    SparseTrace tr(st);
    keep_pool_memory(dev);
    // 1. row_begin = exclusive scan of window counts (+ how many reads each size class has)
    PoolScratch lists(st);   // released on every return path
    unsigned long long* d_cls = nullptr;
    if ((e = lists.get(&d_cls, 4)) != cudaSuccess) return e;
    cudaMemsetAsync(d_cls, 0, 32, st);
    nwin_kernel<<<num_sms * 4, 256, 0, st>>>(length, nS, k, row_begin, d_cls);
    count_launch();
    size_t tmp_bytes = 0;
    if ((e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, row_begin, row_begin, nS + 1, st)) != cudaSuccess) return e;
    void* tmp = nullptr;
    if ((e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st)) != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, row_begin, row_begin, nS + 1, st);
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return e;
    int64_t total = 0;
    unsigned long long cls[4] = {0, 0, 0, 0};
    if ((e = cudaMemcpyAsync(&total, row_begin + nS, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(cls, d_cls, 32, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (total_windows) *total_windows = total;
    if (total > capacity) return cudaErrorInvalidValue;   // caller maps to CFRK_EINVAL "capacity"

    tr.mark("row offsets");
    // 2. short reads: one warp per read
    {
        const int64_t ctas = (nS + kSparseWarps - 1) / kSparseWarps;
        const unsigned grid = (unsigned)(ctas < (int64_t)num_sms * 8 ? ctas : (int64_t)num_sms * 8);
        // which classes occur is not known on the host without a pass over the lengths: launch all
        // three; a class without reads costs one pass over length[] (4 B/read)
        const BasesRef b8 = bases;
        const char* gev = getenv("CFRK_SPARSE_GROUPS");          // 0: always the full network (A/B measurements)
        const bool group_split = !(gev && atoi(gev) == 0);
        const char* hev = getenv("CFRK_SPARSE_HALF");            // 0: round-1 classes (one warp per read for every length)
        const bool use_half = !(hev && atoi(hev) == 0);
        if (use_half) {
            // reads of <= 144 windows: two per warp (sparse_half_kernel; it also writes row_count = 0 for reads
            // without a window); 145..256 and 257..512: one warp each, launched only if such reads exist
            const int64_t hctas = ((nS + 1) / 2 + kSparseWarps - 1) / kSparseWarps;
            const unsigned hgrid = (unsigned)(hctas < (int64_t)num_sms * 8 ? hctas : (int64_t)num_sms * 8);
            sparse_half_kernel<KeyT, FMT><<<hgrid, kSparseWarps * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts);
            count_launch();
            if (cls[1]) {
                sparse_short_kernel<KeyT, FMT, 8><<<grid, SparseCta<8>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts, group_split, kHalfSlots + 1);
                count_launch();
            }
            if (cls[2]) {
                sparse_short_kernel<KeyT, FMT, 16><<<grid, SparseCta<16>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts, group_split, 257);
                count_launch();
            }
        } else {
            sparse_short_kernel<KeyT, FMT, 4><<<grid, SparseCta<4>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts, group_split, 1);
            sparse_short_kernel<KeyT, FMT, 8><<<grid, SparseCta<8>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts, group_split, 129);
            sparse_short_kernel<KeyT, FMT, 16><<<grid, SparseCta<16>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts, group_split, 257);
            count_launch(); count_launch(); count_launch();
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }

    tr.mark("short reads");
    // 3. rows above 512 windows: medium rows (one CTA per row, shared memory only), then the long rows --
    // narrow ones (32-bit suffixes) listed from the front, wide ones from the back -- including the
    // medium rows that turned out to be skewed
    int64_t *long_rows = nullptr, *medium_rows = nullptr;
    unsigned long long* d_n = nullptr;
    const int64_t cap_long = total / kShortMaxWindows + 1;   // such a row has > 512 windows
    const char* mev = getenv("CFRK_SPARSE_MEDIUM");          // 0: medium rows take the long-row path (A/B measurements)
    const bool use_medium = !(mev && atoi(mev) == 0);
    if (cls[3] == 0) return cudaGetLastError();               // no read above 512 windows: done
    if ((e = lists.get(&long_rows, (size_t)cap_long)) != cudaSuccess) return e;
    if ((e = lists.get(&medium_rows, (size_t)cap_long)) != cudaSuccess) return e;
    if ((e = lists.get(&d_n, 3)) != cudaSuccess) return e;
    cudaMemsetAsync(d_n, 0, 24, st);
    collect_long_kernel<<<num_sms * 4, 256, 0, st>>>(length, nS, k, sizeof(KeyT) == 8, use_medium, long_rows, medium_rows, d_n, cap_long);
    count_launch();
    unsigned long long n_rows[3] = {0, 0, 0};
    cudaMemcpyAsync(n_rows, d_n, 24, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    const BasesRef b8 = bases;
    if (n_rows[2] > 0) {
        const size_t dyn = (size_t)(kMedMaxWindows + (kMedThreads / 32) * 256) * (sizeof(KeyT) + 4);
        auto kern = sparse_medium_kernel<KeyT, FMT>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)) != cudaSuccess) return e;
        const int64_t nm = (int64_t)n_rows[2];
        kern<<<(unsigned)std::min<int64_t>(nm, (int64_t)num_sms * (sizeof(KeyT) == 8 ? 3 : 4)), kMedThreads, dyn, st>>>(
            b8, start, length, k, medium_rows, nm, row_begin, row_count, keys, counts, sizeof(KeyT) == 8, long_rows, d_n, cap_long);
        count_launch();
        cudaMemcpyAsync(n_rows, d_n, 16, cudaMemcpyDeviceToHost, st);   // skewed medium rows joined the long lists
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        tr.mark("medium rows");
    }
    if (n_rows[0] > 0)
        e = sparse_long_rows<KeyT, uint32_t, FMT>(b8, start, length, k, row_begin, row_count, keys, counts, long_rows,
                                                  (int64_t)n_rows[0], num_sms, tr, st);
    if (e == cudaSuccess && n_rows[1] > 0)
        e = sparse_long_rows<KeyT, KeyT, FMT>(b8, start, length, k, row_begin, row_count, keys, counts,
                                              long_rows + (cap_long - (int64_t)n_rows[1]), (int64_t)n_rows[1], num_sms, tr, st);
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <typename KeyT>
static cudaError_t sparse_fmt(const BasesRef src, int fmt, const int64_t* start, const int32_t* length, int64_t nS, int k,
                              int64_t* row_begin, int32_t* row_count, KeyT* keys, uint32_t* counts, int64_t capacity,
                              int64_t* total_windows, cudaStream_t st)
{
    if (fmt == FMT_PACKED)
        return sparse_impl<KeyT, FMT_PACKED>(src, start, length, nS, k, row_begin, row_count, keys, counts, capacity, total_windows, st);
    return fmt == FMT_ASCII
               ? sparse_impl<KeyT, FMT_ASCII>(src, start, length, nS, k, row_begin, row_count, keys, counts, capacity, total_windows, st)
               : sparse_impl<KeyT, FMT_CODES>(src, start, length, nS, k, row_begin, row_count, keys, counts, capacity, total_windows, st);
}

// fmt 2 (packed): bases = the uint32 codes, packed_valid = the uint16 validity masks of cfrk_encode_2bit_device
cudaError_t launch_sparse(const void* bases, int fmt, const int64_t* start, const int32_t* length, int64_t nS, int k,
                          int64_t* row_begin, int32_t* row_count, void* keys, int key_bytes, uint32_t* counts,
                          int64_t capacity, int64_t* total_windows, cudaStream_t st, const uint16_t* packed_valid)
{
    const BasesRef src{static_cast<const uint8_t*>(bases), packed_valid};
    if (key_bytes == 4)
        return sparse_fmt<uint32_t>(src, fmt, start, length, nS, k, row_begin, row_count, static_cast<uint32_t*>(keys), counts, capacity,
                                    total_windows, st);
    return sparse_fmt<uint64_t>(src, fmt, start, length, nS, k, row_begin, row_count, static_cast<uint64_t*>(keys), counts, capacity,
                                total_windows, st);
}

}  // namespace cfrk
