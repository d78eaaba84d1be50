// sparse.cu -- sparse per-read k-mer counts for k = 1..31 (exact semantics): for each read the
// sorted distinct k-mers and their multiplicities.  No reference counterpart: the reference's dense
// rows stop being usable at k = 9 (4^k int32 per read; int overflow); this is the
// path BASELINE.json configs 3 and 4 name.
//
//   rows     row r owns keys/counts[row_begin[r] .. row_begin[r] + row_count[r]), where
//            row_begin[r] = sum_{j<r} max(0, len_j - k + 1) (every window distinct: worst case), so
//            rows are written independently; callers compact if they want tight CSR.
//   short reads (<= 512 windows): ONE WARP per read.  The read is encoded once into a per-warp
//            bit stream in shared memory (2-bit codes + validity), every lane extracts E windows
//            with funnel shifts, the 32*E keys are sorted by a bitonic network that lives entirely
//            in registers (blocked layout: strides < E are register swaps, the others shfl.xor) and
//            run-length encoded with two warp scans.  Hand-written, no library.
//   long reads: key generation kernel -> hand-written segmented LSD radix sort (8-bit digits, one
//            warp per 2048-key tile, stable ranks from match.any, per-row offsets from ONE flat
//            uint32 scan used modulo 2^32) -> per-row run-length encode.  cub::DeviceScan is the
//            only library call (plumbing).
#include "kernels.h"
#include "kmer_device.cuh"

#include <cub/device/device_scan.cuh>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace cfrk {

extern void count_launch();

// Scratch comes from the stream-ordered pool.  By default the pool gives freed memory back to the
// OS at the next synchronisation, so every call would map its gigabytes of scratch again (measured:
// 1.4 s per call for 3.3 GB); keep it cached instead.
static void keep_pool_memory(int dev)
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

template <typename KeyT> struct KeyMax;
template <> struct KeyMax<uint32_t> { static constexpr uint32_t value = 0xFFFFFFFFu; };
template <> struct KeyMax<uint64_t> { static constexpr uint64_t value = 0xFFFFFFFFFFFFFFFFull; };

// ------------------------------------------------------------------------------------------
// window counts -> row_begin (exclusive scan done by cub::DeviceScan: plumbing)
__global__ void nwin_kernel(const int32_t* __restrict__ length, int64_t nS, int k, int64_t* __restrict__ out)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nS; r += (int64_t)gridDim.x * blockDim.x)
        out[r] = r < nS ? max(0, length[r] - k + 1) : 0;
}

// ------------------------------------------------------------------------------------------
// keep min(mine, other) if take_min else max: min ^ ((mine ^ other) & mask), mask = take_min ? 0 : ~0
template <typename KeyT>
__device__ __forceinline__ KeyT keep_minmax(KeyT mine, KeyT other, KeyT mask)
{
    const KeyT mn = mine < other ? mine : other;
    return mn ^ ((mine ^ other) & mask);
}

// Bitonic network over 32*E keys in blocked layout (lane L holds elements L*E .. L*E+E-1), in the
// direction-free form: every merge starts with a MIRRORED compare (i with i ^ (size-1)) and goes on
// with the usual strides, and every comparator leaves the minimum at the lower index.  In-lane
// comparators are then pure min/max on compile-time registers; cross-lane ones need one per-lane
// mask per stage.
template <typename KeyT, int E>
__device__ __forceinline__ void bitonic_sort_blocked(KeyT (&key)[E])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int size = 2; size <= 32 * E; size <<= 1) {
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
            const KeyT mask = (lane & (size / (2 * E))) == 0 ? (KeyT)0 : ~(KeyT)0;
            KeyT other[E];
#pragma unroll
            for (int e = 0; e < E; e++) other[e] = __shfl_xor_sync(0xffffffffu, key[E - 1 - e], lm);
#pragma unroll
            for (int e = 0; e < E; e++) key[e] = keep_minmax<KeyT>(key[e], other[e], mask);
        }
        // remaining strides
#pragma unroll
        for (int stride = size >> 2; stride >= 1; stride >>= 1) {
            if (stride >= E) {
                const int ls = stride / E;
                const KeyT mask = (lane & ls) == 0 ? (KeyT)0 : ~(KeyT)0;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const KeyT other = __shfl_xor_sync(0xffffffffu, key[e], ls);
                    key[e] = keep_minmax<KeyT>(key[e], other, mask);
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

// keys sorted ascending in blocked layout, the first nvalid of them real -> (key, count) pairs
template <typename KeyT, int E>
__device__ __forceinline__ int warp_rle_store(const KeyT (&key)[E], int nvalid, KeyT* __restrict__ keys_out,
                                              uint32_t* __restrict__ counts_out, KeyT* __restrict__ stage_k,
                                              uint32_t* __restrict__ stage_c)
{
    const int lane = threadIdx.x & 31;
    KeyT prev = __shfl_up_sync(0xffffffffu, key[E - 1], 1);
    uint32_t heads = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int g = lane * E + e;
        const bool head = g < nvalid && (g == 0 || key[e] != prev);
        heads |= head ? (1u << e) : 0u;
        prev = key[e];
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
    // position of the next head after this lane: suffix-min of each lane's first head
    int first = heads ? lane * E + (__ffs(heads) - 1) : nvalid;
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
        const int g = lane * E + e;
        cnt[e] = (uint32_t)(nxt - g);
        if (heads & (1u << e)) nxt = g;
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

struct WarpStream {
    uint32_t* cw;   // 2-bit codes, 16 bases per word, first base in the top bits
    uint16_t* vh;   // validity, 16 bases per half-word; stored so that 32-bit loads see 32 positions MSB-first
};

// window of k bases starting at stream position P
template <typename KeyT>
__device__ __forceinline__ bool stream_window(const WarpStream& st, int P, int k, KeyT& key)
{
    const int b = P >> 4, o = (P & 15) * 2;
    const uint32_t w0 = st.cw[b], w1 = st.cw[b + 1];
    const uint32_t hi = __funnelshift_l(w1, w0, o);
    if (sizeof(KeyT) == 4) {
        key = (KeyT)(hi >> (32 - 2 * k));
    } else {
        const uint32_t lo = __funnelshift_l(st.cw[b + 2], w1, o);
        key = (KeyT)((((uint64_t)hi << 32) | lo) >> (64 - 2 * k));
    }
    const uint32_t* vw = reinterpret_cast<const uint32_t*>(st.vh);
    const int c = P >> 5, vo = P & 31;
    const uint32_t v = __funnelshift_l(vw[c + 1], vw[c], vo);
    const uint32_t need = 0xFFFFFFFFu << (32 - k);
    return (v & need) == need;
}

template <typename KeyT, int E>
__device__ __forceinline__ int warp_count_read(const WarpStream& st, int a, int nwin, int k,
                                               KeyT* __restrict__ keys_out, uint32_t* __restrict__ counts_out,
                                               KeyT* __restrict__ stage_k, uint32_t* __restrict__ stage_c)
{
    const int lane = threadIdx.x & 31;
    KeyT key[E];
    int myvalid = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int g = lane * E + e;
        KeyT kk = KeyMax<KeyT>::value;
        if (g < nwin) {
            KeyT t;
            if (stream_window<KeyT>(st, a + g, k, t)) { kk = t; myvalid++; }
        }
        key[e] = kk;
    }
    int nvalid = myvalid;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, d);
    bitonic_sort_blocked<KeyT, E>(key);
    return warp_rle_store<KeyT, E>(key, nvalid, keys_out, counts_out, stage_k, stage_c);
}

// E = keys per lane: the kernel handles the reads whose window count falls in its class
// ((E == 4: 1..128, E == 8: 129..256, E == 16: 257..512) and leaves the others to its siblings, so
// that the common short-read case is not compiled with the register budget of the largest network.
template <int E> struct SparseCta { static constexpr int WARPS = E == 16 ? 4 : kSparseWarps; };

template <typename KeyT, int FMT, int E>
__global__ void __launch_bounds__(SparseCta<E>::WARPS * 32) sparse_short_kernel(
    const uint8_t* __restrict__ bases, const int64_t* __restrict__ start, const int32_t* __restrict__ length,
    int64_t nS, int k, const int64_t* __restrict__ row_begin, int32_t* __restrict__ row_count,
    KeyT* __restrict__ keys, uint32_t* __restrict__ counts)
{
    constexpr int WARPS = SparseCta<E>::WARPS;
    __shared__ uint32_t s_cw[WARPS][kStreamBlocks];
    __shared__ __align__(4) uint16_t s_vh[WARPS][2 * ((kStreamBlocks + 1) / 2) + 2];
    __shared__ KeyT s_stage_k[WARPS][32 * E];
    __shared__ uint32_t s_stage_c[WARPS][32 * E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStream st{s_cw[warp], s_vh[warp]};
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < nS; r += nwarps) {
        const int len = length[r];
        const int nwin = len - k + 1;
        if (nwin <= 0) { if (E == 4 && lane == 0) row_count[r] = 0; continue; }
        if (nwin > 32 * E || (E > 4 && nwin <= 16 * E)) continue;  // another class, or the long path
        const int64_t s = start[r];
        const int a = (int)(s & 15);
        const int nblocks = (a + len + 15) >> 4;
        __syncwarp();
        for (int b = lane; b < kStreamBlocks; b += 32) {
            uint32_t c = 0, v = 0;
            if (b < nblocks) {
                encode16<FMT>(ld_block(bases + ((s >> 4) + b) * 16), c, v);
                v &= from_pos(max(0, a - 16 * b)) & ~from_pos(min(16, max(0, a + len - 16 * b)));
            }
            st.cw[b] = c;
            st.vh[b ^ 1] = (uint16_t)v;
        }
        __syncwarp();
        KeyT* ko = keys + row_begin[r];
        uint32_t* co = counts + row_begin[r];
        const int nd = warp_count_read<KeyT, E>(st, a, nwin, k, ko, co, s_stage_k[warp], s_stage_c[warp]);
        if (lane == 0) row_count[r] = nd;
    }
}

// ------------------------------------------------------------------------------------------
// long reads (> 512 windows)
constexpr int kSortTile = 2048;   // keys per warp tile of the radix sort
constexpr int kSortWarps = 8;

__global__ void collect_long_kernel(const int32_t* __restrict__ length, int64_t nS, int k, int64_t* __restrict__ long_rows,
                                    unsigned long long* __restrict__ n_long, int64_t cap)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nS; r += (int64_t)gridDim.x * blockDim.x) {
        if (length[r] - k + 1 > kShortMaxWindows) {
            const unsigned long long slot = atomicAdd(n_long, 1ull);
            if ((int64_t)slot < cap) long_rows[slot] = r;
        }
    }
}

// per long row j: windows and sort tiles (inputs of the two exclusive scans -> loff, ltile)
__global__ void long_sizes_kernel(const int64_t* __restrict__ long_rows, int64_t n_long, const int32_t* __restrict__ length,
                                  int k, int64_t* __restrict__ loff, int64_t* __restrict__ ltile)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= n_long) {
        const int64_t nwin = j < n_long ? length[long_rows[j]] - k + 1 : 0;
        loff[j] = nwin;
        ltile[j] = (nwin + kSortTile - 1) / kSortTile;
    }
}

template <typename KeyT>
__device__ __forceinline__ KeyT invalid_marker(int k)
{
    // sorts behind every real key; when the key type has a spare bit the sort needs 2k+1 bits only
    return 2 * k < (int)sizeof(KeyT) * 8 ? (KeyT)1 << (2 * k) : KeyMax<KeyT>::value;
}

// one CTA per long row: scalar rolling index, 16 windows per thread per pass
template <typename KeyT, int FMT>
__global__ void __launch_bounds__(256) long_keygen_kernel(const uint8_t* __restrict__ bases,
                                                          const int64_t* __restrict__ start,
                                                          const int32_t* __restrict__ length, int k,
                                                          const int64_t* __restrict__ long_rows,
                                                          const int64_t* __restrict__ loff,
                                                          KeyT* __restrict__ keys, int32_t* __restrict__ row_valid)
{
    const int64_t r = long_rows[blockIdx.x];
    const int64_t s = start[r];
    const int len = length[r];
    const int nwin = len - k + 1;
    KeyT* out = keys + loff[blockIdx.x];
    const KeyT mask = k * 2 >= (int)sizeof(KeyT) * 8 ? KeyMax<KeyT>::value : (((KeyT)1 << (2 * k)) - 1);
    const KeyT bad = invalid_marker<KeyT>(k);
    int valid = 0;
    for (int w0 = threadIdx.x * 16; w0 < nwin; w0 += blockDim.x * 16) {
        KeyT key = 0;
        int run = 0;
        const int wend = min(nwin, w0 + 16);
        for (int t = w0; t < wend + k - 1; t++) {       // base t closes the window starting at t-k+1
            const uint32_t c = bases[s + t];
            uint32_t code; bool ok;
            if (FMT == FMT_ASCII) {
                const uint32_t u = c & 0xDFu;
                code = ((c >> 1) ^ (c >> 2)) & 3u;
                ok = u == 'A' || u == 'C' || u == 'G' || u == 'T';
            } else {
                code = c & 3u; ok = !(c & 0x80u);
            }
            key = ((key << 2) | code) & mask;
            run = ok ? run + 1 : 0;
            const int wstart = t - k + 1;
            if (wstart >= w0) {
                const bool good = run >= k;
                out[wstart] = good ? key : bad;
                valid += good;
            }
        }
    }
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    atomicAdd(&s_valid, valid);
    __syncthreads();
    if (threadIdx.x == 0) row_valid[r] = s_valid;
}

// tile t of the flat tile list -> long row j, tile inside the row, keys of the tile
struct SortTile {
    int64_t j, tin, ntiles, key0;   // key0 = offset of the tile's first key in the scratch arrays
    int n;                          // keys in the tile
};
__device__ __forceinline__ SortTile locate_tile(int64_t t, const int64_t* __restrict__ ltile,
                                                const int64_t* __restrict__ loff, int64_t n_long)
{
    int64_t lo = 0, hi = n_long;   // last j with ltile[j] <= t
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ltile[mid] <= t) lo = mid; else hi = mid;
    }
    SortTile st;
    st.j = lo;
    st.tin = t - ltile[lo];
    st.ntiles = ltile[lo + 1] - ltile[lo];
    const int64_t nwin = loff[lo + 1] - loff[lo];
    st.key0 = loff[lo] + st.tin * kSortTile;
    st.n = (int)min((int64_t)kSortTile, nwin - st.tin * kSortTile);
    return st;
}

// digit histogram of every tile; counter layout = row-major, then digit, then tile-in-row, so that
// ONE flat exclusive scan gives, relative to the row's first counter, the destination of every
// (digit, tile) group inside its row (uint32 arithmetic modulo 2^32: rows are < 2^31 keys)
template <typename KeyT>
__global__ void __launch_bounds__(kSortWarps * 32) sort_hist_kernel(const KeyT* __restrict__ src, int shift,
                                                                    const int64_t* __restrict__ ltile,
                                                                    const int64_t* __restrict__ loff, int64_t n_long,
                                                                    int64_t ntiles_total, uint32_t* __restrict__ counters)
{
    __shared__ uint32_t s_h[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* h = s_h[warp];
    for (int64_t t = (int64_t)blockIdx.x * kSortWarps + warp; t < ntiles_total; t += (int64_t)gridDim.x * kSortWarps) {
        const SortTile st = locate_tile(t, ltile, loff, n_long);
        for (int i = lane; i < 256; i += 32) h[i] = 0u;
        __syncwarp();
        for (int i = lane; i < st.n; i += 32) atomicAdd(&h[(uint32_t)(src[st.key0 + i] >> shift) & 255u], 1u);
        __syncwarp();
        uint32_t* c = counters + ltile[st.j] * 256;
        for (int d = lane; d < 256; d += 32) c[(int64_t)d * st.ntiles + st.tin] = h[d];
        __syncwarp();
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortWarps * 32) sort_scatter_kernel(const KeyT* __restrict__ src, KeyT* __restrict__ dst,
                                                                       int shift, const int64_t* __restrict__ ltile,
                                                                       const int64_t* __restrict__ loff, int64_t n_long,
                                                                       int64_t ntiles_total,
                                                                       const uint32_t* __restrict__ scanned)
{
    __shared__ uint32_t s_b[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* base = s_b[warp];
    for (int64_t t = (int64_t)blockIdx.x * kSortWarps + warp; t < ntiles_total; t += (int64_t)gridDim.x * kSortWarps) {
        const SortTile st = locate_tile(t, ltile, loff, n_long);
        const uint32_t* c = scanned + ltile[st.j] * 256;
        const uint32_t row0 = c[0];
        for (int d = lane; d < 256; d += 32) base[d] = c[(int64_t)d * st.ntiles + st.tin] - row0;
        __syncwarp();
        KeyT* out = dst + loff[st.j];
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

// sorted keys of one long row (first nvalid real) -> (key,count) pairs, one CTA per row
template <typename KeyT>
__global__ void __launch_bounds__(256) long_rle_kernel(const int64_t* __restrict__ long_rows,
                                                       const int64_t* __restrict__ loff,
                                                       const int64_t* __restrict__ row_begin,
                                                       const KeyT* __restrict__ sorted, KeyT* __restrict__ keys,
                                                       uint32_t* __restrict__ counts, int32_t* __restrict__ row_count)
{
    constexpr int T = 256;
    const int64_t r = long_rows[blockIdx.x];
    const int64_t base = row_begin[r];
    const int nvalid = row_count[r];   // holds the number of valid windows on entry
    const KeyT* in = sorted + loff[blockIdx.x];
    KeyT* ko = keys + base;
    uint32_t* co = counts + base;      // pass 1: position of each head; pass 2: run lengths
    __shared__ int s_warp[T / 32];
    __shared__ int s_total;
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
        __syncthreads();
        if (head) { ko[pos] = v; co[pos] = (uint32_t)g; }
        nheads += total;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_total = nheads;
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
    if (threadIdx.x == 0) row_count[r] = s_total;
}

// ------------------------------------------------------------------------------------------
template <typename KeyT, int FMT>
static cudaError_t sparse_impl(const void* bases, const int64_t* start, const int32_t* length, int64_t nS, int k,
                               int64_t* row_begin, int32_t* row_count, KeyT* keys, uint32_t* counts,
                               int64_t capacity, int64_t* total_windows, cudaStream_t st)
{
    cudaError_t e;
    int dev = 0, num_sms = 148;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);

    SparseTrace tr(st);
    keep_pool_memory(dev);
    // 1. row_begin = exclusive scan of window counts
    nwin_kernel<<<num_sms * 4, 256, 0, st>>>(length, nS, k, row_begin);
    count_launch();
    size_t tmp_bytes = 0;
    if ((e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, row_begin, row_begin, nS + 1, st)) != cudaSuccess) return e;
    void* tmp = nullptr;
    if ((e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st)) != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, row_begin, row_begin, nS + 1, st);
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return e;
    int64_t total = 0;
    if ((e = cudaMemcpyAsync(&total, row_begin + nS, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
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
        const uint8_t* b8 = static_cast<const uint8_t*>(bases);
        sparse_short_kernel<KeyT, FMT, 4><<<grid, SparseCta<4>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts);
        sparse_short_kernel<KeyT, FMT, 8><<<grid, SparseCta<8>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts);
        sparse_short_kernel<KeyT, FMT, 16><<<grid, SparseCta<16>::WARPS * 32, 0, st>>>(b8, start, length, nS, k, row_begin, row_count, keys, counts);
        count_launch(); count_launch(); count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }

    tr.mark("short reads");
    // 3. long reads
    int64_t* long_rows = nullptr;
    unsigned long long* d_nlong = nullptr;
    const int64_t cap_long = total / kShortMaxWindows + 1;   // a long row has > 512 windows
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&long_rows), (size_t)cap_long * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&d_nlong), 8, st)) != cudaSuccess) return e;
    cudaMemsetAsync(d_nlong, 0, 8, st);
    collect_long_kernel<<<num_sms * 4, 256, 0, st>>>(length, nS, k, long_rows, d_nlong, cap_long);
    count_launch();
    unsigned long long n_long = 0;
    cudaMemcpyAsync(&n_long, d_nlong, 8, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (n_long > 0) {
        const int64_t nl = (int64_t)n_long;
        int64_t *loff = nullptr, *ltile = nullptr;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&loff), (size_t)(nl + 1) * 8, st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&ltile), (size_t)(nl + 1) * 8, st)) != cudaSuccess) return e;
        long_sizes_kernel<<<(unsigned)((nl + 256) / 256), 256, 0, st>>>(long_rows, nl, length, k, loff, ltile);
        count_launch();
        size_t sb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, sb, loff, loff, nl + 1, st);
        void* stmp = nullptr;
        if ((e = cudaMallocAsync(&stmp, sb ? sb : 16, st)) != cudaSuccess) return e;
        if ((e = cub::DeviceScan::ExclusiveSum(stmp, sb, loff, loff, nl + 1, st)) != cudaSuccess) return e;
        if ((e = cub::DeviceScan::ExclusiveSum(stmp, sb, ltile, ltile, nl + 1, st)) != cudaSuccess) return e;
        cudaFreeAsync(stmp, st);
        int64_t total_long = 0, ntiles = 0;
        cudaMemcpyAsync(&total_long, loff + nl, 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&ntiles, ltile + nl, 8, cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;

        KeyT *bufA = nullptr, *bufB = nullptr;
        uint32_t* counters = nullptr;
        const size_t ncount = (size_t)ntiles * 256;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&bufA), (size_t)total_long * sizeof(KeyT), st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&bufB), (size_t)total_long * sizeof(KeyT), st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&counters), ncount * 4, st)) != cudaSuccess) return e;
        tr.mark("long rows: collect + scratch");
        long_keygen_kernel<KeyT, FMT><<<(unsigned)nl, 256, 0, st>>>(static_cast<const uint8_t*>(bases), start, length, k,
                                                                   long_rows, loff, bufA, row_count);
        count_launch();
        tr.mark("long rows: key generation");
        // segmented LSD radix sort, 8 bits per pass, over the bits that can differ
        const int key_bits = (int)sizeof(KeyT) * 8;
        const int nbits = 2 * k + 1 < key_bits ? 2 * k + 1 : key_bits;
        const int passes = (nbits + 7) / 8;
        size_t cb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, cb, counters, counters, (int64_t)ncount, st);
        void* ctmp = nullptr;
        if ((e = cudaMallocAsync(&ctmp, cb ? cb : 16, st)) != cudaSuccess) return e;
        const int64_t ctas = (ntiles + kSortWarps - 1) / kSortWarps;
        const unsigned sgrid = (unsigned)(ctas < (int64_t)num_sms * 8 ? ctas : (int64_t)num_sms * 8);
        KeyT *src = bufA, *dst = bufB;
        for (int p = 0; p < passes; p++) {
            sort_hist_kernel<KeyT><<<sgrid, kSortWarps * 32, 0, st>>>(src, 8 * p, ltile, loff, nl, ntiles, counters);
            tr.mark("  sort pass: histogram");
            if ((e = cub::DeviceScan::ExclusiveSum(ctmp, cb, counters, counters, (int64_t)ncount, st)) != cudaSuccess) return e;
            tr.mark("  sort pass: scan");
            sort_scatter_kernel<KeyT><<<sgrid, kSortWarps * 32, 0, st>>>(src, dst, 8 * p, ltile, loff, nl, ntiles, counters);
            tr.mark("  sort pass: scatter");
            count_launch(); count_launch();
            KeyT* tmpp = src; src = dst; dst = tmpp;
        }
        long_rle_kernel<KeyT><<<(unsigned)nl, 256, 0, st>>>(long_rows, loff, row_begin, src, keys, counts, row_count);
        count_launch();
        tr.mark("long rows: run-length encode");
        cudaFreeAsync(ctmp, st);
        cudaFreeAsync(counters, st);
        cudaFreeAsync(bufA, st);
        cudaFreeAsync(bufB, st);
        cudaFreeAsync(loff, st);
        cudaFreeAsync(ltile, st);
    }
    cudaFreeAsync(long_rows, st);
    cudaFreeAsync(d_nlong, st);
    return cudaGetLastError();
}

cudaError_t launch_sparse(const void* bases, int fmt, const int64_t* start, const int32_t* length, int64_t nS, int k,
                          int64_t* row_begin, int32_t* row_count, void* keys, int key_bytes, uint32_t* counts,
                          int64_t capacity, int64_t* total_windows, cudaStream_t st)
{
    if (key_bytes == 4) {
        auto* kk = static_cast<uint32_t*>(keys);
        return fmt == FMT_ASCII
                   ? sparse_impl<uint32_t, FMT_ASCII>(bases, start, length, nS, k, row_begin, row_count, kk, counts, capacity, total_windows, st)
                   : sparse_impl<uint32_t, FMT_CODES>(bases, start, length, nS, k, row_begin, row_count, kk, counts, capacity, total_windows, st);
    }
    auto* kk = static_cast<uint64_t*>(keys);
    return fmt == FMT_ASCII
               ? sparse_impl<uint64_t, FMT_ASCII>(bases, start, length, nS, k, row_begin, row_count, kk, counts, capacity, total_windows, st)
               : sparse_impl<uint64_t, FMT_CODES>(bases, start, length, nS, k, row_begin, row_count, kk, counts, capacity, total_windows, st);
}

}  // namespace cfrk
