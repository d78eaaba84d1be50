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
//   long reads: key generation kernel -> cub::DeviceSegmentedRadixSort (LIBRARY sort, to be
//            replaced) -> hand-written per-row run-length encode.
#include "kernels.h"
#include "kmer_device.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

#include <atomic>

namespace cfrk {

extern void count_launch();

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
// long reads
__global__ void collect_long_kernel(const int32_t* __restrict__ length, int64_t nS, int k,
                                    const int64_t* __restrict__ row_begin, int64_t* __restrict__ long_rows,
                                    unsigned long long* __restrict__ n_long, int64_t cap)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nS; r += (int64_t)gridDim.x * blockDim.x) {
        if (length[r] - k + 1 > kShortMaxWindows) {
            const unsigned long long slot = atomicAdd(n_long, 1ull);
            if ((int64_t)slot < cap) long_rows[slot] = r;
        }
    }
}

// one CTA per long row: scalar rolling index, 16 windows per thread per pass
template <typename KeyT, int FMT>
__global__ void __launch_bounds__(256) long_keygen_kernel(const uint8_t* __restrict__ bases,
                                                          const int64_t* __restrict__ start,
                                                          const int32_t* __restrict__ length, int k,
                                                          const int64_t* __restrict__ long_rows,
                                                          const int64_t* __restrict__ row_begin,
                                                          KeyT* __restrict__ keys, int32_t* __restrict__ row_valid)
{
    const int64_t r = long_rows[blockIdx.x];
    const int64_t s = start[r];
    const int len = length[r];
    const int nwin = len - k + 1;
    KeyT* out = keys + row_begin[r];
    const KeyT mask = k * 2 >= (int)sizeof(KeyT) * 8 ? KeyMax<KeyT>::value : (((KeyT)1 << (2 * k)) - 1);
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
                out[wstart] = good ? key : KeyMax<KeyT>::value;
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

__global__ void long_segments_kernel(const int64_t* __restrict__ long_rows, int64_t n_long,
                                     const int64_t* __restrict__ row_begin, const int32_t* __restrict__ length, int k,
                                     int64_t* __restrict__ seg_begin, int64_t* __restrict__ seg_end)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_long) {
        const int64_t r = long_rows[j];
        seg_begin[j] = row_begin[r];
        seg_end[j] = row_begin[r] + (length[r] - k + 1);
    }
}

// sorted keys of one long row (first nvalid real) -> (key,count) pairs, one CTA per row
template <typename KeyT>
__global__ void __launch_bounds__(256) long_rle_kernel(const int64_t* __restrict__ long_rows,
                                                       const int64_t* __restrict__ row_begin,
                                                       const KeyT* __restrict__ sorted, KeyT* __restrict__ keys,
                                                       uint32_t* __restrict__ counts, int32_t* __restrict__ row_count)
{
    constexpr int T = 256;
    const int64_t r = long_rows[blockIdx.x];
    const int64_t base = row_begin[r];
    const int nvalid = row_count[r];   // holds the number of valid windows on entry
    const KeyT* in = sorted + base;
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
        __syncthreads();   // all reads of `in` for this chunk are done (ko may alias `sorted`'s row)
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

    // 3. long reads
    int64_t* long_rows = nullptr;
    unsigned long long* d_nlong = nullptr;
    const int64_t cap_long = total / kShortMaxWindows + 1;   // a long row has > 512 windows
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&long_rows), (size_t)cap_long * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&d_nlong), 8, st)) != cudaSuccess) return e;
    cudaMemsetAsync(d_nlong, 0, 8, st);
    collect_long_kernel<<<num_sms * 4, 256, 0, st>>>(length, nS, k, row_begin, long_rows, d_nlong, cap_long);
    count_launch();
    unsigned long long n_long = 0;
    cudaMemcpyAsync(&n_long, d_nlong, 8, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (n_long > 0) {
        KeyT* unsorted = nullptr;
        int64_t *seg_b = nullptr, *seg_e = nullptr;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&unsorted), (size_t)total * sizeof(KeyT), st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&seg_b), (size_t)n_long * 8, st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(reinterpret_cast<void**>(&seg_e), (size_t)n_long * 8, st)) != cudaSuccess) return e;
        long_keygen_kernel<KeyT, FMT><<<(unsigned)n_long, 256, 0, st>>>(static_cast<const uint8_t*>(bases), start, length, k,
                                                                       long_rows, row_begin, unsorted, row_count);
        count_launch();
        long_segments_kernel<<<(unsigned)((n_long + 255) / 256), 256, 0, st>>>(long_rows, (int64_t)n_long, row_begin, length,
                                                                              k, seg_b, seg_e);
        count_launch();
        // LIBRARY: segmented radix sort of the long rows (unsorted -> keys)
        size_t sort_bytes = 0;
        e = cub::DeviceSegmentedRadixSort::SortKeys(nullptr, sort_bytes, unsorted, keys, total, (int)n_long, seg_b, seg_e,
                                                    0, sizeof(KeyT) * 8, st);
        if (e != cudaSuccess) return e;
        void* sort_tmp = nullptr;
        if ((e = cudaMallocAsync(&sort_tmp, sort_bytes ? sort_bytes : 16, st)) != cudaSuccess) return e;
        e = cub::DeviceSegmentedRadixSort::SortKeys(sort_tmp, sort_bytes, unsorted, keys, total, (int)n_long, seg_b, seg_e,
                                                    0, sizeof(KeyT) * 8, st);
        if (e != cudaSuccess) return e;
        // the sorted rows sit in `keys`; copy them back so that the RLE can write `keys` in place
        // from a stable source
        if ((e = cudaMemcpyAsync(unsorted, keys, (size_t)total * sizeof(KeyT), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        long_rle_kernel<KeyT><<<(unsigned)n_long, 256, 0, st>>>(long_rows, row_begin, unsorted, keys, counts, row_count);
        count_launch();
        cudaFreeAsync(sort_tmp, st);
        cudaFreeAsync(unsorted, st);
        cudaFreeAsync(seg_b, st);
        cudaFreeAsync(seg_e, st);
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
