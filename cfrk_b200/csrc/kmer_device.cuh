// kmer_device.cuh -- device-side building blocks of the B200 k-mer counting path.
//
// Replaces the reference's ComputeIndex (src/kmer_kernel.cu:21-49: k byte loads + k powf
// per position, Index[] written to HBM) and ComputeFreqNew (src/kmer_kernel.cu:73-90: one
// 1024-thread block per read, one L2 atomic per k-mer) with one fused pass in which
//   * a lane owns one 16-byte aligned block of the bases buffer (one 128-bit load),
//   * the block is encoded in registers to 16 x 2-bit codes + a 16-bit validity mask,
//   * the k-1 bases of context come from the previous lane by warp shuffle,
//   * window validity for all 16 positions is one bit-parallel AND chain,
//   * the index of the window ending at base j is one funnel shift of (carry:codes).
// Nothing but the bases and the result rows ever touches HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cfrk {

constexpr int FMT_CODES = 0;
constexpr int FMT_ASCII = 1;
constexpr int FMT_PACKED = 2;   // what encode_2bit_kernel writes: 16 bases per uint32 + uint16 validity
constexpr int MODE_COMPAT = 0;
constexpr int MODE_EXACT = 1;
constexpr int kRefBlockThreads = 1024;  // reference blockDim (src/kmer_main.cu:82)
constexpr int kMaxGroupReads = 256;     // reads per work group (tile) at most

// ------------------------------------------------------------------------------------------
// 128-bit streaming load of one block of bases (read once: keep it out of L1)
__device__ __forceinline__ uint4 ld_block(const uint8_t* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// The bases of a batch: one byte per base (CODES / ASCII), or the packed 2-bit words + validity
// masks of FMT_PACKED (p = uint32 codes[], valid = uint16 masks[], one entry per 16-base block).
struct BasesRef {
    const uint8_t* p;
    const uint16_t* valid;
};

// one 16-base block: the 16 raw bytes, or {codes, valid, 0, 0} when the batch is already encoded
template <int FMT>
__device__ __forceinline__ uint4 load_block(const BasesRef& b, int64_t blk)
{
    if (FMT == FMT_PACKED) {
        uint32_t c; uint16_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(c) : "l"(reinterpret_cast<const uint32_t*>(b.p) + blk));
        asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(b.valid + blk));
        return make_uint4(c, (uint32_t)v, 0u, 0u);
    }
    return ld_block(b.p + blk * 16);
}

// ------------------------------------------------------------------------------------------
// 4 bases (one little-endian 32-bit word, byte 0 = first base) -> 8 bits of codes (first
// base in bits 7:6) and a 4-bit validity mask (first base in bit 3).
//
// ASCII: code = ((c>>1)^(c>>2))&3 maps A/C/G/T (either case) to 0/1/2/3 (src/fastaIO.h:123-139);
// a byte is valid iff (c & 0xDF) equals the letter its code stands for, checked for the 4
// bytes at once with a PRMT table lookup.  CODES: value & 3, valid iff bit 7 is clear
// (the only negative value the reference layout holds is -1, src/fastaIO.h:137).
template <int FMT>
__device__ __forceinline__ void encode4(uint32_t w, uint32_t& codes8, uint32_t& valid4)
{
    uint32_t x, z;
    if (FMT == FMT_ASCII) {
        x = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
        // selector nibbles (c0, c2, c1, c3): expected letters in byte order (0, 2, 1, 3)
        uint32_t sel = (x | (x >> 12)) & 0x3333u;
        uint32_t expect = __byte_perm(0x54474341u /* "ACGT" */, 0u, sel);
        uint32_t d = (__byte_perm(w, 0u, 0x3120u) & 0xDFDFDFDFu) ^ expect;
        // bit 7 of each byte <- (byte == 0)
        z = ~(((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
        // bits 7 (base0), 23 (base1), 15 (base2), 31 (base3) -> bits 35, 34, 33, 32
        valid4 = __umulhi(z, (1u << 28) | (1u << 18) | (1u << 11) | (1u << 1)) & 0xFu;
    } else {
        x = w & 0x03030303u;
        z = ~w & 0x80808080u;
        // bits 7, 15, 23, 31 (bases 0..3) -> bits 35, 34, 33, 32
        valid4 = __umulhi(z, (1u << 28) | (1u << 19) | (1u << 10) | (1u << 1)) & 0xFu;
    }
    // 2-bit fields at bits 0, 8, 16, 24 -> bits 30, 28, 26, 24 (no overlapping partial products)
    codes8 = (x * 0x40100401u) >> 24;
}

// 16 bases -> codes (base j in bits 31-2j:30-2j) and valid (base j in bit 15-j)
template <int FMT>
__device__ __forceinline__ void encode16(const uint4 v, uint32_t& codes, uint32_t& valid)
{
    uint32_t c0, c1, c2, c3, m0, m1, m2, m3;
    encode4<FMT>(v.x, c0, m0);
    encode4<FMT>(v.y, c1, m1);
    encode4<FMT>(v.z, c2, m2);
    encode4<FMT>(v.w, c3, m3);
    codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    valid = (m0 << 12) | (m1 << 8) | (m2 << 4) | m3;
}

// what load_block() returned -> codes + validity, whatever the layout
template <int FMT>
__device__ __forceinline__ void decode_block(const uint4 raw, uint32_t& codes, uint32_t& valid)
{
    if (FMT == FMT_PACKED) { codes = raw.x; valid = raw.y; }
    else encode16<FMT == FMT_PACKED ? FMT_CODES : FMT>(raw, codes, valid);
}

// mask with the bits of positions [p, 16) set (position j <-> bit 15-j); p in [0, 16]
__device__ __forceinline__ uint32_t from_pos(int p) { return (0x10000u >> p) - 1u; }

// ------------------------------------------------------------------------------------------
// Per-group (tile) read table in shared memory.
struct ReadTable {
    int64_t*  start;  // byte offset of the read in the bases buffer
    int32_t*  tend;   // bytes [0, tend) of the read carry a counted window end
    int32_t*  extra;  // compat: windows that straddle the terminator (always invalid)
    uint32_t* cum;    // exclusive prefix sum of 16-byte blocks per read; cum[n] = total
};

// Which window-end positions of a read are visited.
//   compat: starts t < min(len-1, 1024) (src/kmer_kernel.cu:85, src/kmer_main.cu:82), so ends
//           < min(len-1,1024) + k-1; those at or beyond len hold the terminator -> `extra`.
//   exact : every window inside the read.
//   compat, len == 0: `threadIdx.x < length[i]-1` is an UNSIGNED compare in the reference, so an
//           empty read lets all 1024 threads through: the block walks over the bytes that follow
//           (terminators, later reads) and counts them into the empty read's row.  Modelled as a
//           read that extends to the end of the buffer (`avail` bytes) with 1024 visited starts.
template <int K>
__device__ __forceinline__ void read_extent(int mode, int len, int64_t avail, int& tend, int& extra)
{
    int vis;
    if (mode == MODE_COMPAT) {
        if (len == 0) {
            len = (int)min(avail, (int64_t)(kRefBlockThreads + K));
            vis = min(len, kRefBlockThreads);
        } else {
            vis = min(len - 1, kRefBlockThreads);
        }
    } else {
        vis = len - K + 1;
    }
    if (vis <= 0) { tend = 0; extra = 0; return; }
    int last = vis + K - 1;
    tend = min(len, last);
    // visited starts t0 >= len-K+1 end at or beyond the terminator
    extra = (mode == MODE_COMPAT) ? max(0, vis - max(0, len - K + 1)) : 0;
}

// Bytes an EMPTY read may walk over in compat mode: up to the end of ITS reference chunk's buffer.  The
// reference runs one kmer_main per chunk on a buffer that holds that chunk's reads only
// (src/main.cu:110-206,222), so the walk of an empty read stops where the next chunk begins; one
// launch here covers many chunks (chunk_size / index_base), and the next chunk's reads follow in the
// same buffer.  Only evaluated for len == 0 (rare).
struct ChunkScope {
    const int64_t* start;
    int64_t nS, nN, chunk_size, index_base;
};
__device__ __forceinline__ int64_t compat_avail(const ChunkScope& c, int64_t r, int64_t s, int len, int mode)
{
    if (mode != MODE_COMPAT || len != 0 || c.chunk_size <= 0) return c.nN - s;
    const int64_t e = r + (c.chunk_size - (c.index_base + r) % c.chunk_size);   // first read of the next chunk
    return (e < c.nS ? c.start[e] : c.nN) - s;
}

// Fill the table for reads [r0, r0+n) (n <= kMaxGroupReads+1).  Call with all threads, then
// __syncthreads(), then scan_read_table() from warp 0, then __syncthreads().
template <int K>
__device__ __forceinline__ void fill_read_table(const ReadTable& tb, const int64_t* __restrict__ start,
                                                const int32_t* __restrict__ length, int64_t r0, int n,
                                                int mode, int64_t nN, int64_t nS = 0, int64_t chunk_size = 0,
                                                int64_t index_base = 0)
{
    const ChunkScope cs{start, nS, nN, chunk_size, index_base};
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        int64_t s = start[r0 + q];
        int len = length[r0 + q];
        int tend, extra;
        read_extent<K>(mode, len, compat_avail(cs, r0 + q, s, len, mode), tend, extra);
        tb.start[q] = s;
        tb.tend[q] = tend;
        tb.extra[q] = extra;
        tb.cum[q] = tend > 0 ? (uint32_t)(((s + tend - 1) >> 4) - (s >> 4) + 1) : 0u;
    }
}

// exclusive scan of tb.cum[0..n) in place, total to tb.cum[n]; one warp
__device__ __forceinline__ void scan_read_table(const ReadTable& tb, int n)
{
    const int lane = threadIdx.x & 31;
    uint32_t run = 0;
    for (int base = 0; base < n; base += 32) {
        int q = base + lane;
        uint32_t v = q < n ? tb.cum[q] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (q < n) tb.cum[q] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) tb.cum[n] = run;
}

template <int N> struct Log2Ceil { static constexpr int value = 1 + Log2Ceil<(N + 1) / 2>::value; };
template <> struct Log2Ceil<1> { static constexpr int value = 0; };

// ------------------------------------------------------------------------------------------
// One item = one 16-byte aligned block of one read's counted byte range.
struct Item {
    uint4 raw;          // the 16 bytes
    int q;              // table-local read
    int t0;             // read-relative position of byte 0 of the block (negative in the first block)
    int tend;           // counted window ends are t < tend
    int extra;          // compat: invalid windows that reach the terminator (charged to block 0)
    bool live, first_block;
};

// Everything that follows the load of an item: encode, context from the previous lane, window
// validity, emission.  Must be called by all 32 lanes (shuffles).  Lane 0 only feeds lane 1.
//   sink.kmer(q, idx)        one valid window of read q with index idx
//   sink.invalid(q, in_read, extra)   compat only: visited windows of read q that held a non-ACGT
//                            byte (`in_read`, data dependent) or reached the terminator (`extra`,
//                            a function of the read length alone, reported with the first block)
// Reads q >= ncount (the halo read of a compat tile) only report invalid windows.
template <int K, int FMT, class Sink>
__device__ __forceinline__ void emit_item(const Item& it, int ncount, int mode, Sink& sink)
{
    static_assert(K >= 1 && K <= 16, "one 32-bit funnel shift per window needs k <= 16");
    constexpr uint32_t IDX_MASK = (K == 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
    const int lane = threadIdx.x & 31;
    uint32_t codes = 0, valid = 0, count_mask = 0;
    if (it.live) {
        if (FMT == FMT_PACKED) { codes = it.raw.x; valid = it.raw.y; }
        else encode16<FMT == FMT_PACKED ? FMT_CODES : FMT>(it.raw, codes, valid);
        const uint32_t upto = ~from_pos(min(16, it.tend - it.t0));
        valid &= from_pos(max(0, -it.t0)) & upto;                       // bases of this read only
        count_mask = from_pos(min(16, max(0, K - 1 - it.t0))) & upto;   // window ends that count
    }
    const uint32_t pcodes = __shfl_up_sync(0xffffffffu, codes, 1);
    uint32_t pvalid = __shfl_up_sync(0xffffffffu, valid, 1);
    if (it.first_block) pvalid = 0;  // no context across a read start
    const bool counting = it.live && lane != 0;
    if (!counting) count_mask = 0;

    // bit b of `ok` <- bases b .. b+K-1 (the window ending at position 15-b) are all valid
    const uint32_t v32 = (pvalid << 16) | valid;
    uint32_t ok = v32;
#pragma unroll
    for (int i = 1; i < K; i++) ok &= v32 >> i;
    const uint32_t good = it.q < ncount ? (ok & count_mask) : 0u;
    if (mode == MODE_COMPAT) {
        const int nbad = __popc(~ok & count_mask);
        const int nextra = (counting && it.first_block) ? it.extra : 0;
        if (nbad | nextra) sink.invalid(it.q, nbad, nextra);
    }
    if constexpr (Sink::kSharedRows) {
        // Rows live in shared memory; when each is aligned to its own size (4^K * 4 bytes,
        // Sink::kRowsAligned) the address of a bin is row | (index << 2), so a window costs one
        // funnel shift, one LOP3, one predicate and one predicated red.shared -- straight-line
        // code, no branches.
        static_assert(2 * K + 2 <= 32, "shared rows are for k <= 8");
        if (good) {   // (warp-uniformly false only for halo-only chunks)
            const uint32_t row = sink.row_saddr(it.q);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int bit = 15 - j;
                const uint32_t t = bit >= 1 ? __funnelshift_r(codes, pcodes, 2 * bit - 2) : (codes << 2);
                const uint32_t addr = Sink::kRowsAligned ? ((t & (IDX_MASK << 2)) | row) : ((t & (IDX_MASK << 2)) + row);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}"
                             :: "r"(addr), "r"(good & (1u << bit)) : "memory");
            }
        }
    } else {
        if (good) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int bit = 15 - j;
                if (good & (1u << bit)) sink.kmer(it.q, __funnelshift_r(codes, pcodes, 2 * bit) & IDX_MASK);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// CTA-cooperative item loop over a shared-memory ReadTable.  Lanes 1..31 of a warp take 31
// consecutive items; lane 0 re-encodes the item before them so that every counting lane gets its
// k-1 bases of context with one shuffle.
//   sink.before_first_emit() (only if Sink::kCtaUniform) is reached by EVERY thread of the CTA
//                            exactly once, after the first chunk's loads were issued and before
//                            anything is emitted (a place for a CTA-wide wait + barrier)
// MAXREADS bounds n (binary search depth).
template <int K, int FMT, int MAXREADS, class Sink>
__device__ __forceinline__ void for_each_window(const BasesRef& bases, const ReadTable& tb,
                                                int n, int ncount, int mode, Sink& sink)
{
    constexpr int STEPS = Log2Ceil<MAXREADS>::value;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const uint32_t total = tb.cum[n];
    const uint32_t nchunks = (total + 30u) / 31u;

    // kCtaUniform: every warp runs the same number of iterations (idle ones with no live lane)
    const uint32_t niter = (nchunks + nwarps - 1) / nwarps;
    for (uint32_t iter = 0, chunk = warp; Sink::kCtaUniform ? (iter < niter) : (chunk < nchunks);
         ++iter, chunk += nwarps) {
        const int64_t item = (int64_t)chunk * 31 + lane - 1;
        Item it;
        it.live = chunk < nchunks && item >= 0 && item < (int64_t)total;
        it.q = 0; it.t0 = 0; it.tend = 0; it.extra = 0; it.first_block = true;
        it.raw = make_uint4(0u, 0u, 0u, 0u);
        if (it.live) {
            // largest q with cum[q] <= item  (cum[0] == 0; interval [lo, hi) halves each step)
            int lo = 0, hi = n;
#pragma unroll
            for (int s = 0; s < STEPS; s++) {
                const int mid = (lo + hi) >> 1;
                const bool right = tb.cum[mid] <= (uint32_t)item;
                lo = right ? mid : lo;
                hi = right ? hi : mid;
            }
            it.q = lo;
            const uint32_t c = (uint32_t)item - tb.cum[lo];
            const int64_t s = tb.start[lo];
            const int64_t blk = (s >> 4) + c;
            it.tend = tb.tend[lo];
            it.extra = tb.extra[lo];
            it.t0 = (int)(blk * 16 - s);
            it.first_block = (c == 0);
            it.raw = load_block<FMT>(bases, blk);
        }
        if constexpr (Sink::kCtaUniform) { if (iter == 0) sink.before_first_emit(); }
        emit_item<K, FMT>(it, ncount, mode, sink);
    }
}

// ------------------------------------------------------------------------------------------
// Warp-autonomous item loop: the read table of a small tile (n <= 32 reads, lane q holds read q)
// lives in registers, lookups are shuffles, the next chunk's blocks are loaded before the current
// chunk is processed.  No shared-memory table, no CTA barrier.
struct LaneRead {      // read q = lane q of the warp
    int64_t start;
    int tend, extra;
    uint32_t nblk, cum;  // 16-byte blocks of this read; exclusive prefix over the tile
};

// have: the lane describes a read (tend / extra are computed); items: its blocks are enumerated
template <int K>
__device__ __forceinline__ LaneRead make_lane_read(bool have, bool items, int64_t s, int len, int mode, int64_t avail)
{
    LaneRead lr;
    lr.start = s; lr.tend = 0; lr.extra = 0; lr.nblk = 0; lr.cum = 0;
    if (have) {
        read_extent<K>(mode, len, avail, lr.tend, lr.extra);
        lr.nblk = (items && lr.tend > 0) ? (uint32_t)(((s + lr.tend - 1) >> 4) - (s >> 4) + 1) : 0u;
    }
    uint32_t inc = lr.nblk;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    lr.cum = inc - lr.nblk;
    return lr;
}

template <int FMT, int NREADS_MAX>
__device__ __forceinline__ Item warp_fetch_item(const BasesRef& bases, const LaneRead& lr, int n,
                                                uint32_t total, uint32_t chunk, bool any_empty)
{
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)chunk * 31 + lane - 1;
    Item it;
    it.live = item >= 0 && item < (int64_t)total;
    // q = last read whose first item is <= item (reads without items share their successor's cum)
    int q = 0;
    if (NREADS_MAX > 4 && !any_empty) {
        // every read has items, so cum is strictly increasing: q(lane) = q(lane 0) + number of
        // reads that begin at items item0+1 .. item0+lane  -- one ballot + one warp OR-reduction
        const int64_t item0 = (int64_t)chunk * 31 - 1;
        const int cnt0 = __popc(__ballot_sync(0xffffffffu, lane < n && (int64_t)lr.cum <= item0));
        const int64_t d = (int64_t)lr.cum - item0;
        const uint32_t bit = (lane < n && d >= 1 && d <= 31) ? (1u << (int)d) : 0u;
        const uint32_t begins = __reduce_or_sync(0xffffffffu, bit);
        q = max(0, cnt0 - 1 + __popc(begins & ((2u << lane) - 1u)));
    } else {
#pragma unroll
        for (int j = 1; j < NREADS_MAX; j++) {
            const uint32_t cj = __shfl_sync(0xffffffffu, lr.cum, j);
            if (j < n && cj <= (uint32_t)item) q = j;
        }
    }
    const uint32_t cq = __shfl_sync(0xffffffffu, lr.cum, q);
    const int64_t s = __shfl_sync(0xffffffffu, lr.start, q);
    it.tend = __shfl_sync(0xffffffffu, lr.tend, q);
    it.extra = __shfl_sync(0xffffffffu, lr.extra, q);
    const uint32_t c = (uint32_t)item - cq;
    const int64_t blk = (s >> 4) + c;
    it.q = q;
    it.t0 = (int)(blk * 16 - s);
    it.first_block = (c == 0) || !it.live;
    it.raw = it.live ? load_block<FMT>(bases, blk) : make_uint4(0u, 0u, 0u, 0u);
    return it;
}

template <int K, int FMT, int NREADS_MAX, class Sink>
__device__ __forceinline__ void warp_for_each_window(const BasesRef& bases, const LaneRead& lr,
                                                     int n, int ncount, int mode, Sink& sink)
{
    const uint32_t total = __shfl_sync(0xffffffffu, lr.cum + lr.nblk, 31);
    const uint32_t nchunks = (total + 30u) / 31u;
    if (nchunks == 0) return;
    const bool any_empty = __ballot_sync(0xffffffffu, (threadIdx.x & 31) < n && lr.nblk == 0) != 0u;
    Item next = warp_fetch_item<FMT, NREADS_MAX>(bases, lr, n, total, 0, any_empty);
    for (uint32_t chunk = 0; chunk < nchunks; chunk++) {
        const Item cur = next;
        if (chunk + 1 < nchunks) next = warp_fetch_item<FMT, NREADS_MAX>(bases, lr, n, total, chunk + 1, any_empty);
        emit_item<K, FMT>(cur, ncount, mode, sink);
    }
}

// ------------------------------------------------------------------------------------------
// TMA (bulk async-copy engine) shared -> global store of one finished row tile.
__device__ __forceinline__ void fence_async_proxy_shared()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_tile(void* gdst, const void* ssrc, uint32_t bytes)
{
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_issue(void* gdst, const void* ssrc, uint32_t bytes)
{
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// order async-proxy (TMA) writes to global before later generic-proxy accesses
__device__ __forceinline__ void fence_async_proxy_global()
{
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace cfrk
