// dense_lane.cu -- dense per-read rows for small k (1..4): a LANE (or a group of 2 / 4 lanes) per read.
//
// Replaces, for k <= 4, SetMatrix x2 + ComputeIndex + ComputeFreqNew (reference src/kmer_kernel.cu:6-90
// as launched by src/kmer_main.cu:107-111).  Why another kernel family: with rows of 16 B .. 1 KiB the
// lane-per-16-byte-block mapping of the other dense kernels (kmer_device.cuh emit_item) makes ~10 lanes
// hammer the same 4..256 counters with red.shared and spends a third of its instructions on finding out
// which read a block belongs to (ncu, round 1: k=4 524 warp-instructions per 512 bases, 7.3 shared
// wavefronts per red.shared; k=2 at 0.21 of the HBM roofline).  Here
//   * a warp owns a tile of R = 32 / SPLIT consecutive reads; lane group q <-> read q, so the read table
//     is the lane's own registers and the k-1 bases of context are the lane's previous block (SPLIT = 1)
//     or one shuffle inside the group (SPLIT = 2, 4: the group's lanes take the read's blocks round-robin);
//   * the tile's bases -- one contiguous span of the buffer -- are staged into shared memory by ONE TMA
//     bulk copy (cp.async.bulk.shared.global + mbarrier) per tile, so the per-lane 16-byte reads at a
//     151-byte stride hit shared memory instead of 32 different L1 lines per load instruction;
//   * k = 1, 2 (4 / 16 bins): no shared-memory counters at all.  The 16 windows of a block are counted
//     BIT-PARALLEL: 2-bit codes and validity stay in "one bit pair per base" form, a bin's count is one
//     LOP3 + POPC + IADD on the whole block, the row lives in 4 / 16 registers of the lane;
//   * k = 3, 4 (256 B / 1 KiB rows): one private row per lane group in shared memory, predicated
//     red.shared as before, but never two reads in one row and never a search for the row;
//   * rows leave through one TMA bulk store per tile (cp.async.bulk.global.shared).
// Tiles whose span does not fit the staging buffer (long reads, scattered reads, empty reads in compat
// mode) are processed warp-wide, read after read, with coalesced global loads -- same code, group
// width 32 -- so every input is handled; there is no second kernel and no host-side dispatch on lengths.
//
// Compat spill (src/kmer_kernel.cu:84-87: an invalid visited window of read i adds 1 to the LAST bin of
// read i-1): inside a tile it moves one lane group to the left; across tiles the same hand-off protocol
// as the warp kernels of kernels.cu -- the tile that owns the read records its invalid windows in a word
// per tile boundary, spill_fixup_kernel adds them to the previous tile's last row after the kernel; the
// last tile of a launch scans the first read of the next range itself.
#include "kernels.h"
#include "kmer_device.cuh"
#include "dense_args.h"

#include <cstdlib>

namespace cfrk {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kEven = 0x55555555u;

// ---- "spread" validity: base j of a 16-base block <-> bit 2*(15-j) (the low bit of its 2-bit code field)
// positions [p, 16) set; p in [0, 16]
__device__ __forceinline__ uint32_t from_pos_s(int p) { return kEven >> min(2 * p, 31); }

// 4 bases (one little-endian word) -> 8 bits of codes (first base in bits 7:6, as encode4) and an 8-bit
// validity field with bits 6, 4, 2, 0 for bases 0..3.  Same classification as kmer_device.cuh encode4
// (src/fastaIO.h:123-139); only the multiplier that gathers the four "byte is valid" bits differs.
template <int FMT>
__device__ __forceinline__ void encode4_s(uint32_t w, uint32_t& codes8, uint32_t& valid8)
{
    uint32_t x, z;
    if (FMT == FMT_ASCII) {
        x = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
        const uint32_t sel = (x | (x >> 12)) & 0x3333u;
        const uint32_t expect = __byte_perm(0x54474341u /* "ACGT" */, 0u, sel);
        const uint32_t d = (__byte_perm(w, 0u, 0x3120u) & 0xDFDFDFDFu) ^ expect;
        z = ~(((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
        // bits 7 (base 0), 23 (base 1), 15 (base 2), 31 (base 3) -> product bits 38, 36, 34, 32
        valid8 = __umulhi(z, (1u << 31) | (1u << 13) | (1u << 19) | (1u << 1)) & 0x55u;
    } else {
        x = w & 0x03030303u;
        z = ~w & 0x80808080u;
        // bits 7, 15, 23, 31 (bases 0..3) -> product bits 38, 36, 34, 32
        valid8 = __umulhi(z, (1u << 31) | (1u << 21) | (1u << 11) | (1u << 1)) & 0x55u;
    }
    codes8 = (x * 0x40100401u) >> 24;
}

__device__ __forceinline__ uint32_t spread16(uint32_t v)   // bit i -> bit 2i
{
    uint32_t x = v & 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// one block: codes (base j in bits 31-2j:30-2j) and spread validity
template <int FMT>
__device__ __forceinline__ void encode16_s(const uint4 v, uint32_t& codes, uint32_t& valid)
{
    if (FMT == FMT_PACKED) { codes = v.x; valid = spread16(v.y); return; }
    constexpr int F = FMT == FMT_PACKED ? FMT_CODES : FMT;
    uint32_t c0, c1, c2, c3, m0, m1, m2, m3;
    encode4_s<F>(v.x, c0, m0);
    encode4_s<F>(v.y, c1, m1);
    encode4_s<F>(v.z, c2, m2);
    encode4_s<F>(v.w, c3, m3);
    codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    valid = (m0 << 24) | (m1 << 16) | (m2 << 8) | m3;
}

// ---- shared-memory / TMA helpers
__device__ __forceinline__ uint4 lds128(uint32_t saddr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
    return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr)
{
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr));
    return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t saddr)
{
    uint16_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(saddr));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared, completion on an mbarrier.  16-byte aligned addresses and size.
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

// ---- geometry per (K, SPLIT)
template <int K, int SPLIT, int FMT>
struct LaneGeo {
    static constexpr int BINS = 1 << (2 * K);
    static constexpr int R = 32 / SPLIT;                       // reads (rows) per warp tile
    static constexpr int OUT_BYTES = R * BINS * 4;             // the tile's rows, contiguous = one bulk store
    static constexpr int ROW_ALIGN = BINS * 4 < 128 ? 128 : BINS * 4;
    // staging: bytes of the bases buffer one tile may span (reads + separators + header lines of a FASTA
    // span): 272 per read on average covers 250-bp reads with short headers; longer -> warp-wide path.
    // Packed reads need 6 bytes per 16 bases: the same span in 104 bytes per read -- more warps per SM.
    static constexpr int SPAN = FMT == FMT_PACKED ? R * 104 : R * 272;
    static constexpr int WARP_BYTES = OUT_BYTES + SPAN;
};

template <int K>
__device__ __forceinline__ void count_planes(uint32_t codes, uint32_t pcodes, uint32_t good, uint32_t (&cnt)[1 << (2 * K)])
{
    static_assert(K == 1 || K == 2, "bit-plane counting is for 4 and 16 bins");
    const uint32_t h2 = codes >> 1;
    if (K == 1) {
        cnt[0] += __popc(~codes & ~h2 & good);
        cnt[1] += __popc(codes & ~h2 & good);
        cnt[2] += __popc(~codes & h2 & good);
        cnt[3] += __popc(codes & h2 & good);
    } else {
        // window ending at base j = (base j-1, base j): bring base j-1's code under base j's bit pair
        const uint32_t c1 = __funnelshift_r(codes, pcodes, 2), h1 = c1 >> 1;
        const uint32_t m0 = ~c1 & ~h1 & good, m1 = c1 & ~h1 & good, m2 = ~c1 & h1 & good, m3 = c1 & h1 & good;
        const uint32_t m[4] = {m0, m1, m2, m3};
#pragma unroll
        for (int a = 0; a < 4; a++) {
            cnt[4 * a + 0] += __popc(m[a] & ~codes & ~h2);
            cnt[4 * a + 1] += __popc(m[a] & codes & ~h2);
            cnt[4 * a + 2] += __popc(m[a] & ~codes & h2);
            cnt[4 * a + 3] += __popc(m[a] & codes & h2);
        }
    }
}

// Two blocks at once: the masks of a block use the even bits only (one per base), so block B's go to the odd bits
// and every POPC counts 32 bases: per 32 bases 16 (k = 2) / 4 (k = 1) LOP3 + POPC + add instead of twice that.
template <int K>
__device__ __forceinline__ void count_planes2(uint32_t codesA, uint32_t pcodesA, uint32_t goodA, uint32_t codesB,
                                              uint32_t pcodesB, uint32_t goodB, uint32_t (&cnt)[1 << (2 * K)])
{
    static_assert(K == 1 || K == 2, "bit-plane counting is for 4 and 16 bins");
    // planes of the window's LAST base: low / high code bit of base j at bit 2(15-j) (A) and 2(15-j)+1 (B)
    const uint32_t lo = (codesA & kEven) + 2u * (codesB & kEven);
    const uint32_t hi = ((codesA >> 1) & kEven) + 2u * ((codesB >> 1) & kEven);
    const uint32_t good = goodA + 2u * goodB;
    if (K == 1) {
        cnt[0] += __popc(~lo & ~hi & good);
        cnt[1] += __popc(lo & ~hi & good);
        cnt[2] += __popc(~lo & hi & good);
        cnt[3] += __popc(lo & hi & good);
    } else {
        // ... and of its FIRST base (base j-1 brought under base j's bit pair)
        const uint32_t c1A = __funnelshift_r(codesA, pcodesA, 2), c1B = __funnelshift_r(codesB, pcodesB, 2);
        const uint32_t lo1 = (c1A & kEven) + 2u * (c1B & kEven);
        const uint32_t hi1 = ((c1A >> 1) & kEven) + 2u * ((c1B >> 1) & kEven);
        const uint32_t m[4] = {~lo1 & ~hi1 & good, lo1 & ~hi1 & good, ~lo1 & hi1 & good, lo1 & hi1 & good};
        const uint32_t s2[4] = {~lo & ~hi, lo & ~hi, ~lo & hi, lo & hi};
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) cnt[4 * a + c] += __popc(m[a] & s2[c]);
    }
}

// k-mer windows of one block into a shared-memory row aligned to its own size (cf. emit_item)
template <int K>
__device__ __forceinline__ void count_row(uint32_t codes, uint32_t pcodes, uint32_t good, uint32_t row_saddr)
{
    constexpr uint32_t IDX_MASK = (1u << (2 * K)) - 1u;
    // every window of the block counts in every lane (blocks inside clean reads: 8 of the 10-11 rounds of a 150-bp
    // tile): 16 plain reductions -- the predicated form below compiles to a branch around every RED
    // (k = 3: 1401 -> 1577 Gbases/s; k = 4 with 4 lanes per read measured 770 -> 751: left as it was)
    if (K == 3 && __all_sync(kFull, good == kEven)) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int bit = 15 - j;
            const uint32_t t = bit >= 1 ? __funnelshift_r(codes, pcodes, 2 * bit - 2) : (codes << 2);
            const uint32_t addr = (t & (IDX_MASK << 2)) | row_saddr;
            asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
        }
        return;
    }
    if (good) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int bit = 15 - j;
            const uint32_t t = bit >= 1 ? __funnelshift_r(codes, pcodes, 2 * bit - 2) : (codes << 2);
            const uint32_t addr = (t & (IDX_MASK << 2)) | row_saddr;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}"
                         :: "r"(addr), "r"(good & (1u << (2 * bit))) : "memory");
        }
    }
}

// One block of one read for one lane of a W-lane group (W = 1, 2, 4: the tile's lane groups; W = 32: the
// warp-wide path).  The group's lanes hold consecutive blocks; block b-1 is the previous lane's, or -- for
// the group's first lane -- the last lane's block of the previous round (carry_*).
//   out: codes / pcodes (for the index), good = window ends of this block that count (spread bits)
template <int K, int FMT, int W>
__device__ __forceinline__ void lane_block(bool live, const uint4 raw, int t0, int tend, int mode, uint32_t& carry_c,
                                           uint32_t& carry_v, uint32_t& codes, uint32_t& pcodes, uint32_t& good, int& nbad)
{
    uint32_t valid = 0, cmask = 0;
    codes = 0;
    // packed words: a block without N (validity 0xFFFF: nearly all) needs no bit spreading -- a warp-uniform branch,
    // 2 instructions instead of 14
    bool all_valid = false;
    if (FMT == FMT_PACKED) all_valid = !__any_sync(kFull, live && raw.y != 0xFFFFu);
    if (W == 1) {
        // a lane per read: when this block lies inside its read in EVERY lane, without N here or in the block before
        // (8 of the 10-11 blocks of a 150-bp tile), all 16 window ends count and nothing else is to be found out
        uint32_t c = raw.x, v = kEven;
        if (FMT != FMT_PACKED) encode16_s<FMT>(raw, c, v);
        const bool inside = live && t0 >= K - 1 && t0 + 16 <= tend && carry_v == kEven &&
                            (FMT == FMT_PACKED ? raw.y == 0xFFFFu : v == kEven);
        if (__all_sync(kFull, inside)) {
            codes = c; pcodes = carry_c; good = kEven;
            carry_c = c;
            return;
        }
    }
    if (live) {
        if (FMT == FMT_PACKED && all_valid) { codes = raw.x; valid = kEven; }
        else encode16_s<FMT>(raw, codes, valid);
        if (t0 >= K - 1 && t0 + 16 <= tend) {
            cmask = kEven;                                                  // a block inside the read: every window end counts
        } else {
            const uint32_t upto = ~from_pos_s(min(16, tend - t0));
            valid &= from_pos_s(max(0, -t0)) & upto;                        // bases of this read only
            cmask = from_pos_s(min(16, max(0, K - 1 - t0))) & upto;         // window ends that are visited
        }
    }
    uint32_t pvalid;
    if (W == 1) {
        pcodes = carry_c; pvalid = carry_v;
    } else {
        const uint32_t up_c = __shfl_up_sync(kFull, codes, 1, W), up_v = __shfl_up_sync(kFull, valid, 1, W);
        const uint32_t last_c = __shfl_sync(kFull, carry_c, W - 1, W), last_v = __shfl_sync(kFull, carry_v, W - 1, W);
        const bool first = ((threadIdx.x & 31) & (W - 1)) == 0;
        pcodes = first ? last_c : up_c;
        pvalid = first ? last_v : up_v;
    }
    carry_c = codes; carry_v = valid;
    // bit of base j <- bases j-K+1 .. j all valid
    uint32_t ok = valid;
#pragma unroll
    for (int i = 1; i < K; i++) ok &= __funnelshift_r(valid, pvalid, 2 * i);
    good = ok & cmask;
    if (mode == MODE_COMPAT) {
        const uint32_t bad = ~ok & cmask;
        if (bad) nbad += __popc(bad);       // (clean reads: only the windows that reach the terminator)
    }
}

struct Stage {            // where the tile's blocks are: shared memory (staged) or the global buffer
    uint32_t sbase;       // shared address of block `blk_lo` (bytes), or of its codes word (packed)
    uint32_t svalid;      // packed: shared address of block blk_lo's validity half-word
    int64_t blk_lo;
};

template <int FMT, bool STAGED>
__device__ __forceinline__ uint4 fetch_block(const BasesRef& bases, const Stage& sg, int64_t blk)
{
    if (!STAGED) return load_block<FMT>(bases, blk);
    const uint32_t d = (uint32_t)(blk - sg.blk_lo);
    if (FMT == FMT_PACKED) return make_uint4(lds32(sg.sbase + d * 4u), lds16(sg.svalid + d * 2u), 0u, 0u);
    return lds128(sg.sbase + d * 16u);
}

}  // namespace

// ------------------------------------------------------------------------------------------
template <int K, int FMT, int SPLIT, int WARPS, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB) dense_lane_kernel(const DenseArgs a)
{
    using G = LaneGeo<K, SPLIT, FMT>;
    constexpr int BINS = G::BINS, R = G::R;
    constexpr bool PLANES = K <= 2;
    static_assert(!PLANES || SPLIT == 1, "register rows belong to one lane");
    static_assert(SPLIT == 1 || SPLIT == 2 || SPLIT == 4, "lane groups of 1, 2 or 4");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t raw_saddr = (uint32_t)__cvta_generic_to_shared(smem_raw);
    unsigned char* smem = smem_raw + ((G::ROW_ALIGN - (raw_saddr & (G::ROW_ALIGN - 1))) & (G::ROW_ALIGN - 1));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane / SPLIT, g = lane % SPLIT;
    uint32_t* out_s = reinterpret_cast<uint32_t*>(smem + (size_t)warp * G::OUT_BYTES);         // [R][BINS]
    unsigned char* stage = smem + (size_t)WARPS * G::OUT_BYTES + (size_t)warp * G::SPAN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * G::WARP_BYTES);
    const uint32_t out_saddr = (uint32_t)__cvta_generic_to_shared(out_s);
    const uint32_t stage_saddr = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + warp);
    const BasesRef bases{a.bases, a.valid};
    const bool compat = a.mode == MODE_COMPAT;

    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    int64_t tile = (int64_t)blockIdx.x * WARPS + warp;
    if (tile >= a.num_tiles) return;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_proxy_shared();
    __syncwarp();
    uint32_t parity = 0;

    // packed input: blocks per staging buffer (4-byte codes + 2-byte validity each), multiple of 8
    constexpr int kPackedBlocks = (G::SPAN / 6) / 8 * 8;

    // (start, length) of the tile's reads -- lane group q <-> read r0 + q; always one tile ahead
    int64_t s = 0; int len = 0;
    auto load_meta = [&](int64_t t) {
        const int64_t r = a.read_begin + t * R + q;
        s = 0; len = 0;
        if (r < a.read_end) { s = a.start[r]; len = a.length[r]; }
    };
    load_meta(tile);

    for (; tile < a.num_tiles;) {
        const int64_t r0 = a.read_begin + tile * R;
        const int nrows = (int)min((int64_t)R, a.read_end - r0);
        const bool have = q < nrows;
        int tend = 0, extra = 0;
        const ChunkScope cs{a.start, a.nS, a.nN, a.chunk_size, a.index_base};
        if (have) read_extent<K>(a.mode, len, compat_avail(cs, r0 + q, s, len, a.mode), tend, extra);
        const int64_t blk0 = s >> 4;
        const int off = (int)(s & 15);
        const int nblk = tend > 0 ? (int)(((s + tend - 1) >> 4) - blk0 + 1) : 0;

        // chunk openers (their spill is dropped): tile-local reads first, first + period, ...
        int first = -1, period = 1;
        if (a.chunk_size > 0) {
            const int64_t phase = (a.index_base + r0) % a.chunk_size;
            const int64_t f = phase == 0 ? 0 : a.chunk_size - phase;
            if (f <= R) { first = (int)f; period = (int)min(a.chunk_size, (int64_t)1 << 20); }
        } else if (r0 == 0) {
            first = 0; period = 1 << 20;
        }
        // (one opener per tile unless the chunks are shorter than a tile: no division in the common case)
        const bool many = period <= R;
        auto opens = [&](int qq) { return qq == first || (many && first >= 0 && qq > first && (qq - first) % period == 0); };
        // The spill of the read after the tile lands in the tile's last row.  Inside a launch the tile that
        // OWNS that read records it (hand-off word, added by spill_fixup_kernel); the last tile of a launch
        // scans that read itself (the owner is another launch).
        const bool scan_next = compat && (a.handoff == nullptr || tile == a.num_tiles - 1) && (r0 + nrows < a.nS) && !opens(nrows);
        int64_t sn = 0;
        int ex_next = 0, tend_next = 0;
        if (scan_next) {
            sn = a.start[r0 + nrows];
            const int ln = a.length[r0 + nrows];
            read_extent<K>(a.mode, ln, compat_avail(cs, r0 + nrows, sn, ln, a.mode), tend_next, ex_next);
        }

        // span of the tile in 16-byte blocks, relative to the first read with blocks
        const uint32_t with_blocks = __ballot_sync(kFull, nblk > 0);
        const int64_t base_blk = __shfl_sync(kFull, blk0, with_blocks ? __ffs(with_blocks) - 1 : 0);
        const int64_t dl = blk0 - base_blk, dh = dl + nblk;
        const int big = 1 << 28;
        const int lo32 = nblk > 0 ? (int)max((int64_t)-big, min((int64_t)big, dl)) : big;
        const int hi32 = nblk > 0 ? (int)max((int64_t)-big, min((int64_t)big, dh)) : -big;
        int span_lo = __reduce_min_sync(kFull, lo32), span_hi = __reduce_max_sync(kFull, hi32);
        bool staged = with_blocks != 0u;
        if (FMT == FMT_PACKED) {
            // 16-byte aligned pieces of codes[] and valid[]: whole groups of 8 blocks, inside the arrays
            const bool sane = staged && span_lo > -big && span_hi < big;
            const int64_t abs_lo = sane ? ((base_blk + span_lo) & ~(int64_t)7) : 0;
            const int64_t abs_hi = sane ? ((base_blk + span_hi + 7) & ~(int64_t)7) : 0;
            staged = sane && abs_hi - abs_lo <= kPackedBlocks && abs_hi <= ((a.nN + 15) >> 4) &&
                     ((reinterpret_cast<uintptr_t>(a.bases) | reinterpret_cast<uintptr_t>(a.valid)) & 15) == 0;
            span_lo = (int)(abs_lo - base_blk);
            span_hi = (int)(abs_hi - base_blk);
        } else {
            staged = staged && (int64_t)(span_hi - span_lo) * 16 <= G::SPAN && span_lo > -big && span_hi < big;
        }
        Stage sg;
        sg.blk_lo = base_blk + span_lo;
        sg.sbase = stage_saddr;
        sg.svalid = stage_saddr + kPackedBlocks * 4;
        if (staged && lane == 0) {
            fence_async_proxy_shared();     // the previous tile's reads of the buffer come before this write
            const uint32_t nb = (uint32_t)(span_hi - span_lo);
            if (FMT == FMT_PACKED) {
                mbar_expect_tx(bar, nb * 6u);
                bulk_load(sg.sbase, reinterpret_cast<const uint32_t*>(a.bases) + sg.blk_lo, nb * 4u, bar);
                bulk_load(sg.svalid, a.valid + sg.blk_lo, nb * 2u, bar);
            } else {
                mbar_expect_tx(bar, nb * 16u);
                bulk_load(sg.sbase, a.bases + sg.blk_lo * 16, nb * 16u, bar);
            }
        }

        // the bulk store of the previous tile has read the row buffer: clear it (register rows are
        // written whole, so the staged bit-plane path needs no clearing)
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        if (!PLANES || !staged) {
            uint4* o4 = reinterpret_cast<uint4*>(out_s);
#pragma unroll 4
            for (int i = lane; i < G::OUT_BYTES / 16; i += 32) o4[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
        }

        // next tile's offsets are in flight while this one is counted
        const int64_t next_tile = tile + nwarps;
        if (next_tile < a.num_tiles) load_meta(next_tile);

        uint32_t cnt[PLANES ? BINS : 1];
#pragma unroll
        for (int i = 0; i < (PLANES ? BINS : 1); i++) cnt[i] = 0u;
        int carry0 = 0;          // data-dependent invalid windows of the tile's first read (owed to the previous tile)
        const uint32_t row_saddr = out_saddr + (uint32_t)q * (BINS * 4);

        if (staged) {
            mbar_wait(bar, parity);
            parity ^= 1u;
            // ---- lane group per read
            const int rounds = __reduce_max_sync(kFull, (nblk + SPLIT - 1) / SPLIT);
            uint32_t carry_c = 0, carry_v = 0;
            int nbad = 0;
            uint4 raw_next = make_uint4(0u, 0u, 0u, 0u);
            if constexpr (PLANES && SPLIT == 1) {
                // a lane per read, two blocks per step (count_planes2); the next two are on their way from shared memory
                const int steps = rounds >> 1;          // an odd last block (11 blocks for most 150-bp tiles) goes alone
                uint4 raw_next2 = make_uint4(0u, 0u, 0u, 0u);
                if (0 < nblk) raw_next = fetch_block<FMT, true>(bases, sg, blk0);
                if (1 < nblk) raw_next2 = fetch_block<FMT, true>(bases, sg, blk0 + 1);
                for (int i = 0; i < steps; i++) {
                    const int b = 2 * i;
                    const uint4 rawA = raw_next, rawB = raw_next2;
                    if (b + 2 < nblk) raw_next = fetch_block<FMT, true>(bases, sg, blk0 + b + 2);
                    if (b + 3 < nblk) raw_next2 = fetch_block<FMT, true>(bases, sg, blk0 + b + 3);
                    uint32_t cA, pA, gA, cB, pB, gB;
                    lane_block<K, FMT, 1>(b < nblk, rawA, b * 16 - off, tend, a.mode, carry_c, carry_v, cA, pA, gA, nbad);
                    lane_block<K, FMT, 1>(b + 1 < nblk, rawB, b * 16 + 16 - off, tend, a.mode, carry_c, carry_v, cB, pB, gB, nbad);
                    count_planes2<K>(cA, pA, gA, cB, pB, gB, cnt);
                }
                if (rounds & 1) {
                    const int b = 2 * steps;
                    uint32_t codes, pcodes, good;
                    lane_block<K, FMT, 1>(b < nblk, raw_next, b * 16 - off, tend, a.mode, carry_c, carry_v, codes, pcodes, good, nbad);
                    count_planes<K>(codes, pcodes, good, cnt);
                }
            } else {
            if (g < nblk) raw_next = fetch_block<FMT, true>(bases, sg, blk0 + g);
            for (int i = 0; i < rounds; i++) {
                const int b = i * SPLIT + g;
                const bool live = b < nblk;
                const uint4 raw = raw_next;
                // the next round's block is on its way from shared memory while this one is counted
                if (b + SPLIT < nblk) raw_next = fetch_block<FMT, true>(bases, sg, blk0 + b + SPLIT);
                uint32_t codes, pcodes, good;
                lane_block<K, FMT, SPLIT>(live, raw, b * 16 - off, tend, a.mode, carry_c, carry_v, codes, pcodes, good, nbad);
                if constexpr (PLANES) count_planes<K>(codes, pcodes, good, cnt);
                else count_row<K>(codes, pcodes, good, row_saddr);
            }
            }
            if (compat) {
#pragma unroll
                for (int d = 1; d < SPLIT; d <<= 1) nbad += __shfl_xor_sync(kFull, nbad, d);
                const bool drop = !have || opens(q);
                const int inv = drop ? 0 : nbad + extra;      // to the last bin of read q - 1
                if (q == 0) carry0 = inv;                     // ... of the previous tile: through the hand-off word
                if constexpr (PLANES) {
                    const int from_right = __shfl_down_sync(kFull, inv, 1);
                    if (lane + 1 < nrows) cnt[BINS - 1] += (uint32_t)from_right;
                } else {
                    if (g == 0 && q >= 1 && inv > 0)
                        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(row_saddr - 4u), "r"((uint32_t)inv) : "memory");
                }
            }
            if constexpr (PLANES) {
                uint4* o4 = reinterpret_cast<uint4*>(out_s + lane * BINS);
#pragma unroll
                for (int i = 0; i < BINS / 4; i++) o4[i] = make_uint4(cnt[4 * i], cnt[4 * i + 1], cnt[4 * i + 2], cnt[4 * i + 3]);
            }
        } else {
            // ---- warp-wide: read after read, lanes take consecutive blocks, coalesced global loads
            for (int qq = 0; qq < nrows; qq++) {
                const int src = qq * SPLIT;
                const int64_t bq = __shfl_sync(kFull, blk0, src);
                const int offq = __shfl_sync(kFull, off, src), tendq = __shfl_sync(kFull, tend, src);
                const int nblkq = __shfl_sync(kFull, nblk, src), extraq = __shfl_sync(kFull, extra, src);
                uint32_t carry_c = 0, carry_v = 0;
                int nbad = 0;
#pragma unroll
                for (int i = 0; i < (PLANES ? BINS : 1); i++) cnt[i] = 0u;
                const uint32_t rowq = out_saddr + (uint32_t)qq * (BINS * 4);
                for (int b0 = 0; b0 < nblkq; b0 += 32) {
                    const int b = b0 + lane;
                    const bool live = b < nblkq;
                    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                    if (live) raw = fetch_block<FMT, false>(bases, sg, bq + b);
                    uint32_t codes, pcodes, good;
                    lane_block<K, FMT, 32>(live, raw, b * 16 - offq, tendq, a.mode, carry_c, carry_v, codes, pcodes, good, nbad);
                    if constexpr (PLANES) count_planes<K>(codes, pcodes, good, cnt);
                    else count_row<K>(codes, pcodes, good, rowq);
                }
                if constexpr (PLANES) {
#pragma unroll
                    for (int i = 0; i < BINS; i++)
                        if (cnt[i]) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(rowq + 4u * i), "r"(cnt[i]) : "memory");
                }
                if (compat) {
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) nbad += __shfl_xor_sync(kFull, nbad, d);
                    const bool drop = opens(qq);
                    if (qq == 0) carry0 = drop ? 0 : nbad + extraq;
                    else if (!drop && lane == 0 && nbad + extraq > 0)
                        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(rowq - 4u), "r"((uint32_t)(nbad + extraq)) : "memory");
                }
            }
        }

        if (scan_next && tend_next > 0) {
            // last tile of the launch: nobody records the data-dependent spill of the read after it -- scan
            // that read here (warp-wide, nothing counted)
            const int64_t bq = sn >> 4;
            const int offq = (int)(sn & 15);
            const int nblkq = (int)(((sn + tend_next - 1) >> 4) - bq + 1);
            uint32_t carry_c = 0, carry_v = 0;
            int nbad = 0;
            for (int b0 = 0; b0 < nblkq; b0 += 32) {
                const int b = b0 + lane;
                const bool live = b < nblkq;
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (live) raw = fetch_block<FMT, false>(bases, sg, bq + b);
                uint32_t codes, pcodes, good;
                lane_block<K, FMT, 32>(live, raw, b * 16 - offq, tend_next, a.mode, carry_c, carry_v, codes, pcodes, good, nbad);
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) nbad += __shfl_xor_sync(kFull, nbad, d);
            __syncwarp();
            if (lane == 0 && nbad + ex_next > 0)
                asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(out_saddr + (uint32_t)(nrows * BINS - 1) * 4u), "r"((uint32_t)(nbad + ex_next)) : "memory");
        }

        // rows -> HBM: one TMA bulk store per tile
        fence_async_proxy_shared();
        __syncwarp();
        if (lane == 0) {
            bulk_store_tile(a.out + (r0 - a.read_begin) * BINS, out_s, (uint32_t)nrows * BINS * 4u);
            if (compat && a.handoff != nullptr && carry0 > 0 && tile != 0) a.handoff[tile - 1] = (uint32_t)carry0;   // (0 for a chunk opener)
        }
        tile = next_tile;
    }
    if (lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
template <int K, int FMT, int SPLIT, int WARPS, int MINB = 1>
static cudaError_t launch_lane_t(const DenseArgs& a0, cudaStream_t st)
{
    using G = LaneGeo<K, SPLIT, FMT>;
    auto kern = dense_lane_kernel<K, FMT, SPLIT, WARPS, MINB>;
    constexpr int smem = WARPS * G::WARP_BYTES + WARPS * 8 + G::ROW_ALIGN;
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
    a.num_tiles = (a.read_end - a.read_begin + G::R - 1) / G::R;
    if (a.num_tiles <= 0) return cudaSuccess;
    const int64_t ctas_needed = (a.num_tiles + WARPS - 1) / WARPS;
    const int64_t resident = (int64_t)num_sms * ctas_per_sm;
    const unsigned grid = (unsigned)(ctas_needed < resident ? ctas_needed : resident);
    a.handoff = nullptr;
    if (a.mode == MODE_COMPAT && a.num_tiles > 1) {
        // one word per tile boundary, zeroed for this launch
        if ((e = stream_scratch(st, (size_t)a.num_tiles * 4, reinterpret_cast<void**>(&a.handoff))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(a.handoff, 0, (size_t)a.num_tiles * 4, st)) != cudaSuccess) return e;
    }
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    count_launch();
    e = cudaGetLastError();
    if (a.handoff) {
        const int64_t blocks = (a.num_tiles + 255) / 256;
        (void)blocks;
        const cudaError_t e2 = launch_spill_fixup(a.handoff, a.num_tiles, a.out, (int64_t)G::R * G::BINS, st);
        if (e == cudaSuccess) e = e2;
    }
    return e;
}

static int lane_env(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// lane groups per k: measured choices (profiles/r2_notes.md); the environment switches re-run the A/B
static int lane_split(int k, int fmt)
{
    // k = 3 from packed reads: the small staging buffer leaves room for a lane per read (32-read tiles: the
    // tile set-up is paid once per 32 reads instead of 16) at 20 warps per SM
    static const int s3 = lane_env("CFRK_LANE_SPLIT_K3", 0), s4 = lane_env("CFRK_LANE_SPLIT_K4", 4);
    return k <= 2 ? 1 : (k == 3 ? (s3 ? s3 : (fmt == FMT_PACKED ? 1 : 2)) : s4);
}

// granularity for callers that cut a batch into ranges: a multiple of the tile of every layout
int dense_lane_reads_per_tile(int k) { return k <= 3 ? 32 : 32 / lane_split(k, FMT_ASCII); }

template <int FMT>
static cudaError_t launch_lane_fmt(int k, const DenseArgs& a, cudaStream_t st)
{
    const int sp = lane_split(k, FMT);
    switch (k) {
    case 1: return launch_lane_t<1, FMT, 1, 4>(a, st);
    case 2: {
        static const int minb = lane_env("CFRK_LANE_MINB_K2", 1);      // 8: cap at 64 registers (A/B)
        return minb == 8 ? launch_lane_t<2, FMT, 1, 4, 8>(a, st) : launch_lane_t<2, FMT, 1, 4>(a, st);
    }
    case 3:
        if (sp == 1) return launch_lane_t<3, FMT, 1, 4>(a, st);
        if (sp == 4) return launch_lane_t<3, FMT, 4, 4>(a, st);
        return launch_lane_t<3, FMT, 2, 4>(a, st);
    case 4:
        if (sp == 1) return launch_lane_t<4, FMT, 1, 2>(a, st);
        if (sp == 2) return launch_lane_t<4, FMT, 2, 4>(a, st);
        return launch_lane_t<4, FMT, 4, 4>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_dense_lane(int k, int fmt, const DenseArgs& a, cudaStream_t st)
{
    if (fmt == FMT_PACKED) return launch_lane_fmt<FMT_PACKED>(k, a, st);
    return fmt == FMT_ASCII ? launch_lane_fmt<FMT_ASCII>(k, a, st) : launch_lane_fmt<FMT_CODES>(k, a, st);
}

}  // namespace cfrk
