// stream_device.cuh -- a span of bases as a 2-bit stream + validity bits in shared memory, and the
// extraction of consecutive k-mer windows from it.  Shared by the sparse per-read path (sparse.cu)
// and the partitioned whole-dataset histogram (hist_split.cu).
#pragma once

#include "kmer_device.cuh"

namespace cfrk {

template <typename KeyT> struct KeyMax;
template <> struct KeyMax<uint32_t> { static constexpr uint32_t value = 0xFFFFFFFFu; };
template <> struct KeyMax<uint64_t> { static constexpr uint64_t value = 0xFFFFFFFFFFFFFFFFull; };

struct WarpStream {
    uint32_t* cw;   // 2-bit codes, 16 bases per word, first base in the top bits
    uint16_t* vh;   // validity, 16 bases per half-word; stored so that 32-bit loads see 32 positions MSB-first
};

// E consecutive windows starting at stream position P0: the 48 bases and 64 validity bits behind P0
// are pulled into registers once (4 + 3 shared-memory words), every window is then two funnel
// shifts.  A window that is not entirely inside the read is invalid by the validity masks alone
// (positions outside the read are masked when the stream is built), so no window count is needed.
// Invalid windows get the maximum key.  Returns the valid windows of this lane as a bit mask.
template <typename KeyT, int E>
__device__ __forceinline__ uint32_t extract_windows(const WarpStream& st, int P0, int k, KeyT (&key)[E])
{
    static_assert(E <= 16, "2*e must stay below 32");
    const int b = P0 >> 4, o = (P0 & 15) * 2;
    const uint32_t w0 = st.cw[b], w1 = st.cw[b + 1], w2 = st.cw[b + 2], w3 = st.cw[b + 3];
    const uint32_t x0 = __funnelshift_l(w1, w0, o), x1 = __funnelshift_l(w2, w1, o), x2 = __funnelshift_l(w3, w2, o);
    const uint32_t* vw = reinterpret_cast<const uint32_t*>(st.vh);
    const int c = P0 >> 5, vo = P0 & 31;
    const uint32_t u0 = vw[c], u1 = vw[c + 1], u2 = vw[c + 2];
    const uint32_t v0 = __funnelshift_l(u1, u0, vo), v1 = __funnelshift_l(u2, u1, vo);
    const uint32_t need = 0xFFFFFFFFu << (32 - k);
    uint32_t valid = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t hi = __funnelshift_l(x1, x0, 2 * e);
        KeyT kk;
        if (sizeof(KeyT) == 4) {
            kk = (KeyT)(hi >> (32 - 2 * k));
        } else {
            const uint32_t lo = __funnelshift_l(x2, x1, 2 * e);
            kk = (KeyT)((((uint64_t)hi << 32) | lo) >> (64 - 2 * k));
        }
        const bool ok = (__funnelshift_l(v1, v0, e) & need) == need;
        key[e] = ok ? kk : KeyMax<KeyT>::value;
        valid |= ok ? 1u << e : 0u;
    }
    return valid;
}

}  // namespace cfrk
