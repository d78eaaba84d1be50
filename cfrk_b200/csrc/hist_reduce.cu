// hist_reduce.cu -- the one exchange step of the path: the whole-dataset histogram tables of the GPUs of one box,
// summed IN PLACE over NVLink peer memory by one kernel per GPU (no NCCL call, no staging buffer).
//
// Every rank holds a 4^k uint32 table in memory that its peers can address (symmetric allocations: the host side
// hands in the peers' addresses; NVSwitch gives every pair of GPUs the full link rate).  Two shots:
//   1. rank r SUMS slice r of every table (16-byte loads from the peers, own table included),
//   2. and WRITES the sums into slice r of every table (16-byte stores to the peers).
// Per GPU (W - 1) / W of the table crosses NVLink in each direction: 56 MiB of the 64 MiB table (k = 12) on 8 GPUs.
// Ranks meet twice through flag words in each other's memory (st.release.sys / ld.acquire.sys): before the sums (a
// peer's table is complete: its count kernel precedes this kernel in its stream) and after the writes.  Every CTA
// meets the CTA of the same index on every peer; the grid is small enough to be resident as a whole.
//
// The NVLS variant (multimem.ld_reduce / multimem.st on a multicast address: the switch adds, one load per element
// instead of W, half the link traffic on 8 GPUs) is used when the caller passes a multicast mapping; integer adds
// exist there for 32- and 64-bit words only (no vectors), so it moves 8 bytes per instruction (two counters: their
// sums do not carry, the total number of windows is below 2^32 -- the caller's contract).
// Measured, 64 MiB table (k = 12), ms per call, same run (tests/dist_hist_reduce.py):
//            peer loads/stores   NVLS     NCCL all-reduce
//   2 GPUs   0.126               0.195    0.150
//   8 GPUs   0.213               0.177    0.254
// 2 GPUs: 64 MiB cross each direction at best (0.09 ms at 750 GB/s); 8 GPUs: 112 MiB (p2p), 64 MiB (NVLS).
#include "../../include/cfrk_b200.h"
#include "kernels.h"
#include "internal.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>

namespace cfrk {
extern void count_launch();

namespace {

constexpr int kMaxRanks = CFRK_HIST_REDUCE_MAX_RANKS;
constexpr int kMaxCtas = 512;
constexpr int kThreads = 512;
// flag words of one rank: [2 meetings][kMaxCtas][kMaxRanks], then one status word
constexpr int kStatusWord = 2 * kMaxCtas * kMaxRanks;
static_assert((kStatusWord + 1) * 4 <= CFRK_HIST_REDUCE_FLAG_BYTES, "flag buffer too small");

struct Peers {
    uint32_t* table[kMaxRanks];
    uint32_t* flag[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_sys_v4(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_v4(uint32_t* p, uint4 v)
{
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// CTA b of this rank meets CTA b of every peer.  A peer that does not show up within ~2 s (a failed rank) is
// recorded in the status word instead of spinning for ever; the result is then undefined and the caller sees it.
__device__ __forceinline__ void meet(const Peers& p, int rank, int world, int meeting, uint32_t epoch)
{
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int t = threadIdx.x;
        const size_t slot = ((size_t)meeting * kMaxCtas + blockIdx.x) * kMaxRanks;
        __threadfence_system();
        st_release_sys(p.flag[t] + slot + rank, epoch);
        const uint32_t* mine = p.flag[rank] + slot + t;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > 4000000000ll) { p.flag[rank][kStatusWord] = 1u + (uint32_t)t; break; }
        }
    }
    __syncthreads();
}

template <int UNROLL>
__global__ void __launch_bounds__(kThreads, 2) hist_reduce_kernel(const Peers p, int rank, int world, int64_t n4, uint32_t epoch)
{
    meet(p, rank, world, 0, epoch);
    const int64_t begin = n4 * rank / world, end = n4 * (rank + 1) / world;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i0 = begin + (int64_t)blockIdx.x * kThreads + threadIdx.x; i0 < end; i0 += stride * UNROLL) {
        uint4 acc[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc[u] = make_uint4(0, 0, 0, 0);
        for (int q = 0; q < world; q++) {
            const uint32_t* src = p.table[(rank + q) % world];      // staggered: the ranks start on different peers
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int64_t i = i0 + u * stride;
                if (i < end) {
                    const uint4 v = ld_sys_v4(src + 4 * i);
                    acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
                }
            }
        }
        for (int q = 0; q < world; q++) {
            uint32_t* dst = p.table[(rank + q) % world];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int64_t i = i0 + u * stride;
                if (i < end) st_sys_v4(dst + 4 * i, acc[u]);
            }
        }
    }
    meet(p, rank, world, 1, epoch);
}

// NVLS: the switch adds.  mc = multicast address of the tables; 8 bytes (two counters) per instruction.  More
// instructions in flight per thread do not help (8 GPUs, 64 MiB: UNROLL 1 0.177 ms, 4 0.194, 8 0.206): the rate of
// 8-byte multimem operations is the limit, not their latency.
template <int UNROLL>
__global__ void __launch_bounds__(kThreads, 2) hist_reduce_nvls_kernel(const Peers p, uint64_t* mc, int rank, int world, int64_t n2,
                                                                       uint32_t epoch)
{
    meet(p, rank, world, 0, epoch);
    const int64_t begin = n2 * rank / world, end = n2 * (rank + 1) / world;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i0 = begin + (int64_t)blockIdx.x * kThreads + threadIdx.x; i0 < end; i0 += stride * UNROLL) {
        uint64_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = i0 + u * stride;
            if (i < end) asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v[u]) : "l"(mc + i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = i0 + u * stride;
            if (i < end) asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(mc + i), "l"(v[u]) : "memory");
        }
    }
    meet(p, rank, world, 1, epoch);
}

}  // namespace

cudaError_t launch_hist_reduce(void* const* tables, void* const* flags, int rank, int world, int64_t n_bins, uint32_t epoch,
                               void* multicast, cudaStream_t st)
{
    Peers p{};
    for (int i = 0; i < world; i++) {
        p.table[i] = static_cast<uint32_t*>(tables[i]);
        p.flag[i] = static_cast<uint32_t*>(flags[i]);
    }
    int dev = 0, sms = 148;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // every CTA must be resident (CTAs of the same index wait for each other across the GPUs): 2 per SM by the
    // launch bounds, the same grid on every rank
    static const int per_sm = [] { const char* ev = getenv("CFRK_HIST_REDUCE_CTAS"); return ev && atoi(ev) == 1 ? 1 : 2; }();
    static const int unroll = [] { const char* ev = getenv("CFRK_HIST_REDUCE_UNROLL"); return ev ? atoi(ev) : 2; }();
    const int grid = std::min(per_sm * sms, kMaxCtas);
    static const int nvls_unroll = [] { const char* ev = getenv("CFRK_HIST_REDUCE_NVLS_UNROLL"); return ev ? atoi(ev) : 1; }();
    if (multicast && nvls_unroll == 4)
        hist_reduce_nvls_kernel<4><<<grid, kThreads, 0, st>>>(p, static_cast<uint64_t*>(multicast), rank, world, n_bins / 2, epoch);
    else if (multicast)
        hist_reduce_nvls_kernel<1><<<grid, kThreads, 0, st>>>(p, static_cast<uint64_t*>(multicast), rank, world, n_bins / 2, epoch);
    else if (unroll == 1)
        hist_reduce_kernel<1><<<grid, kThreads, 0, st>>>(p, rank, world, n_bins / 4, epoch);
    else if (unroll == 4)
        hist_reduce_kernel<4><<<grid, kThreads, 0, st>>>(p, rank, world, n_bins / 4, epoch);
    else
        hist_reduce_kernel<2><<<grid, kThreads, 0, st>>>(p, rank, world, n_bins / 4, epoch);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cfrk

extern "C" int cfrk_hist_allreduce_device(void* const* peer_tables, void* const* peer_flags, int rank, int world,
                                          int64_t n_bins, uint32_t epoch, void* multicast_table, void* stream)
{
    if (!peer_tables || !peer_flags) { cfrk::set_last_error("null pointer array"); return CFRK_EINVAL; }
    if (world < 1 || world > CFRK_HIST_REDUCE_MAX_RANKS || rank < 0 || rank >= world) { cfrk::set_last_error("rank / world out of range (<= 8 ranks)"); return CFRK_EINVAL; }
    if (n_bins < 4 || (n_bins & 3)) { cfrk::set_last_error("n_bins must be a positive multiple of 4 (4^k, k >= 1)"); return CFRK_EINVAL; }
    if (epoch == 0) { cfrk::set_last_error("epoch starts at 1 (the flag words start at 0)"); return CFRK_EINVAL; }
    for (int i = 0; i < world; i++) {
        if (!peer_tables[i] || !peer_flags[i]) { cfrk::set_last_error("null peer pointer"); return CFRK_EINVAL; }
        if (reinterpret_cast<uintptr_t>(peer_tables[i]) & 15) { cfrk::set_last_error("tables must be 16-byte aligned"); return CFRK_EINVAL; }
    }
    if (world == 1) return CFRK_OK;
    const cudaError_t e = cfrk::launch_hist_reduce(peer_tables, peer_flags, rank, world, n_bins, epoch, multicast_table,
                                                   static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { cfrk::set_last_error(std::string("hist_reduce_kernel launch: ") + cudaGetErrorString(e)); return CFRK_ECUDA; }
    return CFRK_OK;
}
