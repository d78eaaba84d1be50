// api.cu -- the C ABI of libcfrk_b200.so (include/cfrk_b200.h) over the kernels.
//
// cfrk_count_dense_host is the operator that replaces the reference's kmer_main
// (src/kmer_main.cu:20-128): same inputs and output, but
//   * device buffers are cached per host thread instead of 5 cudaMalloc + 5 cudaFree per call
//     (src/kmer_main.cu:59-63,120-124),
//   * rows leave the device through a two-slot ring so that the device->host copy of slice i
//     overlaps the kernel of slice i+1 (the reference does one synchronous cudaMemcpy of the
//     whole Freq, src/kmer_main.cu:116),
//   * part of the rows is written by host threads from GPU-compacted (bin, count) pairs (below), so that the
//     PCIe link is not the only path into the caller's buffer,
//   * errors are returned, not printed.
#include "../../include/cfrk_b200.h"
#include "kernels.h"
#include "kmer_device.cuh"
#include "internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string t_err;

int fail(int code, const char* what)
{
    t_err = what;
    return code;
}
int fail_cuda(cudaError_t e, const char* where)
{
    t_err = std::string(where) + ": " + cudaGetErrorString(e);
    return CFRK_ECUDA;
}
#define CU(call)                                            \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call); \
    } while (0)

constexpr size_t kRingSlotBytes = (size_t)256 << 20;  // rows per kernel launch of the host path

// Per host thread: kmer_main is called concurrently from several pthreads in the reference
// driver (src/main.cu:279-289), each needs its own streams and scratch.  A context belongs to one
// thread at a time; when the thread exits the context goes back to a process-wide idle list and
// the next new thread takes it over, so thread churn does not grow the footprint (at most one
// context per CONCURRENT caller).  Nothing is freed at process exit (the CUDA runtime may already
// be gone); cfrk_release() frees the idle contexts, the calling thread's own and the pinned arena.
// Cached per context: the bases buffer of the largest call, start/length arrays, up to
// 2 x 256 MiB of row ring, 2 streams, 4 events.
struct HostCtx {
    int device = -1;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    void* d_bases = nullptr;  size_t cap_bases = 0;
    int64_t* d_start = nullptr; int32_t* d_length = nullptr; size_t cap_reads = 0;
    int32_t* d_ring[2] = {nullptr, nullptr}; size_t cap_ring = 0;
    // host-expanded part of a call: a device slot for its dense rows, their (bin, count) pairs (device +
    // pinned mirror), per-row offsets and counts, a stream for the small copies, events per slice
    cudaStream_t aux = nullptr;
    int32_t* d_hslot = nullptr; size_t cap_hslot = 0;
    uint32_t* d_pk = nullptr; uint32_t* d_pc = nullptr; size_t cap_pairs = 0;
    int64_t* d_poff = nullptr; int32_t* d_prc = nullptr; size_t cap_prows = 0;
    uint32_t* h_pk = nullptr; uint32_t* h_pc = nullptr; size_t cap_hpairs = 0;
    int32_t* h_prc = nullptr; size_t cap_hprc = 0;
    std::vector<cudaEvent_t> packed, ready;     // packed[0]: inputs uploaded; ready[s]: pairs of host slice s in host memory
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    double dma_frac[CFRK_DENSE_MAX_K + 1] = {0};   // share of the rows that goes through the DMA engine, per k (adaptive)

    void release()
    {
        if (device < 0) return;
        cudaSetDevice(device);
        cudaFree(d_bases); cudaFree(d_start); cudaFree(d_length);
        cudaFree(d_ring[0]); cudaFree(d_ring[1]);
        cudaFree(d_hslot); cudaFree(d_pk); cudaFree(d_pc); cudaFree(d_poff); cudaFree(d_prc);
        if (h_pk) cudaFreeHost(h_pk);
        if (h_pc) cudaFreeHost(h_pc);
        if (h_prc) cudaFreeHost(h_prc);
        for (cudaEvent_t e : packed) cudaEventDestroy(e);
        for (cudaEvent_t e : ready) cudaEventDestroy(e);
        if (aux) cudaStreamDestroy(aux);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        for (int i = 0; i < 2; i++) {
            if (done[i]) cudaEventDestroy(done[i]);
            if (drained[i]) cudaEventDestroy(drained[i]);
        }
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
        cudaGetLastError();
        *this = HostCtx();
    }
};

std::mutex g_ctx_mu;
std::vector<HostCtx*> g_idle_ctx;   // contexts whose thread has exited

struct CtxHolder {
    HostCtx* c = nullptr;
    ~CtxHolder()
    {
        if (!c) return;
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        g_idle_ctx.push_back(c);   // no CUDA call here: this runs during thread / process teardown
        c = nullptr;
    }
};
thread_local CtxHolder t_holder;

int ensure_ctx(int device, HostCtx** out)
{
    HostCtx*& c = t_holder.c;
    if (c && c->device == device) { CU(cudaSetDevice(device)); *out = c; return CFRK_OK; }
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (c) { g_idle_ctx.push_back(c); c = nullptr; }
        for (size_t i = 0; i < g_idle_ctx.size(); i++)
            if (g_idle_ctx[i]->device == device) {
                c = g_idle_ctx[i];
                g_idle_ctx.erase(g_idle_ctx.begin() + (long)i);
                break;
            }
    }
    CU(cudaSetDevice(device));
    if (!c) {
        c = new HostCtx();
        CU(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->drained[i], cudaEventDisableTiming));
        }
        CU(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
        CU(cudaEventCreate(&c->t0));
        CU(cudaEventCreate(&c->t1));
        c->device = device;
    }
    *out = c;
    return CFRK_OK;
}

template <class T>
int grow(T*& p, size_t& cap, size_t need)
{
    if (need <= cap) return CFRK_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want);
    if (e != cudaSuccess) { t_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return CFRK_ENOMEM; }
    cap = want;
    return CFRK_OK;
}

// ------------------------------------------------------------------------------------------
// Host side of the host-buffer operator.  The output of kmer_main's contract is 4^k int32 PER READ in
// HOST memory (256 KiB per 150-bp read at k = 8, of which <= 150 words are not zero) and a PCIe Gen5
// link moves 55 GB/s: with every row crossing the bus the operator ran at 0.12 Gbases/s on a k = 4..8
// sweep, below a 16-thread CPU counter.  So the rows are split: one share is written by
// the GPU's DMA engine as before; the rows of the other share are counted on the GPU like all the others
// (compat spill included), compacted there to the (bin, count) pairs of their non-zero bins
// (row_pairs.cu: <= 1.2 KB per 150-bp read), and `nt` host threads EXPAND the pairs into the caller's
// rows -- zeros and counts in one pass of streaming stores, no read-for-ownership.  The split adapts per
// k to whichever side finishes first.  Nothing is counted on the host.
std::atomic<int> g_host_threads{-2};    // -2: not configured yet

int host_threads()
{
    int n = g_host_threads.load();
    if (n == -2) {
        const char* e = getenv("CFRK_HOST_THREADS");
        n = e ? atoi(e) : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        if (n < 0) n = 0;
        g_host_threads.store(n);
    }
    return n < 0 ? 0 : n;
}

inline int64_t visited_windows(int len, int k, int mode)
{
    const int v = mode == CFRK_MODE_COMPAT ? std::min(len - 1, cfrk::kRefBlockThreads) : len - k + 1;
    return v > 0 ? v : 0;
}

class HostPool {
public:
    void run(int nparts, const std::function<void(int)>& fn)
    {
        std::unique_lock<std::mutex> lk(mu_);
        while ((int)th_.size() < nparts) {
            const int id = (int)th_.size();
            th_.emplace_back([this, id] { loop(id); });
        }
        busy_cv_.wait(lk, [&] { return fn_ == nullptr; });     // one batch at a time (several caller threads share the pool)
        fn_ = &fn; nparts_ = nparts; pending_ = nparts; gen_++;
        cv_.notify_all();
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
        busy_cv_.notify_one();
    }

private:
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (id >= nparts_) continue;
                fn = fn_;
            }
            (*fn)(id);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;    // never joined: the pool lives as long as the process
    std::mutex mu_;
    std::condition_variable cv_, done_cv_, busy_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int nparts_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
};
HostPool* g_pool = nullptr;
std::once_flag g_pool_once;

inline void stream_zero_line(int32_t* p)      // 64 bytes, 64-byte aligned
{
#if defined(__SSE2__)
    const __m128i z = _mm_setzero_si128();
    _mm_stream_si128(reinterpret_cast<__m128i*>(p), z);
    _mm_stream_si128(reinterpret_cast<__m128i*>(p) + 1, z);
    _mm_stream_si128(reinterpret_cast<__m128i*>(p) + 2, z);
    _mm_stream_si128(reinterpret_cast<__m128i*>(p) + 3, z);
#else
    memset(p, 0, 64);
#endif
}

// One row from the (bin, count) pairs of its non-zero bins (ascending bins, as row_pairs.cu emits them).
// Rows of >= 16 KiB that are 64-byte aligned are written in ONE streaming pass: zero lines with
// non-temporal stores, the few lines that hold counts built in a register buffer -- no read for
// ownership, nothing is read back; smaller rows are cleared and patched in cache.
void expand_one(const uint32_t* keys, const uint32_t* counts, int n, size_t fourk, int32_t* row)
{
    const bool aligned = (reinterpret_cast<uintptr_t>(row) & 63) == 0;
    if (fourk < 4096 || !aligned) {
        memset(row, 0, fourk * 4);
        for (int i = 0; i < n; i++) row[keys[i]] = (int32_t)counts[i];
        return;
    }
    size_t line = 0;                        // next 16-word line to write
    const size_t nlines = fourk / 16;
    int i = 0;
    while (i < n) {
        const size_t l = keys[i] >> 4;
        for (; line < l; line++) stream_zero_line(row + line * 16);
        alignas(64) int32_t buf[16] = {0};
        while (i < n && (keys[i] >> 4) == l) { buf[keys[i] & 15] = (int32_t)counts[i]; i++; }
#if defined(__SSE2__)
        for (int q = 0; q < 4; q++)
            _mm_stream_si128(reinterpret_cast<__m128i*>(row + l * 16) + q, _mm_load_si128(reinterpret_cast<const __m128i*>(buf) + q));
#else
        memcpy(row + l * 16, buf, 64);
#endif
        line = l + 1;
    }
    for (; line < nlines; line++) stream_zero_line(row + line * 16);
}

// rows [r0, r1) of the host share (indices relative to its first row): pairs at poff[], counts in prc[]
void expand_rows(const uint32_t* h_pk, const uint32_t* h_pc, const int64_t* poff, const int32_t* h_prc, int64_t r0, int64_t r1,
                 size_t fourk, int32_t* rows_out /* row 0 of the host share */, int nt)
{
    std::call_once(g_pool_once, [] { g_pool = new HostPool(); });
    const int64_t n = r1 - r0;
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n / 4));
    const std::function<void(int)> fn = [&](int p) {
        const int64_t a = r0 + n * p / parts, b = r0 + n * (p + 1) / parts;
        for (int64_t i = a; i < b; i++)
            expand_one(h_pk + poff[i], h_pc + poff[i], h_prc[i], fourk, rows_out + (size_t)i * fourk);
#if defined(__SSE2__)
        _mm_sfence();
#endif
    };
    g_pool->run(parts, fn);
}

int check_common(int fmt, int k, int kmax, int mode)
{
    if (fmt != CFRK_FMT_CODES && fmt != CFRK_FMT_ASCII) return fail(CFRK_EINVAL, "fmt must be CFRK_FMT_CODES or CFRK_FMT_ASCII");
    if (k < 1 || k > kmax) return fail(CFRK_EINVAL, "k out of range for this entry point");
    if (mode != CFRK_MODE_COMPAT && mode != CFRK_MODE_EXACT) return fail(CFRK_EINVAL, "mode must be CFRK_MODE_COMPAT or CFRK_MODE_EXACT");
    return CFRK_OK;
}

}  // namespace

namespace cfrk {
void set_last_error(const std::string& msg) { t_err = msg; }

// Pinned host memory for the rows kmer_main hands out (rd->Freq).  The reference pins a fresh
// buffer per call (cudaMallocHost, src/kmer_main.cu:115: ~0.35 ms per MiB, i.e. 1 s for the 2.86 GB
// of one k = 4..8 sweep over a chunk) and never frees it.  Here the buffers come from a process-wide
// arena: cfrk_free_host() puts a buffer back and the next call of that size class takes it over, so
// a caller that releases its rows pays the pinning once.  A caller that never frees (the reference
// driver) gets exactly the reference's behaviour.
namespace {
std::mutex g_arena_mu;
std::map<void*, size_t> g_arena_live;            // handed out: pointer -> capacity
std::multimap<size_t, void*> g_arena_free;       // cached: capacity -> pointer
constexpr size_t kArenaGrain = (size_t)2 << 20;
}

void* pinned_alloc(size_t bytes)
{
    const size_t cap = (std::max<size_t>(bytes, 1) + kArenaGrain - 1) / kArenaGrain * kArenaGrain;
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        auto it = g_arena_free.lower_bound(cap);
        if (it != g_arena_free.end() && it->first <= cap + cap / 4) {   // close fit only: no 2.6 GB block for a 8 MB row set
            void* p = it->second;
            g_arena_live[p] = it->first;
            g_arena_free.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(g_arena_mu);
    g_arena_live[p] = cap;
    return p;
}

void pinned_free(void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_arena_mu);
    auto it = g_arena_live.find(p);
    if (it == g_arena_live.end()) { cudaFreeHost(p); cudaGetLastError(); return; }   // not ours: plain pinned memory
    g_arena_free.emplace(it->second, p);
    g_arena_live.erase(it);
}

void pinned_release_cached()
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (auto& kv : g_arena_free) cudaFreeHost(kv.second);
    g_arena_free.clear();
    cudaGetLastError();
}
}  // namespace cfrk

extern "C" {

const char* cfrk_version(void) { return "cfrk_b200 0.1 (sm_100a)"; }
const char* cfrk_last_error(void) { return t_err.c_str(); }
uint64_t cfrk_launch_count(void) { return cfrk::launch_count(); }

int cfrk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cfrk_dense_reads_per_tile(int k) { return cfrk::dense_reads_per_tile(k); }

void cfrk_free_host(void* p) { cfrk::pinned_free(p); }

int cfrk_release(void)
{
    std::vector<HostCtx*> victims;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        victims.swap(g_idle_ctx);
        if (t_holder.c) { victims.push_back(t_holder.c); t_holder.c = nullptr; }
    }
    int dev = -1;
    const bool have_dev = cudaGetDevice(&dev) == cudaSuccess;
    for (HostCtx* c : victims) { c->release(); delete c; }
    cfrk::release_stream_scratch();
    cfrk::pinned_release_cached();
    if (have_dev) cudaSetDevice(dev);
    cudaGetLastError();
    return CFRK_OK;
}

int cfrk_count_dense_device(const void* d_bases, int fmt, const int64_t* d_start, const int32_t* d_length,
                            int64_t nN, int64_t nS, int64_t read_begin, int64_t read_end, int k, int mode,
                            int64_t chunk_size, int64_t first_read_index, int32_t* d_freq, void* stream)
{
    int rc = check_common(fmt, k, CFRK_DENSE_MAX_K, mode);
    if (rc) return rc;
    if (nS < 0 || nN < 0 || read_begin < 0 || read_end > nS || read_begin > read_end)
        return fail(CFRK_EINVAL, "bad read range");
    if (chunk_size < 0 || first_read_index < 0) return fail(CFRK_EINVAL, "negative chunk_size / first_read_index");
    if (read_begin == read_end) return CFRK_OK;
    if (!d_bases || !d_start || !d_length || !d_freq) return fail(CFRK_EINVAL, "null device pointer");
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (reinterpret_cast<uintptr_t>(d_freq) & 15))
        return fail(CFRK_EINVAL, "d_bases and d_freq must be 16-byte aligned");
    cudaError_t e = cfrk::launch_dense(d_bases, fmt, d_start, d_length, nN, nS, read_begin, read_end, k, mode,
                                       chunk_size, first_read_index, d_freq, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "dense_count_kernel launch");
    return CFRK_OK;
}

int cfrk_count_dense_packed_device(const uint32_t* d_codes, const uint16_t* d_valid, const int64_t* d_start,
                                   const int32_t* d_length, int64_t nN, int64_t nS, int64_t read_begin,
                                   int64_t read_end, int k, int mode, int64_t chunk_size, int64_t first_read_index,
                                   int32_t* d_freq, void* stream)
{
    int rc = check_common(CFRK_FMT_CODES, k, CFRK_DENSE_MAX_K, mode);
    if (rc) return rc;
    if (nS < 0 || nN < 0 || read_begin < 0 || read_end > nS || read_begin > read_end)
        return fail(CFRK_EINVAL, "bad read range");
    if (chunk_size < 0 || first_read_index < 0) return fail(CFRK_EINVAL, "negative chunk_size / first_read_index");
    if (read_begin == read_end) return CFRK_OK;
    if (!d_codes || !d_valid || !d_start || !d_length || !d_freq) return fail(CFRK_EINVAL, "null device pointer");
    if (reinterpret_cast<uintptr_t>(d_freq) & 15) return fail(CFRK_EINVAL, "d_freq must be 16-byte aligned");
    cudaError_t e = cfrk::launch_dense(d_codes, cfrk::FMT_PACKED, d_start, d_length, nN, nS, read_begin, read_end, k,
                                       mode, chunk_size, first_read_index, d_freq, static_cast<cudaStream_t>(stream),
                                       d_valid);
    if (e != cudaSuccess) return fail_cuda(e, "dense kernel launch (packed reads)");
    return CFRK_OK;
}

int cfrk_count_dense_host(const void* bases, int fmt, const int64_t* start, const int32_t* length,
                          int64_t nN, int64_t nS, int k, int mode, int device, int32_t* freq_out)
{
    int rc = check_common(fmt, k, CFRK_DENSE_MAX_K, mode);
    if (rc) return rc;
    if (nS < 0 || nN < 0) return fail(CFRK_EINVAL, "negative size");
    if (nS == 0) return CFRK_OK;
    if (!bases || !start || !length || !freq_out) return fail(CFRK_EINVAL, "null pointer");
    if (cfrk_device_count() <= device || device < 0)
        return fail(CFRK_ECUDA, "no such CUDA device (this library has no CPU fallback)");
    const auto wall0 = std::chrono::steady_clock::now();
    HostCtx* cp = nullptr;
    rc = ensure_ctx(device, &cp);
    if (rc) return rc;
    HostCtx& c = *cp;

    const size_t fourk = (size_t)1 << (2 * k);
    const size_t row_bytes = fourk * 4;
    const int rpt = cfrk::dense_reads_per_tile(k);

    // ---- split: rows [0, nD) leave as dense rows through the DMA engine, rows [nD, nS) as (bin, count) pairs
    // that host threads expand (see above).  Not for tiny rows / batches (nothing to win) and
    // not when a compat batch holds an empty read (its walk over the following reads stays on the GPU).
    const int nt = host_threads();
    int64_t nD = nS;
    if (nt > 0 && row_bytes >= 4096 && (size_t)nS * row_bytes >= ((size_t)32 << 20)) {
        bool ok = true;
        if (mode == CFRK_MODE_COMPAT)
            for (int64_t i = 0; i < nS && ok; i++) ok = length[i] != 0;
        if (ok) {
            double f = c.dma_frac[k];
            if (f <= 0.0) f = c.dma_frac[0] > 0.0 ? c.dma_frac[0] : 0.40;   // [0]: the share the last call of any k settled on
            nD = (int64_t)((double)nS * f);
            nD = std::max<int64_t>(0, std::min<int64_t>(nS, nD / rpt * rpt));
            if (nS - nD < 64) nD = nS;
        }
    }
    const int64_t nH = nS - nD;                     // rows expanded on the host
    // pairs of host row i (relative to nD) at poff[i]: capacity = min(4^k, visited windows + 1)
    std::vector<int64_t> poff;
    int64_t pairs_total = 0;
    if (nH > 0) {
        poff.resize((size_t)nH + 1);
        for (int64_t i = 0; i < nH; i++) {
            poff[(size_t)i] = pairs_total;
            pairs_total += std::min<int64_t>((int64_t)fourk, visited_windows(length[nD + i], k, mode) + 1);
        }
        poff[(size_t)nH] = pairs_total;
    }

    int64_t slice = (int64_t)std::max<size_t>(1, kRingSlotBytes / row_bytes);
    slice = std::max<int64_t>(rpt, slice / rpt * rpt);
    const int64_t dslice = std::min(slice, std::max<int64_t>(rpt, (nD + rpt - 1) / rpt * rpt));
    const int64_t hslice = std::min(slice, std::max<int64_t>(rpt, (nH + rpt - 1) / rpt * rpt));
    const int64_t nslices = nD > 0 ? (nD + dslice - 1) / dslice : 0;
    const int64_t nhslices = nH > 0 ? (nH + hslice - 1) / hslice : 0;

    if ((rc = grow(c.d_bases, c.cap_bases, (size_t)nN + CFRK_PAD))) return rc;
    {
        size_t need = (size_t)nS;
        if (need > c.cap_reads) {
            cudaFree(c.d_start); cudaFree(c.d_length); c.d_start = nullptr; c.d_length = nullptr; c.cap_reads = 0;
            size_t want = need + need / 4 + 64;
            if (cudaMalloc(reinterpret_cast<void**>(&c.d_start), want * 8) != cudaSuccess ||
                cudaMalloc(reinterpret_cast<void**>(&c.d_length), want * 4) != cudaSuccess) {
                cudaGetLastError();
                return fail(CFRK_ENOMEM, "cudaMalloc(start/length)");
            }
            c.cap_reads = want;
        }
    }
    if (nD > 0) {
        size_t need = (size_t)std::min<int64_t>(dslice, nD) * row_bytes;
        if (need > c.cap_ring) {
            cudaFree(c.d_ring[0]); cudaFree(c.d_ring[1]); c.d_ring[0] = c.d_ring[1] = nullptr; c.cap_ring = 0;
            const int slots = nslices > 1 ? 2 : 1;
            for (int i = 0; i < slots; i++)
                if (cudaMalloc(reinterpret_cast<void**>(&c.d_ring[i]), need) != cudaSuccess) {
                    cudaGetLastError();
                    return fail(CFRK_ENOMEM, "cudaMalloc(row ring)");
                }
            c.cap_ring = need;
        } else if (nslices > 1 && !c.d_ring[1]) {
            if (cudaMalloc(reinterpret_cast<void**>(&c.d_ring[1]), c.cap_ring) != cudaSuccess) {
                cudaGetLastError();
                return fail(CFRK_ENOMEM, "cudaMalloc(row ring)");
            }
        }
    }
    if (nH > 0) {
        if ((rc = grow(c.d_hslot, c.cap_hslot, (size_t)std::min(hslice, nH) * row_bytes))) return rc;
        const size_t np = (size_t)std::max<int64_t>(pairs_total, 1);
        if (np * 4 > c.cap_pairs) {
            cudaFree(c.d_pk); cudaFree(c.d_pc); c.d_pk = c.d_pc = nullptr; c.cap_pairs = 0;
            const size_t want = np * 4 + np + 256;
            if (cudaMalloc(reinterpret_cast<void**>(&c.d_pk), want) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&c.d_pc), want) != cudaSuccess) {
                cudaGetLastError();
                return fail(CFRK_ENOMEM, "cudaMalloc(pairs)");
            }
            c.cap_pairs = want;
        }
        if ((size_t)nH + 1 > c.cap_prows) {
            cudaFree(c.d_poff); cudaFree(c.d_prc); c.d_poff = nullptr; c.d_prc = nullptr; c.cap_prows = 0;
            const size_t want = (size_t)nH + (size_t)nH / 4 + 64;
            if (cudaMalloc(reinterpret_cast<void**>(&c.d_poff), want * 8) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&c.d_prc), want * 4) != cudaSuccess) {
                cudaGetLastError();
                return fail(CFRK_ENOMEM, "cudaMalloc(pair offsets)");
            }
            c.cap_prows = want;
        }
        if (np * 4 > c.cap_hpairs) {
            if (c.h_pk) cudaFreeHost(c.h_pk);
            if (c.h_pc) cudaFreeHost(c.h_pc);
            c.h_pk = c.h_pc = nullptr; c.cap_hpairs = 0;
            const size_t want = np * 4 + np + 256;
            if (cudaMallocHost(reinterpret_cast<void**>(&c.h_pk), want) != cudaSuccess || cudaMallocHost(reinterpret_cast<void**>(&c.h_pc), want) != cudaSuccess) {
                cudaGetLastError();
                return fail(CFRK_ENOMEM, "cudaMallocHost(pairs)");
            }
            c.cap_hpairs = want;
        }
        if ((size_t)nH * 4 > c.cap_hprc) {
            if (c.h_prc) cudaFreeHost(c.h_prc);
            c.h_prc = nullptr; c.cap_hprc = 0;
            const size_t want = (size_t)nH * 5 + 256;
            if (cudaMallocHost(reinterpret_cast<void**>(&c.h_prc), want) != cudaSuccess) { cudaGetLastError(); return fail(CFRK_ENOMEM, "cudaMallocHost(pair counts)"); }
            c.cap_hprc = want;
        }
        while ((int64_t)c.ready.size() < nhslices) {
            cudaEvent_t e1 = nullptr, e2 = nullptr;
            CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            c.packed.push_back(e1); c.ready.push_back(e2);
        }
    }

    CU(cudaMemcpyAsync(c.d_bases, bases, (size_t)nN, cudaMemcpyHostToDevice, c.compute));
    // the kernel loads whole 16-byte blocks: define the tail
    CU(cudaMemsetAsync(static_cast<char*>(c.d_bases) + nN, 0xFF, CFRK_PAD, c.compute));
    CU(cudaMemcpyAsync(c.d_start, start, (size_t)nS * 8, cudaMemcpyHostToDevice, c.compute));
    CU(cudaMemcpyAsync(c.d_length, length, (size_t)nS * 4, cudaMemcpyHostToDevice, c.compute));
    if (nH > 0) CU(cudaMemcpyAsync(c.d_poff, poff.data(), ((size_t)nH + 1) * 8, cudaMemcpyHostToDevice, c.compute));

    // The host share runs on its own stream (dense rows into a private slot -> pairs -> small D2H copies), so it
    // is not held back by the DMA slices, whose kernels wait for their ring slot to drain at PCIe pace; every
    // row -- compat spill included -- is counted by the same kernels.
    CU(cudaEventRecord(c.t0, c.compute));
    if (nhslices > 0) {
        CU(cudaEventRecord(c.packed[0], c.compute));            // inputs are in HBM
        CU(cudaStreamWaitEvent(c.aux, c.packed[0], 0));
    }
    for (int64_t s = 0; s < nhslices; s++) {
        const int64_t h0 = s * hslice, h1 = std::min(nH, h0 + hslice);        // relative to nD
        cudaError_t e = cfrk::launch_dense(c.d_bases, fmt, c.d_start, c.d_length, nN, nS, nD + h0, nD + h1, k, mode,
                                           0, 0, c.d_hslot, c.aux);
        if (e != cudaSuccess) return fail_cuda(e, "dense_count_kernel launch");
        e = cfrk::launch_rows_to_pairs(c.d_hslot, h1 - h0, (int)fourk, c.d_poff + h0, c.d_pk, c.d_pc, c.d_prc + h0, c.aux);
        if (e != cudaSuccess) return fail_cuda(e, "rows_to_pairs_kernel launch");
        const size_t p0 = (size_t)poff[(size_t)h0], p1 = (size_t)poff[(size_t)h1];
        // pinned host memory is mapped into the device's address space (unified addressing): SM stores, not the
        // copy engine, which is busy with the dense rows of the DMA share for milliseconds at a time
        e = cfrk::launch_pairs_to_host(reinterpret_cast<const uint32_t*>(c.d_prc + h0), reinterpret_cast<uint32_t*>(c.h_prc + h0), h1 - h0,
                                       c.d_pk + p0, c.h_pk + p0, (int64_t)(p1 - p0), c.d_pc + p0, c.h_pc + p0, (int64_t)(p1 - p0), c.aux);
        if (e != cudaSuccess) return fail_cuda(e, "pairs_to_host_kernel launch");
        CU(cudaEventRecord(c.ready[(size_t)s], c.aux));
    }
    for (int64_t s = 0; s < nslices; s++) {
        {
            const int slot = (int)(s & 1);
            const int64_t r0 = s * dslice, r1 = std::min(nD, r0 + dslice);
            if (s >= 2) CU(cudaStreamWaitEvent(c.compute, c.drained[slot], 0));
            cudaError_t e = cfrk::launch_dense(c.d_bases, fmt, c.d_start, c.d_length, nN, nS, r0, r1, k, mode,
                                               0, 0, c.d_ring[slot], c.compute);
            if (e != cudaSuccess) return fail_cuda(e, "dense_count_kernel launch");
            CU(cudaEventRecord(c.done[slot], c.compute));
            CU(cudaStreamWaitEvent(c.copy, c.done[slot], 0));
            CU(cudaMemcpyAsync(freq_out + (size_t)r0 * fourk, c.d_ring[slot], (size_t)(r1 - r0) * row_bytes,
                               cudaMemcpyDeviceToHost, c.copy));
            CU(cudaEventRecord(c.drained[slot], c.copy));
        }
    }
    CU(cudaEventRecord(c.t1, c.copy));
    double host_ms = 0.0;
    for (int64_t s = 0; s < nhslices; s++) {
        CU(cudaEventSynchronize(c.ready[(size_t)s]));
        const int64_t h0 = s * hslice, h1 = std::min(nH, h0 + hslice);
        const auto t_a = std::chrono::steady_clock::now();
        expand_rows(c.h_pk, c.h_pc, poff.data(), c.h_prc, h0, h1, fourk, freq_out + (size_t)nD * fourk, nt);
        host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count();
    }
    CU(cudaStreamSynchronize(c.copy));
    CU(cudaStreamSynchronize(c.aux));
    CU(cudaStreamSynchronize(c.compute));
    static const bool trace = getenv("CFRK_TRACE") != nullptr;
    if (trace) {
        float ms = 0.f;
        if (nD > 0) cudaEventElapsedTime(&ms, c.t0, c.t1);
        fprintf(stderr, "[cfrk host op] k=%d nS=%lld dma rows=%lld (%.2f ms) host rows=%lld (%.2f ms, %d threads) call %.2f ms\n", k,
                (long long)nS, (long long)nD, ms, (long long)nH, host_ms, nt,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
        cudaGetLastError();
    }
    if (nH > 0 && nD > 0) {
        // steer the split towards equal finishing times of the DMA side and the host side
        float dma_ms = 0.f;
        if (cudaEventElapsedTime(&dma_ms, c.t0, c.t1) == cudaSuccess && dma_ms > 0.f && host_ms > 0.0) {
            const double rate_d = (double)nD / dma_ms, rate_h = (double)nH / host_ms;
            const double target = rate_d / (rate_d + rate_h);
            c.dma_frac[k] = std::min(0.95, std::max(0.05, 0.3 * ((double)nD / (double)nS) + 0.7 * target));
            c.dma_frac[0] = c.dma_frac[k];
        }
        cudaGetLastError();
    }
    return CFRK_OK;
}

void cfrk_set_host_threads(int n)
{
    g_host_threads.store(n < 0 ? -2 : std::min(n, 256));    // negative: back to the default
}

int cfrk_encode_2bit_device(const void* d_bases, int fmt, int64_t n, uint32_t* d_codes, uint16_t* d_valid,
                            void* stream)
{
    if (fmt != CFRK_FMT_CODES && fmt != CFRK_FMT_ASCII) return fail(CFRK_EINVAL, "bad fmt");
    if (n < 0) return fail(CFRK_EINVAL, "negative size");
    if (n == 0) return CFRK_OK;
    if (!d_bases || !d_codes || !d_valid) return fail(CFRK_EINVAL, "null device pointer");
    if (reinterpret_cast<uintptr_t>(d_bases) & 15) return fail(CFRK_EINVAL, "d_bases must be 16-byte aligned");
    cudaError_t e = cfrk::launch_encode_2bit(d_bases, fmt, n, d_codes, d_valid, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "encode_2bit_kernel launch");
    return CFRK_OK;
}

int cfrk_count_sparse_device(const void* d_bases, int fmt, const int64_t* d_start, const int32_t* d_length,
                             int64_t nN, int64_t nS, int k, int key_bytes, int64_t* d_row_begin, int32_t* d_row_count,
                             void* d_keys, uint32_t* d_counts, int64_t capacity, int64_t* total_windows, void* stream)
{
    int rc = check_common(fmt, k, CFRK_SPARSE_MAX_K, CFRK_MODE_EXACT);
    if (rc) return rc;
    if (key_bytes != 4 && key_bytes != 8) return fail(CFRK_EINVAL, "key_bytes must be 4 or 8");
    if (key_bytes == 4 && k > 16) return fail(CFRK_EINVAL, "uint32 keys hold k <= 16");
    if (nS < 0 || nN < 0 || capacity < 0) return fail(CFRK_EINVAL, "negative size");
    if (!d_row_begin) return fail(CFRK_EINVAL, "null device pointer");
    if (total_windows) *total_windows = 0;
    if (nS == 0) return CFRK_OK;
    if (!d_bases || !d_start || !d_length || !d_row_count || !d_keys || !d_counts)
        return fail(CFRK_EINVAL, "null device pointer");
    if (reinterpret_cast<uintptr_t>(d_bases) & 15) return fail(CFRK_EINVAL, "d_bases must be 16-byte aligned");
    int64_t total = -1;
    cudaError_t e = cfrk::launch_sparse(d_bases, fmt, d_start, d_length, nS, k, d_row_begin, d_row_count, d_keys,
                                        key_bytes, d_counts, capacity, &total, static_cast<cudaStream_t>(stream));
    if (total_windows) *total_windows = total < 0 ? 0 : total;
    // the capacity case is recognised by the totals, not by the error code (cudaErrorInvalidValue can
    // also come from a launch configuration or an attribute call)
    if (total > capacity) { cudaGetLastError(); return fail(CFRK_EINVAL, "capacity smaller than the number of windows"); }
    if (e != cudaSuccess) return fail_cuda(e, "sparse path");
    return CFRK_OK;
}

int cfrk_count_sparse_packed_device(const uint32_t* d_codes, const uint16_t* d_valid, const int64_t* d_start,
                                    const int32_t* d_length, int64_t nN, int64_t nS, int k, int key_bytes,
                                    int64_t* d_row_begin, int32_t* d_row_count, void* d_keys, uint32_t* d_counts,
                                    int64_t capacity, int64_t* total_windows, void* stream)
{
    int rc = check_common(CFRK_FMT_CODES, k, CFRK_SPARSE_MAX_K, CFRK_MODE_EXACT);
    if (rc) return rc;
    if (key_bytes != 4 && key_bytes != 8) return fail(CFRK_EINVAL, "key_bytes must be 4 or 8");
    if (key_bytes == 4 && k > 16) return fail(CFRK_EINVAL, "uint32 keys hold k <= 16");
    if (nS < 0 || nN < 0 || capacity < 0) return fail(CFRK_EINVAL, "negative size");
    if (!d_row_begin) return fail(CFRK_EINVAL, "null device pointer");
    if (total_windows) *total_windows = 0;
    if (nS == 0) return CFRK_OK;
    if (!d_codes || !d_valid || !d_start || !d_length || !d_row_count || !d_keys || !d_counts)
        return fail(CFRK_EINVAL, "null device pointer");
    int64_t total = -1;
    cudaError_t e = cfrk::launch_sparse(d_codes, cfrk::FMT_PACKED, d_start, d_length, nS, k, d_row_begin, d_row_count, d_keys,
                                        key_bytes, d_counts, capacity, &total, static_cast<cudaStream_t>(stream), d_valid);
    if (total_windows) *total_windows = total < 0 ? 0 : total;
    if (total > capacity) { cudaGetLastError(); return fail(CFRK_EINVAL, "capacity smaller than the number of windows"); }
    if (e != cudaSuccess) return fail_cuda(e, "sparse path (packed reads)");
    return CFRK_OK;
}

int cfrk_scan_fasta_device(const void* d_bytes, int64_t n, int is_final, int64_t* d_header, int64_t* d_start,
                           int32_t* d_length, int64_t capacity, int64_t* n_headers, void* stream)
{
    if (n_headers) *n_headers = 0;
    if (n < 0 || capacity < 0) return fail(CFRK_EINVAL, "negative size");
    if (n == 0) return CFRK_OK;
    if (!d_bytes || !d_header || !d_start || !d_length) return fail(CFRK_EINVAL, "null device pointer");
    if (reinterpret_cast<uintptr_t>(d_bytes) & 15) return fail(CFRK_EINVAL, "d_bytes must be 16-byte aligned");
    int64_t out[2];
    cudaError_t e = cfrk::launch_fasta_scan(static_cast<const uint8_t*>(d_bytes), n, is_final, d_header, d_start,
                                            d_length, capacity, out, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "fasta scan");
    if (n_headers) *n_headers = out[0];
    switch (out[1]) {
    case 0: return CFRK_OK;
    case 1: return fail(CFRK_EFORMAT, "'>' inside a sequence line (grep -c over-counts nS in the reference, src/fastaIO.h:16)");
    case 2: return fail(CFRK_EFORMAT, "sequence text before the first '>' header (undefined in the reference, src/fastaIO.h:49-52)");
    case 3: return fail(CFRK_EFORMAT, "record longer than 2^31-1 bytes (length is int in the reference, src/tipos.h:26)");
    default: return fail(CFRK_EINVAL, "capacity smaller than the number of headers");
    }
}

int cfrk_global_hist_device(const void* d_bases, int fmt, const int64_t* d_start, const int32_t* d_length,
                            int64_t nN, int64_t nS, int k, uint32_t* d_hist, void* stream)
{
    int rc = check_common(fmt, k, CFRK_HIST_MAX_K, CFRK_MODE_EXACT);
    if (rc) return rc;
    if (nS < 0 || nN < 0) return fail(CFRK_EINVAL, "negative size");
    if (nS == 0) return CFRK_OK;
    if (!d_bases || !d_start || !d_length || !d_hist) return fail(CFRK_EINVAL, "null device pointer");
    if (reinterpret_cast<uintptr_t>(d_bases) & 15) return fail(CFRK_EINVAL, "d_bases must be 16-byte aligned");
    cudaError_t e = cfrk::launch_global_hist(d_bases, fmt, d_start, d_length, nN, nS, k, d_hist,
                                             static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "global_hist_kernel launch");
    return CFRK_OK;
}

}  // extern "C"
