// api.cu -- the C ABI of libcfrk_b200.so (include/cfrk_b200.h) over the kernels.
//
// cfrk_count_dense_host is the operator that replaces the reference's kmer_main
// (src/kmer_main.cu:20-128): same inputs and output, but
//   * device buffers are cached per host thread instead of 5 cudaMalloc + 5 cudaFree per call
//     (src/kmer_main.cu:59-63,120-124),
//   * rows leave the device through a two-slot ring so that the device->host copy of slice i
//     overlaps the kernel of slice i+1 (the reference does one synchronous cudaMemcpy of the
//     whole Freq, src/kmer_main.cu:116),
//   * part of the rows is written by host threads from index lists (HostExpand below), so that the
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
    // host-expanded part of a call: index lists (device + pinned mirror), their offsets, timing events
    uint32_t* d_idx = nullptr; size_t cap_idx = 0;
    int64_t* d_ibeg = nullptr; size_t cap_ibeg = 0;
    uint32_t* h_idx = nullptr; size_t cap_hidx = 0;
    cudaEvent_t idx_ready = nullptr, t0 = nullptr, t1 = nullptr;
    double dma_frac[CFRK_DENSE_MAX_K + 1] = {0};   // share of the rows that goes through the DMA engine, per k (adaptive)

    void release()
    {
        if (device < 0) return;
        cudaSetDevice(device);
        cudaFree(d_bases); cudaFree(d_start); cudaFree(d_length);
        cudaFree(d_ring[0]); cudaFree(d_ring[1]);
        cudaFree(d_idx); cudaFree(d_ibeg);
        if (h_idx) cudaFreeHost(h_idx);
        if (idx_ready) cudaEventDestroy(idx_ready);
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
        CU(cudaEventCreateWithFlags(&c->idx_ready, cudaEventDisableTiming));
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
// the GPU's DMA engine as before, for the other share only the k-mer index of every visited window
// crosses the bus (dense_index.cu: 4 bytes per window) and `nt` host threads write those rows -- zeros
// and counts in one pass of streaming stores, no read-for-ownership.  The split adapts per k to
// whichever side finishes first.  The k-mers are still computed on the GPU only.
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

// one row from its index list: bins[v]++ for every valid entry; returns the number of invalid entries
// (compat: they are owed to the previous row's last bin).  Rows of >= 16 KiB are written in one
// streaming pass (sorted indices, zero lines with non-temporal stores, the few lines that hold counts
// built in a register buffer); smaller rows are cleared and counted in cache.
int expand_one(const uint32_t* idx, int64_t n, size_t fourk, int32_t* row, uint32_t* tmp)
{
    int invalid = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(row) & 63) == 0;
    if (fourk < 4096 || !aligned) {
        memset(row, 0, fourk * 4);
        for (int64_t t = 0; t < n; t++) {
            const uint32_t v = idx[t];
            if (v < fourk) row[v]++; else invalid++;
        }
        return invalid;
    }
    int64_t m = 0;
    for (int64_t t = 0; t < n; t++) {
        const uint32_t v = idx[t];
        if (v < fourk) tmp[m++] = v; else invalid++;
    }
    std::sort(tmp, tmp + m);
    size_t line = 0;                        // next 16-word line to write
    const size_t nlines = fourk / 16;
    int64_t i = 0;
    while (i < m) {
        const size_t l = tmp[i] >> 4;
        for (; line < l; line++) stream_zero_line(row + line * 16);
        alignas(64) int32_t buf[16] = {0};
        while (i < m && (tmp[i] >> 4) == l) buf[tmp[i++] & 15]++;
#if defined(__SSE2__)
        for (int q = 0; q < 4; q++)
            _mm_stream_si128(reinterpret_cast<__m128i*>(row + l * 16) + q, _mm_load_si128(reinterpret_cast<const __m128i*>(buf) + q));
#else
        memcpy(row + l * 16, buf, 64);
#endif
        line = l + 1;
    }
    for (; line < nlines; line++) stream_zero_line(row + line * 16);
    return invalid;
}

// rows [nD, nS) of freq_out from the index lists (h_idx, ibeg relative to read nD)
void expand_rows(const uint32_t* h_idx, const int64_t* ibeg, int64_t nD, int64_t nS, size_t fourk, int mode,
                 int32_t* freq_out, int nt)
{
    std::call_once(g_pool_once, [] { g_pool = new HostPool(); });
    const int64_t nH = nS - nD;
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(nt, nH / 4));
    const std::function<void(int)> fn = [&](int p) {
        const int64_t a = nH * p / parts, b = nH * (p + 1) / parts;
        std::vector<uint32_t> tmp((size_t)cfrk::kRefBlockThreads + 64);
        for (int64_t i = a; i < b; i++) {
            const int64_t n = ibeg[i + 1] - ibeg[i];
            if ((size_t)n > tmp.size()) tmp.resize((size_t)n);
            expand_one(h_idx + ibeg[i], n, fourk, freq_out + (size_t)(nD + i) * fourk, tmp.data());
        }
#if defined(__SSE2__)
        _mm_sfence();
#endif
        if (mode == CFRK_MODE_COMPAT) {
            // spill: the invalid windows of read i+1 land in the last bin of row i (src/kmer_kernel.cu:84-87).
            // Row nD-1 is the GPU's (it scans read nD itself); the thread that owns row i looks at list i+1.
            for (int64_t i = a; i < b; i++) {
                if (i + 1 >= nH) break;
                int inv = 0;
                for (int64_t t = ibeg[i + 1]; t < ibeg[i + 2]; t++) inv += h_idx[t] >= fourk;
                if (inv) freq_out[(size_t)(nD + i) * fourk + fourk - 1] += inv;
            }
        }
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

    // ---- split: rows [0, nD) leave as dense rows through the DMA engine, rows [nD, nS) as index lists
    // that host threads expand (see HostExpand above).  Not for tiny rows / batches (nothing to win) and
    // not when a compat batch holds an empty read (its walk over the following reads stays on the GPU).
    const int nt = host_threads();
    int64_t nD = nS;
    std::vector<int64_t> ibeg;
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
    int64_t idx_total = 0;
    if (nH > 0) {
        ibeg.resize((size_t)nH + 1);
        for (int64_t i = 0; i < nH; i++) {
            ibeg[(size_t)i] = idx_total;
            idx_total += visited_windows(length[nD + i], k, mode);
        }
        ibeg[(size_t)nH] = idx_total;
    }

    int64_t slice = (int64_t)std::max<size_t>(1, kRingSlotBytes / row_bytes);
    slice = std::max<int64_t>(rpt, slice / rpt * rpt);
    if (slice > nD) slice = std::max<int64_t>(rpt, (nD + rpt - 1) / rpt * rpt);
    const int64_t nslices = nD > 0 ? (nD + slice - 1) / slice : 0;

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
        size_t need = (size_t)std::min<int64_t>(slice, nD) * row_bytes;
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
        if ((rc = grow(c.d_idx, c.cap_idx, (size_t)std::max<int64_t>(idx_total, 1) * 4))) return rc;
        if ((rc = grow(c.d_ibeg, c.cap_ibeg, ((size_t)nH + 1) * 8))) return rc;
        const size_t need = (size_t)std::max<int64_t>(idx_total, 1) * 4;
        if (need > c.cap_hidx) {
            if (c.h_idx) cudaFreeHost(c.h_idx);
            c.h_idx = nullptr; c.cap_hidx = 0;
            if (cudaMallocHost(reinterpret_cast<void**>(&c.h_idx), need + need / 4) != cudaSuccess) {
                cudaGetLastError();
                return fail(CFRK_ENOMEM, "cudaMallocHost(index lists)");
            }
            c.cap_hidx = need + need / 4;
        }
    }

    CU(cudaMemcpyAsync(c.d_bases, bases, (size_t)nN, cudaMemcpyHostToDevice, c.compute));
    // the kernel loads whole 16-byte blocks: define the tail
    CU(cudaMemsetAsync(static_cast<char*>(c.d_bases) + nN, 0xFF, CFRK_PAD, c.compute));
    CU(cudaMemcpyAsync(c.d_start, start, (size_t)nS * 8, cudaMemcpyHostToDevice, c.compute));
    CU(cudaMemcpyAsync(c.d_length, length, (size_t)nS * 4, cudaMemcpyHostToDevice, c.compute));

    if (nH > 0) {
        // index lists first: they are small, and the host threads can start while the dense rows move
        CU(cudaMemcpyAsync(c.d_ibeg, ibeg.data(), ((size_t)nH + 1) * 8, cudaMemcpyHostToDevice, c.compute));
        cudaError_t e = cfrk::launch_dense_index(c.d_bases, fmt, c.d_start, c.d_length, c.d_ibeg, nD, nS, k, mode, c.d_idx, c.compute);
        if (e != cudaSuccess) return fail_cuda(e, "dense_index_kernel launch");
        CU(cudaMemcpyAsync(c.h_idx, c.d_idx, (size_t)idx_total * 4, cudaMemcpyDeviceToHost, c.compute));
        CU(cudaEventRecord(c.idx_ready, c.compute));
    }
    CU(cudaEventRecord(c.t0, c.compute));
    for (int64_t s = 0; s < nslices; s++) {
        const int slot = (int)(s & 1);
        const int64_t r0 = s * slice, r1 = std::min(nD, r0 + slice);
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
    CU(cudaEventRecord(c.t1, c.copy));
    double host_ms = 0.0;
    if (nH > 0) {
        CU(cudaEventSynchronize(c.idx_ready));
        const auto h0 = std::chrono::steady_clock::now();
        expand_rows(c.h_idx, ibeg.data(), nD, nS, fourk, mode, freq_out, nt);
        host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
    }
    CU(cudaStreamSynchronize(c.copy));
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
