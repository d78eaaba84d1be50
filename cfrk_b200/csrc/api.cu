// api.cu -- the C ABI of libcfrk_b200.so (include/cfrk_b200.h) over the kernels.
//
// cfrk_count_dense_host is the operator that replaces the reference's kmer_main
// (src/kmer_main.cu:20-128): same inputs and output, but
//   * device buffers are cached per host thread instead of 5 cudaMalloc + 5 cudaFree per call
//     (src/kmer_main.cu:59-63,120-124),
//   * rows leave the device through a two-slot ring so that the device->host copy of slice i
//     overlaps the kernel of slice i+1 (the reference does one synchronous cudaMemcpy of the
//     whole Freq, src/kmer_main.cu:116),
//   * errors are returned, not printed.
#include "../../include/cfrk_b200.h"
#include "kernels.h"
#include "kmer_device.cuh"
#include "internal.h"

#include <algorithm>
#include <cstdio>
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

    void release()
    {
        if (device < 0) return;
        cudaSetDevice(device);
        cudaFree(d_bases); cudaFree(d_start); cudaFree(d_length);
        cudaFree(d_ring[0]); cudaFree(d_ring[1]);
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
    HostCtx* cp = nullptr;
    rc = ensure_ctx(device, &cp);
    if (rc) return rc;
    HostCtx& c = *cp;

    const size_t fourk = (size_t)1 << (2 * k);
    const size_t row_bytes = fourk * 4;
    const int rpt = cfrk::dense_reads_per_tile(k);
    int64_t slice = (int64_t)std::max<size_t>(1, kRingSlotBytes / row_bytes);
    slice = std::max<int64_t>(rpt, slice / rpt * rpt);
    if (slice > nS) slice = (nS + rpt - 1) / rpt * rpt;
    const int64_t nslices = (nS + slice - 1) / slice;

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
    {
        size_t need = (size_t)std::min<int64_t>(slice, nS) * row_bytes;
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

    CU(cudaMemcpyAsync(c.d_bases, bases, (size_t)nN, cudaMemcpyHostToDevice, c.compute));
    // the kernel loads whole 16-byte blocks: define the tail
    CU(cudaMemsetAsync(static_cast<char*>(c.d_bases) + nN, 0xFF, CFRK_PAD, c.compute));
    CU(cudaMemcpyAsync(c.d_start, start, (size_t)nS * 8, cudaMemcpyHostToDevice, c.compute));
    CU(cudaMemcpyAsync(c.d_length, length, (size_t)nS * 4, cudaMemcpyHostToDevice, c.compute));

    for (int64_t s = 0; s < nslices; s++) {
        const int slot = (int)(s & 1);
        const int64_t r0 = s * slice, r1 = std::min(nS, r0 + slice);
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
    CU(cudaStreamSynchronize(c.copy));
    CU(cudaStreamSynchronize(c.compute));
    return CFRK_OK;
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
