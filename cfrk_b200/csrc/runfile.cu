// runfile.cu -- cfrk_run_file(): FASTA file -> GPU count -> .cfrk text, the whole of the
// reference's main() (src/main.cu:232-305) re-designed around the GPU.
//
//   reader   FastaStreamer: a thread preads the file into a ring of PINNED buffers; the raw
//            bytes are what the GPU consumes (CFRK_FMT_ASCII).  The record text of the
//            reference parser (src/fastaIO.h:38-69: every non-header line, '\n' included,
//            minus the last byte) is a contiguous span of the file, so a record is described
//            by (start, length) into the raw bytes and nothing is copied or encoded on the
//            host -- newline-as-base, CRLF, missing final newline all fall out.
//            Replaces popen("grep -c") + getline + 3 malloc/record + strcat + per-base switch
//            + ProcessData + SelectChunk (src/fastaIO.h:12-148, src/main.cu:110-206).
//   scan     on the GPU (fasta_scan.cu): header positions, (start, length) per record; the host
//            gets the small table back and never looks at a base.  A '>' that does not start a
//            line, or text before the first header, is where the reference is undefined:
//            CFRK_EFORMAT.
//   count    dense_count_kernel over the buffer, rows through a two-slot device/pinned ring.
//   writer   nt threads format rows ("bin:count ", src/main.cu:53-55) into private buffers
//            that are written in order; "\n" before every row but the first, none at EOF.
//
// Default (compat) output = only reads [ (nS/chunkSize)*chunkSize, nS ), because the reference
// re-opens the output with "w" for the remainder chunk (src/main.cu:34,303-305): the file is
// scanned once for headers and only that tail is uploaded.  CFRK_RUN_ALL_ROWS streams every read.
#include "../../include/cfrk_b200.h"
#include "kernels.h"
#include "kmer_device.cuh"
#include "internal.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <memory>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct Err {
    int code = CFRK_OK;
    std::string msg;
};

// CFRK_TRACE=1: wall-clock trace of the file pipeline on stderr (the reference only has
// commented-out time(NULL) probes, src/main.cu:259-268,302-306)
struct Trace {
    bool on = getenv("CFRK_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    void mark(const char* what, size_t bytes = 0)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cfrk trace] %9.3f ms (+%8.3f) %s", std::chrono::duration<double, std::milli>(now - t0).count(),
                std::chrono::duration<double, std::milli>(now - last).count(), what);
        if (bytes) fprintf(stderr, " %zu bytes", bytes);
        fputc('\n', stderr);
        last = now;
    }
};
#define RF_CU(call)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            err.code = CFRK_ECUDA;                                                   \
            err.msg = std::string(#call) + ": " + cudaGetErrorString(e_);            \
            return false;                                                            \
        }                                                                            \
    } while (0)

// ------------------------------------------------------------------------------------------
// Pinned, multi-buffered sequential file reader.  Each buffer has `headroom` bytes in front of
// the region the reader fills, so the consumer can prepend the unfinished records of the
// previous buffer and hand the GPU one contiguous span.
class FastaStreamer {
public:
    static constexpr int NB = 3;
    struct Buf { char* base = nullptr; size_t n = 0; bool eof = false; bool filled = false; };

    bool open(const char* path, size_t chunk, size_t headroom, Err& err)
    {
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) { err.code = CFRK_EIO; err.msg = std::string("cannot open ") + path; return false; }
        struct stat st;
        if (fstat(fd_, &st) != 0) { err.code = CFRK_EIO; err.msg = "fstat failed"; return false; }
        size_ = (size_t)st.st_size;
        chunk_ = chunk; headroom_ = headroom;
        for (int i = 0; i < NB; i++) {
            if (cudaMallocHost(reinterpret_cast<void**>(&bufs_[i].base), headroom + chunk + CFRK_PAD) != cudaSuccess) {
                cudaGetLastError();
                err.code = CFRK_ENOMEM; err.msg = "cudaMallocHost(stream buffer)";
                return false;
            }
        }
        th_ = std::thread([this] { run(); });
        return true;
    }
    // i-th buffer of the file (blocks until read). The new bytes are at base+headroom.
    Buf* acquire(size_t i)
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return bufs_[i % NB].filled && seq_[i % NB] == i; });
        return &bufs_[i % NB];
    }
    void release(size_t i)
    {
        std::lock_guard<std::mutex> lk(mu_);
        bufs_[i % NB].filled = false;
        cv_.notify_all();
    }
    char* next_base(size_t i) { return bufs_[(i + 1) % NB].base; }
    size_t headroom() const { return headroom_; }
    size_t file_size() const { return size_; }
    int fd() const { return fd_; }
    ~FastaStreamer()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            cv_.notify_all();
        }
        if (th_.joinable()) th_.join();
        for (int i = 0; i < NB; i++) if (bufs_[i].base) cudaFreeHost(bufs_[i].base);
        if (fd_ >= 0) ::close(fd_);
    }

private:
    void run()
    {
        size_t off = 0;
        for (size_t i = 0;; i++) {
            Buf& b = bufs_[i % NB];
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !b.filled; });
                if (stop_) return;
            }
            size_t want = std::min(chunk_, size_ - off), got = 0;
            while (got < want) {
                ssize_t r = pread(fd_, b.base + headroom_ + got, want - got, (off_t)(off + got));
                if (r <= 0) break;
                got += (size_t)r;
            }
            off += got;
            {
                std::lock_guard<std::mutex> lk(mu_);
                b.n = got; b.eof = (off >= size_) || got < want; b.filled = true; seq_[i % NB] = i;
                cv_.notify_all();
            }
            if (b.eof) return;
        }
    }
    int fd_ = -1;
    size_t size_ = 0, chunk_ = 0, headroom_ = 0;
    Buf bufs_[NB];
    size_t seq_[NB] = {0, 0, 0};
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// ------------------------------------------------------------------------------------------
// Record table of a span, as fasta_scan.cu produces it on the device.
struct RecordIndex {
    std::vector<int64_t> start;   // records whose end is known: all of them in a final span, else all but the last
    std::vector<int32_t> length;
    std::vector<size_t> header;   // position of each header's '>'
};

// ------------------------------------------------------------------------------------------
// .cfrk text writer
// "bin:" labels as fixed 8-byte records (bins < 65536 -> at most "65535:"): one 64-bit store per token
struct BinLabels {
    std::vector<uint64_t> rec;
    std::vector<uint8_t> len;
    explicit BinLabels(size_t bins) : rec(bins), len(bins)
    {
        char tmp[16];
        for (size_t b = 0; b < bins; b++) {
            memset(tmp, 0, sizeof tmp);
            len[b] = (uint8_t)snprintf(tmp, sizeof tmp, "%zu:", b);
            memcpy(&rec[b], tmp, 8);
        }
    }
};

// output scratch that is NOT value-initialised (a std::vector would zero-fill the worst-case size
// of every slice: measured as most of the writer's time) and is kept between slices
struct RawBuf {
    char* data = nullptr;
    size_t size = 0, cap = 0;
    void reserve(size_t n)
    {
        if (n > cap) {
            free(data);
            data = static_cast<char*>(malloc(n));
            cap = data ? n : 0;
        }
        size = 0;
    }
    RawBuf() = default;
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { free(data); }
};

inline char* put_int(char* p, int32_t v)
{
    uint32_t u = (uint32_t)v;
    if (v < 0) { *p++ = '-'; u = (uint32_t)(-(int64_t)v); }  // cannot happen (counts), kept for "%d" fidelity
    if (u < 10) { *p++ = (char)('0' + u); return p; }
    if (u < 100) { *p++ = (char)('0' + u / 10); *p++ = (char)('0' + u % 10); return p; }
    char tmp[12];
    int l = 0;
    do { tmp[l++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (l) *p++ = tmp[--l];
    return p;
}

void format_rows(const int32_t* rows, size_t nrows, size_t bins, const BinLabels& lab, bool sparse,
                 bool first_row_of_file, RawBuf& out)
{
    // worst case per token: 8-byte label store + 11 digits + space
    out.reserve(nrows * (bins * 20 + 1) + 16);
    char* p = out.data;
    for (size_t r = 0; r < nrows; r++) {
        if (!(first_row_of_file && r == 0)) *p++ = '\n';
        const int32_t* row = rows + r * bins;
        for (size_t b = 0; b < bins; b++) {
            const int32_t v = row[b];
            if (sparse && v == 0) continue;
            memcpy(p, &lab.rec[b], 8);
            p += lab.len[b];
            p = put_int(p, v);
            *p++ = ' ';
        }
    }
    out.size = (size_t)(p - out.data);
}

class CfrkWriter {
public:
    bool open(const char* path, int k, int nt, bool sparse, Err& err)
    {
        fd_ = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);   // fopen(path, "w"), src/main.cu:34
        if (fd_ < 0) { err.code = CFRK_EIO; err.msg = std::string("cannot open output ") + path; return false; }
        seekable_ = lseek(fd_, 0, SEEK_CUR) != (off_t)-1;
        bins_ = k <= CFRK_CLI_DENSE_MAX_K ? (size_t)1 << (2 * k) : 0;
        if (bins_) labels_.reset(new BinLabels(bins_));
        nt_ = std::max(1, std::min(nt, 64));
        sparse_ = sparse;
        return true;
    }
    bool write_rows(const int32_t* rows, size_t nrows, Err& err)
    {
        if (!nrows) return true;
        const int nt = (int)std::min<size_t>((size_t)nt_, nrows);
        if (parts_.size() < (size_t)nt_) parts_ = std::vector<RawBuf>(nt_);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) {
            const size_t a = nrows * t / nt, b = nrows * (t + 1) / nt;
            const bool first = first_ && a == 0;
            th.emplace_back([=] { format_rows(rows + a * bins_, b - a, bins_, *labels_, sparse_, first, parts_[t]); });
        }
        for (auto& x : th) x.join();
        first_ = false;
        return flush_parts(nt, err);
    }
    // rows given as (key, count) pairs: the same "bin:count " tokens, non-zero bins only
    bool write_pairs(const int64_t* row_begin, const int32_t* row_count, const uint64_t* keys,
                     const uint32_t* counts, size_t nrows, Err& err)
    {
        if (!nrows) return true;
        const int nt = (int)std::min<size_t>((size_t)nt_, nrows);
        if (parts_.size() < (size_t)nt_) parts_ = std::vector<RawBuf>(nt_);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) {
            const size_t a = nrows * t / nt, b = nrows * (t + 1) / nt;
            const bool first = first_ && a == 0;
            th.emplace_back([=] {
                size_t npairs = 0;
                for (size_t r = a; r < b; r++) npairs += (size_t)row_count[r];
                RawBuf& out = parts_[t];
                out.reserve(npairs * 33 + (b - a) + 1);   // 20 + 1 + 10 + 1 per token, '\n' per row
                char* p = out.data;
                for (size_t r = a; r < b; r++) {
                    if (!(first && r == a)) *p++ = '\n';
                    const uint64_t* kk = keys + row_begin[r];
                    const uint32_t* cc = counts + row_begin[r];
                    for (int32_t i = 0; i < row_count[r]; i++) {
                        char tmp[24];
                        int l = 0;
                        uint64_t u = kk[i];
                        do { tmp[l++] = (char)('0' + u % 10); u /= 10; } while (u);
                        while (l) *p++ = tmp[--l];
                        *p++ = ':';
                        p = put_int(p, (int32_t)cc[i]);
                        *p++ = ' ';
                    }
                }
                out.size = (size_t)(p - out.data);
            });
        }
        for (auto& x : th) x.join();
        first_ = false;
        return flush_parts(nt, err);
    }
    ~CfrkWriter() { if (fd_ >= 0) ::close(fd_); }

private:
    // formatted parts -> file, in order.  Regular files: every part is written at its own offset
    // by its own thread (pwrite); pipes (the Swift stdout form): sequential write.
    bool flush_parts(int nparts, Err& err)
    {
        bool ok = true;
        if (seekable_) {
            std::vector<off_t> at(nparts);
            for (int i = 0; i < nparts; i++) { at[i] = off_; off_ += (off_t)parts_[i].size; }
            std::vector<std::thread> th;
            std::vector<char> good(nparts, 1);
            for (int i = 0; i < nparts; i++)
                th.emplace_back([&, i] {
                    size_t done = 0;
                    while (done < parts_[i].size) {
                        ssize_t w = pwrite(fd_, parts_[i].data + done, parts_[i].size - done, at[i] + (off_t)done);
                        if (w <= 0) { good[i] = 0; break; }
                        done += (size_t)w;
                    }
                });
            for (auto& x : th) x.join();
            for (char g : good) ok = ok && g;
        } else {
            for (int i = 0; i < nparts; i++) {
                size_t done = 0;
                while (ok && done < parts_[i].size) {
                    ssize_t w = ::write(fd_, parts_[i].data + done, parts_[i].size - done);
                    if (w <= 0) ok = false; else done += (size_t)w;
                }
            }
        }
        if (!ok) { err.code = CFRK_EIO; err.msg = "short write"; }
        return ok;
    }
    std::vector<RawBuf> parts_;
    int fd_ = -1;
    bool seekable_ = false;
    off_t off_ = 0;
    size_t bins_ = 0;
    std::unique_ptr<BinLabels> labels_;
    int nt_ = 1;
    bool sparse_ = false, first_ = true;
};

// ------------------------------------------------------------------------------------------
// GPU side of the file pipeline: input double buffer + row ring.
struct Pipeline {
    static constexpr size_t kSlotBytes = (size_t)128 << 20;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t done[2] = {}, drained[2] = {};
    char* d_in = nullptr; size_t cap_in = 0;
    int64_t* d_start = nullptr; int32_t* d_length = nullptr; int64_t* d_header = nullptr; size_t cap_reads = 0;
    // exact mode: unwrapped copy of the span (fasta_scan.cu launch_unwrap)
    char* d_packed = nullptr; size_t cap_packed = 0;
    int64_t* d_start2 = nullptr; int32_t* d_length2 = nullptr; size_t cap_reads2 = 0;
    // what the count kernels read: the raw span or its unwrapped copy
    const char* cur_bases = nullptr; const int64_t* cur_start = nullptr; const int32_t* cur_length = nullptr;
    size_t n_headers = 0;
    int32_t* d_rows[2] = {}; int32_t* h_rows[2] = {}; size_t cap_rows = 0;

    bool init(Err& err)
    {
        RF_CU(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
        RF_CU(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            RF_CU(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
            RF_CU(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
        }
        return true;
    }
    bool reserve(size_t in_bytes, size_t nreads, size_t row_bytes, Err& err)
    {
        if (in_bytes && in_bytes + CFRK_PAD > cap_in) {
            cudaFree(d_in);
            cap_in = in_bytes + CFRK_PAD + in_bytes / 8;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_in), cap_in));
        }
        if (nreads > cap_reads) {
            cudaFree(d_start); cudaFree(d_length); cudaFree(d_header);
            cap_reads = nreads + nreads / 4 + 64;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_start), cap_reads * 8));
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_length), cap_reads * 4));
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_header), cap_reads * 8));
        }
        const size_t need = row_bytes;
        if (row_bytes && need > cap_rows) {
            for (int i = 0; i < 2; i++) {
                cudaFree(d_rows[i]); if (h_rows[i]) cudaFreeHost(h_rows[i]);
                RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_rows[i]), need));
                RF_CU(cudaMallocHost(reinterpret_cast<void**>(&h_rows[i]), need));
            }
            cap_rows = need;
        }
        return true;
    }
    // Span of raw file bytes -> HBM, record table built THERE (fasta_scan.cu); the host gets the
    // (small) table back for its bookkeeping and never looks at a base.
    bool upload_and_scan(const char* h_in, size_t in_bytes, bool final_span, RecordIndex& ri, Err& err)
    {
        ri.start.clear(); ri.length.clear(); ri.header.clear();
        if (in_bytes == 0) return true;
        if (!reserve(in_bytes, std::max<size_t>(cap_reads, in_bytes / 64 + 64), 0, err)) return false;
        RF_CU(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, compute));
        RF_CU(cudaMemsetAsync(d_in + in_bytes, 0, CFRK_PAD, compute));
        int64_t out[2];
        for (;;) {
            cudaError_t e = cfrk::launch_fasta_scan(reinterpret_cast<const uint8_t*>(d_in), (int64_t)in_bytes, final_span,
                                                    d_header, d_start, d_length, (int64_t)cap_reads, out, compute);
            if (e != cudaSuccess) { err.code = CFRK_ECUDA; err.msg = std::string("fasta scan: ") + cudaGetErrorString(e); return false; }
            if (out[1] != 4) break;
            if (!reserve(in_bytes, (size_t)out[0], 0, err)) return false;   // more records than guessed: grow, rescan
        }
        if (out[1] == 1) { err.code = CFRK_EFORMAT; err.msg = "'>' inside a sequence line (grep -c over-counts nS in the reference, src/fastaIO.h:16)"; return false; }
        if (out[1] == 2) { err.code = CFRK_EFORMAT; err.msg = "sequence text before the first '>' header (undefined in the reference, src/fastaIO.h:49-52)"; return false; }
        if (out[1] == 3) { err.code = CFRK_EFORMAT; err.msg = "record longer than 2^31-1 bytes (length is int in the reference, src/tipos.h:26)"; return false; }
        const size_t nh = (size_t)out[0];
        n_headers = nh;
        cur_bases = d_in; cur_start = d_start; cur_length = d_length;
        const size_t complete = final_span ? nh : (nh ? nh - 1 : 0);
        ri.header.resize(nh); ri.start.resize(complete); ri.length.resize(complete);
        std::vector<int64_t> hdr(nh);
        if (nh) RF_CU(cudaMemcpyAsync(hdr.data(), d_header, nh * 8, cudaMemcpyDeviceToHost, compute));
        if (complete) {
            RF_CU(cudaMemcpyAsync(ri.start.data(), d_start, complete * 8, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaMemcpyAsync(ri.length.data(), d_length, complete * 4, cudaMemcpyDeviceToHost, compute));
        }
        RF_CU(cudaStreamSynchronize(compute));
        for (size_t i = 0; i < nh; i++) ri.header[i] = (size_t)hdr[i];
        return true;
    }
    // Exact mode: k-mers do not stop at line ends.  Replace the span by its unwrapped copy.
    bool unwrap_scanned(size_t in_bytes, size_t nreads, Err& err)
    {
        if (nreads == 0) return true;
        if (in_bytes + CFRK_PAD > cap_packed) {
            cudaFree(d_packed);
            cap_packed = in_bytes + CFRK_PAD + in_bytes / 8;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_packed), cap_packed));
        }
        if (nreads + 1 > cap_reads2) {
            cudaFree(d_start2); cudaFree(d_length2);
            cap_reads2 = nreads + nreads / 4 + 64;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_start2), cap_reads2 * 8));
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_length2), cap_reads2 * 4));
        }
        RF_CU(cudaMemsetAsync(d_packed, 0, in_bytes + CFRK_PAD, compute));
        cudaError_t e = cfrk::launch_unwrap(reinterpret_cast<const uint8_t*>(d_in), (int64_t)in_bytes, d_header,
                                            (int64_t)n_headers, d_start, (int64_t)nreads,
                                            reinterpret_cast<uint8_t*>(d_packed), d_start2, d_length2, compute);
        if (e != cudaSuccess) { err.code = CFRK_ECUDA; err.msg = std::string("unwrap: ") + cudaGetErrorString(e); return false; }
        cur_bases = d_packed; cur_start = d_start2; cur_length = d_length2;
        return true;
    }
    // Rows [0, nrows) of the span that upload_and_scan() left in HBM.
    bool count_scanned(const char* h_in, size_t in_bytes, const RecordIndex& ri, size_t nrows, int k, int mode,
                       int64_t chunk_size, int64_t index_base, CfrkWriter& w, Err& err)
    {
        if (nrows == 0) return true;
        if (mode == CFRK_MODE_COMPAT && std::find(ri.length.begin(), ri.length.end(), 0) != ri.length.end())
            return run(h_in, in_bytes, ri, nrows, k, mode, chunk_size, index_base, w, err);   // needs the packed layout
        if (mode == CFRK_MODE_EXACT && !unwrap_scanned(in_bytes, ri.start.size(), err)) return false;
        return count_rows(in_bytes, ri.start.size(), nrows, k, mode, chunk_size, index_base, w, err);
    }
    // k > 8: sparse rows (exact semantics) of the span that upload_and_scan() left in HBM
    bool count_scanned_sparse(size_t in_bytes, const RecordIndex& ri, size_t nrows, int k, CfrkWriter& w, Err& err)
    {
        if (nrows == 0) return true;
        const size_t nreads = ri.start.size();
        if (!unwrap_scanned(in_bytes, nreads, err)) return false;   // k > 8 is exact mode only
        int64_t cap = 0;
        for (int32_t l : ri.length) cap += std::max(0, l + 1 - k + 1);   // unwrapped text <= raw length + 1
        int64_t* d_rb = nullptr; int32_t* d_rc = nullptr; uint64_t* d_k = nullptr; uint32_t* d_c = nullptr;
        RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_rb), (nreads + 1) * 8));
        RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_rc), nreads * 4));
        RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_k), (size_t)std::max<int64_t>(cap, 1) * 8));
        RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_c), (size_t)std::max<int64_t>(cap, 1) * 4));
        int64_t total = 0;
        cudaError_t e = cfrk::launch_sparse(cur_bases, cfrk::FMT_ASCII, cur_start, cur_length, (int64_t)nreads, k, d_rb, d_rc, d_k, 8,
                                            d_c, cap, &total, compute);
        bool ok = e == cudaSuccess;
        if (!ok) { err.code = CFRK_ECUDA; err.msg = std::string("sparse path: ") + cudaGetErrorString(e); }
        std::vector<int64_t> rb(nrows + 1);
        std::vector<int32_t> rc(nrows);
        std::vector<uint64_t> kk;
        std::vector<uint32_t> cc;
        if (ok) {
            ok = cudaMemcpyAsync(rb.data(), d_rb, (nrows + 1) * 8, cudaMemcpyDeviceToHost, compute) == cudaSuccess &&
                 cudaMemcpyAsync(rc.data(), d_rc, nrows * 4, cudaMemcpyDeviceToHost, compute) == cudaSuccess &&
                 cudaStreamSynchronize(compute) == cudaSuccess;
            const size_t used = (size_t)rb[nrows];
            kk.resize(std::max<size_t>(used, 1)); cc.resize(std::max<size_t>(used, 1));
            ok = ok && cudaMemcpy(kk.data(), d_k, used * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
                 cudaMemcpy(cc.data(), d_c, used * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
            if (!ok) { err.code = CFRK_ECUDA; err.msg = "sparse path: device to host copy failed"; }
        }
        cudaFree(d_rb); cudaFree(d_rc); cudaFree(d_k); cudaFree(d_c);
        return ok && w.write_pairs(rb.data(), rc.data(), kk.data(), cc.data(), nrows, err);
    }
    // Host-side record table (the rare packed layout): upload bytes + table, then count.
    bool run(const char* h_in, size_t in_bytes, const RecordIndex& ri_in, size_t nrows, int k, int mode,
             int64_t chunk_size, int64_t index_base, CfrkWriter& w, Err& err)
    {
        if (nrows == 0) return true;
        // An EMPTY read makes the reference walk over the bytes that FOLLOW it in its batch layout
        // (kmer_device.cuh read_extent).  Raw file bytes have header lines there, so for such a
        // (rare) span the records are first compacted into the reference layout: text, one
        // separator, next text, ...
        const RecordIndex* rip = &ri_in;
        RecordIndex packed;
        std::vector<char> pack_buf;
        if (mode == CFRK_MODE_COMPAT &&
            std::find(ri_in.length.begin(), ri_in.length.end(), 0) != ri_in.length.end()) {
            size_t total = 0;
            for (int32_t l : ri_in.length) total += (size_t)l + 1;
            pack_buf.resize(total + CFRK_PAD);
            size_t wpos = 0;
            for (size_t i = 0; i < ri_in.start.size(); i++) {
                packed.start.push_back((int64_t)wpos);
                packed.length.push_back(ri_in.length[i]);
                memcpy(pack_buf.data() + wpos, h_in + ri_in.start[i], (size_t)ri_in.length[i]);
                wpos += (size_t)ri_in.length[i];
                pack_buf[wpos++] = '\n';
            }
            h_in = pack_buf.data();
            in_bytes = wpos;
            rip = &packed;
        }
        const RecordIndex& ri = *rip;
        const size_t nreads = ri.start.size();
        if (!reserve(in_bytes, nreads, 0, err)) return false;
        RF_CU(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, compute));
        RF_CU(cudaMemsetAsync(d_in + in_bytes, 0, CFRK_PAD, compute));
        RF_CU(cudaMemcpyAsync(d_start, ri.start.data(), nreads * 8, cudaMemcpyHostToDevice, compute));
        RF_CU(cudaMemcpyAsync(d_length, ri.length.data(), nreads * 4, cudaMemcpyHostToDevice, compute));
        cur_bases = d_in; cur_start = d_start; cur_length = d_length;
        const bool ok = count_rows(in_bytes, nreads, nrows, k, mode, chunk_size, index_base, w, err);
        RF_CU(cudaStreamSynchronize(compute));   // pack_buf / ri must outlive the copies
        return ok;
    }
    // d_in / d_start / d_length hold the span: kernels slice by slice, rows through the ring to the writer.
    bool count_rows(size_t in_bytes, size_t nreads, size_t nrows, int k, int mode, int64_t chunk_size,
                    int64_t index_base, CfrkWriter& w, Err& err)
    {
        const size_t bins = (size_t)1 << (2 * k), row_bytes = bins * 4;
        // ring slots: as large as the span needs, at most kSlotBytes
        if (!reserve(0, 0, std::max(row_bytes, std::min(kSlotBytes, nrows * row_bytes)), err)) return false;
        const size_t rpt = (size_t)cfrk::dense_reads_per_tile(k);
        size_t slice = std::max<size_t>(1, cap_rows / row_bytes);
        slice = std::max(rpt, slice / rpt * rpt);

        const size_t nslices = (nrows + slice - 1) / slice;
        for (size_t s = 0; s <= nslices; s++) {
            if (s < nslices) {
                const int slot = (int)(s & 1);
                const size_t r0 = s * slice, r1 = std::min(nrows, r0 + slice);
                // slot reuse: the host finished formatting slice s-2 before we get here (below);
                // the kernel must not overwrite d_rows[slot] before the D2H of slice s-2 is done
                if (s >= 2) RF_CU(cudaStreamWaitEvent(compute, drained[slot], 0));
                cudaError_t e = cfrk::launch_dense(cur_bases, cfrk::FMT_ASCII, cur_start, cur_length, (int64_t)in_bytes, (int64_t)nreads,
                                                   (int64_t)r0, (int64_t)r1, k, mode, chunk_size, index_base,
                                                   d_rows[slot], compute);
                if (e != cudaSuccess) { err.code = CFRK_ECUDA; err.msg = std::string("dense_count_kernel: ") + cudaGetErrorString(e); return false; }
                RF_CU(cudaEventRecord(done[slot], compute));
                RF_CU(cudaStreamWaitEvent(copy, done[slot], 0));
                RF_CU(cudaMemcpyAsync(h_rows[slot], d_rows[slot], (r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, copy));
                RF_CU(cudaEventRecord(drained[slot], copy));
            }
            if (s >= 1) {  // format slice s-1 while slice s is on the GPU
                const size_t q = s - 1;
                const int slot = (int)(q & 1);
                const size_t r0 = q * slice, r1 = std::min(nrows, r0 + slice);
                RF_CU(cudaEventSynchronize(drained[slot]));
                if (!w.write_rows(h_rows[slot], r1 - r0, err)) return false;
            }
        }
        RF_CU(cudaStreamSynchronize(compute));
        return true;
    }
    ~Pipeline()
    {
        cudaFree(d_in); cudaFree(d_start); cudaFree(d_length); cudaFree(d_header);
        cudaFree(d_packed); cudaFree(d_start2); cudaFree(d_length2);
        for (int i = 0; i < 2; i++) {
            cudaFree(d_rows[i]);
            if (h_rows[i]) cudaFreeHost(h_rows[i]);
            if (done[i]) cudaEventDestroy(done[i]);
            if (drained[i]) cudaEventDestroy(drained[i]);
        }
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
    }
};

bool run_file(const char* fasta, const char* out_path, int k, int nt, int64_t chunk_size, int flags, int device,
              Err& err)
{
    const bool all_rows = flags & CFRK_RUN_ALL_ROWS;
    const int mode = (flags & CFRK_RUN_EXACT) ? CFRK_MODE_EXACT : CFRK_MODE_COMPAT;
    RF_CU(cudaSetDevice(device));

    Trace tr;
    tr.mark("cudaSetDevice");
    // 64 MiB streaming window (+ as much headroom for carried records); small files get small
    // pinned buffers: pinning costs ~0.5 ms/MiB and test-sized inputs should start instantly
    size_t window = (size_t)64 << 20;
    {
        struct stat st0;
        if (stat(fasta, &st0) == 0 && (size_t)st0.st_size < window)
            window = std::max<size_t>((size_t)1 << 20, ((size_t)st0.st_size + 4095) & ~(size_t)4095);
    }
    const size_t kChunk = window, kHeadroom = window;
    FastaStreamer rd;
    if (!rd.open(fasta, kChunk, kHeadroom, err)) return false;
    tr.mark("streamer open (pinned ring)");
    CfrkWriter w;
    if (!w.open(out_path, k, nt, flags & CFRK_RUN_SPARSE, err)) return false;
    Pipeline gpu;
    if (!gpu.init(err)) return false;
    tr.mark("pipeline init");

    RecordIndex ri;
    size_t carry = 0;            // bytes of unfinished records prepended to the current buffer
    int64_t reads_done = 0;      // index of the first record of the current span
    int64_t tail_off = -1;       // file offset of the header that opens the last (partial) chunk
    size_t span_file_off = 0;    // file offset of data[0]
    for (size_t i = 0;; i++) {
        FastaStreamer::Buf* b = rd.acquire(i);
        char* data = b->base + rd.headroom() - carry;
        const size_t n = carry + b->n;
        tr.mark("buffer acquired", n);
        if (!gpu.upload_and_scan(data, n, b->eof, ri, err)) return false;
        tr.mark("uploaded + scanned on device");
        const size_t m = ri.start.size();  // records whose end is known

        size_t keep_from;  // data[keep_from, n) goes in front of the next buffer
        if (all_rows) {
            // Rows for all complete records but a held-back tail: the record after the last row
            // is needed for its spill into that row, and an empty read walks up to 1024+k bytes
            // into its successors (kmer_device.cuh read_extent).  The held-back records are
            // counted again at the head of the next span.
            size_t nrows = m;
            if (!b->eof) {
                size_t held = 0;
                while (nrows > 0 && (nrows == m || held < (size_t)cfrk::kRefBlockThreads + 64)) {
                    nrows--;
                    held += (size_t)ri.length[nrows] + 1;
                }
            }
            if (k > CFRK_CLI_DENSE_MAX_K ? !gpu.count_scanned_sparse(n, ri, nrows, k, w, err)
                                     : !gpu.count_scanned(data, n, ri, nrows, k, mode, chunk_size, reads_done, w, err)) return false;
            reads_done += (int64_t)nrows;
            tr.mark("rows counted + written");
            keep_from = b->eof ? n : (m ? ri.header[nrows] : 0);
        } else {
            for (size_t r = 0; r < m; r++)
                if ((reads_done + (int64_t)r) % chunk_size == 0) tail_off = (int64_t)(span_file_off + ri.header[r]);
            reads_done += (int64_t)m;
            keep_from = b->eof ? n : (ri.header.empty() ? 0 : ri.header[m]);
        }
        if (b->eof) { rd.release(i); break; }
        carry = n - keep_from;
        if (carry > rd.headroom()) {
            err.code = CFRK_EFORMAT;
            err.msg = "a single FASTA record (plus the records held back with it) exceeds the streaming window";
            return false;
        }
        memcpy(rd.next_base(i) + rd.headroom() - carry, data + keep_from, carry);
        span_file_off += keep_from;
        rd.release(i);
    }

    if (!all_rows) {
        // reference: only the remainder chunk reaches the file; nothing when nS % chunkSize == 0
        const int64_t nS = reads_done;
        if (nS % chunk_size != 0 && tail_off >= 0) {
            const size_t bytes = rd.file_size() - (size_t)tail_off;
            char* h = nullptr;
            if (cudaMallocHost(reinterpret_cast<void**>(&h), bytes + CFRK_PAD) != cudaSuccess) {
                cudaGetLastError();
                err.code = CFRK_ENOMEM; err.msg = "cudaMallocHost(tail chunk)"; return false;
            }
            size_t got = 0;
            while (got < bytes) {
                ssize_t r = pread(rd.fd(), h + got, bytes - got, (off_t)((size_t)tail_off + got));
                if (r <= 0) break;
                got += (size_t)r;
            }
            bool ok = got == bytes && gpu.upload_and_scan(h, bytes, true, ri, err);
            if (got != bytes) { err.code = CFRK_EIO; err.msg = "short read of the tail chunk"; }
            tr.mark("tail chunk scanned", bytes);
            if (ok) ok = k > CFRK_CLI_DENSE_MAX_K
                             ? gpu.count_scanned_sparse(bytes, ri, ri.start.size(), k, w, err)
                             : gpu.count_scanned(h, bytes, ri, ri.start.size(), k, mode, chunk_size, (nS / chunk_size) * chunk_size, w, err);
            tr.mark("tail rows counted + written");
            cudaFreeHost(h);
            if (!ok) return false;
        }
    }
    return true;
}

}  // namespace

extern "C" int cfrk_run_file(const char* fasta_path, const char* out_path, int k, int nt, int64_t chunk_size,
                             int flags, int device)
{
    Err err;
    if (!fasta_path || !out_path) { err.code = CFRK_EINVAL; err.msg = "null path"; }
    else if (k < 1 || k > CFRK_SPARSE_MAX_K) { err.code = CFRK_EINVAL; err.msg = "k must be in 1..31"; }
    else if (k > CFRK_CLI_DENSE_MAX_K && (flags & (CFRK_RUN_SPARSE | CFRK_RUN_EXACT)) != (CFRK_RUN_SPARSE | CFRK_RUN_EXACT)) {
        err.code = CFRK_EINVAL;
        err.msg = "k > 8: rows have 4^k bins; pass --sparse --exact (non-zero bins only, intended semantics)";
    }
    else if (chunk_size <= 0) { err.code = CFRK_EINVAL; err.msg = "chunkSize must be positive"; }
    else {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
            cudaGetLastError();
            err.code = CFRK_ECUDA; err.msg = "no such CUDA device (this library has no CPU fallback)";
        } else {
            run_file(fasta_path, out_path, k, nt, chunk_size, flags, device, err);
        }
    }
    if (err.code != CFRK_OK) cfrk::set_last_error(err.msg);
    return err.code;
}
