// runfile.cu -- cfrk_run_file(): FASTA file -> GPU count -> .cfrk text, the whole of the
// reference's main() (src/main.cu:232-305) re-designed around one or several GPUs.
//
//   reader   SpanReader: ONE thread reads the file (FASTA or 4-line FASTQ, plain or gzip) into a pool of PINNED buffers and
//            cuts it into SPANS at header lines, on the host, by looking at a few hundred bytes at the
//            end of each window (the bases are never touched).  A span = complete records; its
//            last >= 1088 text bytes of records are LOOKAHEAD: they get no rows in this span (they
//            open the next one) but are uploaded with it, because the last row needs the spill of
//            the read after it and an empty read walks up to 1024+k bytes into its successors
//            (kmer_device.cuh read_extent).  A record larger than the window makes the buffer grow.
//            The raw bytes are what the GPU consumes (CFRK_FMT_ASCII): the record text of the
//            reference parser (src/fastaIO.h:38-69: every non-header line, '\n' included, minus the
//            last byte) is a contiguous piece of the file, so a record is (start, length) into the
//            raw bytes -- newline-as-base, CRLF, missing final newline all fall out.
//            Replaces popen("grep -c") + getline + 3 malloc/record + strcat + per-base switch
//            + ProcessData + SelectChunk (src/fastaIO.h:12-148, src/main.cu:110-206).
//   workers  two host threads per GPU (the reference: devCount pthreads on ONE GPU,
//            src/main.cu:208-230,277-289), each with its own streams, device buffers and pinned row
//            ring; spans are dealt out in file order.  Per span: H2D, record table built on the GPU
//            (fasta_scan.cu), dense rows slice by slice (or sparse rows) into the ring.  While one
//            worker formats and writes span i, the other has span i+1 uploaded, scanned and counting.
//   order    rows reach the file in read order: the global index of a span's first read is handed
//            from span to span as soon as a span is scanned (chunk openers depend on it), and the
//            writer is entered span by span.
//   writer   nt pooled threads format rows ("bin:count ", src/main.cu:53-55) into private buffers that
//            are written in order; "\n" before every row but the first, none at EOF.  --sparse rows of
//            k >= 5 are compacted ON THE DEVICE to (bin, count) pairs: dense rows never cross PCIe.
//
// Default (compat) output = only reads [ (nS/chunkSize)*chunkSize, nS ), because the reference
// re-opens the output with "w" for the remainder chunk (src/main.cu:34,303-305): pass 1 scans the
// file for headers only (all GPUs), pass 2 streams that tail.  CFRK_RUN_ALL_ROWS streams every read.
#include "../../include/cfrk_b200.h"
#include "kernels.h"
#include "kmer_device.cuh"
#include "internal.h"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <thread>
#include <unistd.h>
#include <vector>
#include <zlib.h>

namespace cfrk {
extern void count_launch();
}

namespace {

struct Err {
    int code = CFRK_OK;
    std::string msg;
};

// CFRK_TRACE=1: wall-clock trace of the file pipeline on stderr (the reference only has
// commented-out time(NULL) probes, src/main.cu:259-268,302-306)
struct Trace {
    bool on = getenv("CFRK_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    std::mutex mu;
    void mark(const char* what, size_t a = 0, size_t b = 0)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::lock_guard<std::mutex> lk(mu);
        fprintf(stderr, "[cfrk trace] %9.3f ms %s %zu %zu\n", std::chrono::duration<double, std::milli>(now - t0).count(), what, a, b);
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

constexpr size_t kLookText = (size_t)cfrk::kRefBlockThreads + 64;   // text bytes of records held back as lookahead

// ------------------------------------------------------------------------------------------
// Input: a plain file (pread) or a gzip file (zlib; the reference includes <zlib.h> but never uses it,
// src/fastaIO.h:7).  Sequential reads plus a restart at an (uncompressed) offset for the tail pass.
class Source {
public:
    bool open(const char* path, Err& err)
    {
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) { err.code = CFRK_EIO; err.msg = std::string("cannot open ") + path; return false; }
        unsigned char magic[2] = {0, 0};
        const ssize_t got = pread(fd_, magic, 2, 0);
        struct stat st;
        if (fstat(fd_, &st) != 0) { err.code = CFRK_EIO; err.msg = "fstat failed"; return false; }
        size_ = (size_t)st.st_size;
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            gz_ = gzdopen(dup(fd_), "rb");
            if (!gz_) { err.code = CFRK_EIO; err.msg = "gzdopen failed"; return false; }
            gzbuffer(gz_, 1 << 20);
        }
        return true;
    }
    bool is_gzip() const { return gz_ != nullptr; }
    // '@' as the first byte of the (uncompressed) input: 4-line FASTQ records instead of FASTA
    bool detect_fastq(Err& err)
    {
        char c = 0; size_t got = 0; bool eof = false;
        if (!restart(0, err) || !read(&c, 1, &got, &eof, err)) return false;
        fastq_ = got == 1 && c == '@';
        return restart(0, err);
    }
    bool is_fastq() const { return fastq_; }
    size_t size_hint() const { return gz_ ? size_ * 4 : size_; }   // a guess for gzip; only sizes the buffers
    bool restart(size_t off, Err& err)
    {
        pos_ = off;
        if (gz_ && gzseek(gz_, (z_off_t)off, SEEK_SET) < 0) { err.code = CFRK_EIO; err.msg = "gzseek failed"; return false; }
        return true;
    }
    // up to n bytes; *eof when the input ends inside or right after them.  false = I/O error.
    bool read(char* dst, size_t n, size_t* got, bool* eof, Err& err)
    {
        *got = 0; *eof = false;
        while (*got < n) {
            ssize_t r;
            if (gz_) {
                r = gzread(gz_, dst + *got, (unsigned)std::min<size_t>(n - *got, (size_t)1 << 30));
                if (r < 0) { int en = 0; err.code = CFRK_EIO; err.msg = std::string("gzread: ") + gzerror(gz_, &en); return false; }
            } else {
                r = pread(fd_, dst + *got, n - *got, (off_t)(pos_ + *got));
                if (r < 0) { err.code = CFRK_EIO; err.msg = std::string("read error: ") + strerror(errno); return false; }
            }
            if (r == 0) { *eof = true; break; }
            *got += (size_t)r;
        }
        pos_ += *got;
        if (!gz_) {
            // a file that shrank under us is an error, not an end: the rows would be silently truncated
            if (*eof && pos_ < size_) { err.code = CFRK_EIO; err.msg = "input file was truncated while reading"; return false; }
            if (pos_ >= size_) *eof = true;
        }
        return true;
    }
    ~Source()
    {
        if (gz_) gzclose(gz_);
        if (fd_ >= 0) ::close(fd_);
    }

private:
    int fd_ = -1;
    gzFile gz_ = nullptr;
    size_t size_ = 0, pos_ = 0;
    bool fastq_ = false;
};

// ------------------------------------------------------------------------------------------
struct Span {
    char* buf = nullptr;       // pinned
    size_t cap = 0;
    size_t n = 0;              // bytes [0, n): complete records, uploaded
    size_t rows_end = 0;       // records whose header lies before rows_end get rows; the rest is lookahead
    size_t file_off = 0;       // (uncompressed) file offset of buf[0]
    bool final = false;
    int64_t index = 0;
};

// last q in [lo, p) with buf[q] == '>' at a line start, or SIZE_MAX
inline size_t prev_header(const char* buf, size_t lo, size_t p)
{
    while (p > lo) {
        const void* m = memrchr(buf + lo, '>', p - lo);
        if (!m) return SIZE_MAX;
        const size_t q = (size_t)(static_cast<const char*>(m) - buf);
        if (q == 0 || buf[q - 1] == '\n') return q;
        p = q;
    }
    return SIZE_MAX;
}

// FASTQ: last q in [lo, p) that begins a record: '@' at a line start whose line after next begins with '+'
// (a quality line may begin with '@' too, but is never followed two lines later by a '+' line).  *text =
// bytes of its sequence line, '\n' included.  Candidates whose next two lines are not in [0, fill) yet are
// skipped.
inline size_t prev_fastq_record(const char* buf, size_t lo, size_t p, size_t fill, size_t* text)
{
    while (p > lo) {
        const void* m = memrchr(buf + lo, '@', p - lo);
        if (!m) return SIZE_MAX;
        const size_t q = (size_t)(static_cast<const char*>(m) - buf);
        p = q;
        if (q != 0 && buf[q - 1] != '\n') continue;
        const void* n1 = memchr(buf + q, '\n', fill - q);
        if (!n1) continue;
        const size_t l1 = (size_t)(static_cast<const char*>(n1) - buf) + 1;      // sequence line
        const void* n2 = l1 < fill ? memchr(buf + l1, '\n', fill - l1) : nullptr;
        if (!n2) continue;
        const size_t l2 = (size_t)(static_cast<const char*>(n2) - buf) + 1;      // '+' line
        if (l2 >= fill || buf[l2] != '+') continue;
        if (text) *text = l2 - l1;
        return q;
    }
    return SIZE_MAX;
}

class SpanReader {
public:
    // window: bytes read per span; nslots: pinned buffers in flight
    bool open(Source* src, size_t window, int nslots, Err& err)
    {
        src_ = src; window_ = window;
        slots_.resize(nslots);
        for (Span& s : slots_)
            if (!grow(s, window + ((size_t)1 << 16), err)) return false;
        for (int i = 0; i < nslots; i++) free_.push_back(i);
        return true;
    }
    void start(size_t file_off)
    {
        file_off_ = file_off;
        th_ = std::thread([this] { run(); });
    }
    // next span in file order (nullptr: end of input or error -- see error())
    Span* next()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !ready_.empty() || done_; });
        if (ready_.empty()) return nullptr;
        Span* s = ready_.front();
        ready_.pop_front();
        return s;
    }
    void release(Span* s)
    {
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back((int)(s - slots_.data()));
        cv_.notify_all();
    }
    void stop()     // may be called by several failing workers and by the owner: only one joins
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            cv_.notify_all();
        }
        std::lock_guard<std::mutex> jk(join_mu_);
        if (th_.joinable()) th_.join();
    }
    const Err& error() const { return err_; }
    ~SpanReader()
    {
        stop();
        for (Span& s : slots_) if (s.buf) cudaFreeHost(s.buf);
    }

private:
    // (re)allocate the pinned buffer of a span, keeping its first `keep` bytes
    static bool grow(Span& s, size_t cap, Err& err, size_t keep = 0)
    {
        char* nb = nullptr;
        if (cudaMallocHost(reinterpret_cast<void**>(&nb), cap + CFRK_PAD) != cudaSuccess) {
            cudaGetLastError();
            err.code = CFRK_ENOMEM; err.msg = "cudaMallocHost(stream buffer)";
            return false;
        }
        if (s.buf) { memcpy(nb, s.buf, keep); cudaFreeHost(s.buf); }
        s.buf = nb; s.cap = cap;
        return true;
    }
    void finish()
    {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
        cv_.notify_all();
    }
    void run()
    {
        std::vector<char> carry;      // bytes that open the next span
        int64_t index = 0;
        bool eof = false;
        while (!eof) {
            int si;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !free_.empty(); });
                if (stop_) { done_ = true; cv_.notify_all(); return; }
                si = free_.back(); free_.pop_back();
            }
            Span& s = slots_[si];
            s.n = 0; s.final = false; s.index = index; s.file_off = file_off_;
            if (carry.size() + window_ > s.cap && !grow(s, carry.size() + window_ + ((size_t)1 << 16), err_)) return finish();
            memcpy(s.buf, carry.data(), carry.size());
            size_t fill = carry.size();
            for (;;) {
                size_t got = 0;
                if (!src_->read(s.buf + fill, s.cap - fill, &got, &eof, err_)) return finish();
                fill += got;
                if (eof) {
                    // FASTQ: blank lines at the end of the file are not a fifth line of the last record
                    if (src_->is_fastq()) while (fill > 1 && s.buf[fill - 1] == '\n' && (s.buf[fill - 2] == '\n' || s.buf[fill - 2] == '\r')) fill--;
                    s.n = fill; s.rows_end = fill; s.final = true;
                    break;
                }
                // cut at the last header line; hold back >= kLookText text bytes of complete records
                const bool fq = src_->is_fastq();
                const size_t h_last = fq ? prev_fastq_record(s.buf, 1, fill, fill, nullptr) : prev_header(s.buf, 1, fill);
                size_t cut = SIZE_MAX;
                if (h_last != SIZE_MAX) {
                    size_t held = 0, next = h_last, h = h_last;
                    while (h > 0) {
                        size_t text = 0;
                        h = fq ? prev_fastq_record(s.buf, 0, h, fill, &text) : prev_header(s.buf, 0, h);
                        if (h == SIZE_MAX) break;
                        if (!fq) {
                            const void* eol = memchr(s.buf + h, '\n', next - h);
                            text = eol ? next - ((size_t)(static_cast<const char*>(eol) - s.buf) + 1) : 0;
                        }
                        held += text;
                        next = h;
                        if (held >= kLookText) { cut = h; break; }
                    }
                }
                if (cut != SIZE_MAX && cut > 0) {
                    s.n = h_last; s.rows_end = cut;
                    carry.assign(s.buf + cut, s.buf + fill);
                    file_off_ += cut;
                    break;
                }
                // one record (plus its lookahead) larger than the buffer: make room and read on
                if (!grow(s, s.cap * 2, err_, fill)) return finish();
            }
            if (s.final) carry.clear();
            index++;
            {
                std::lock_guard<std::mutex> lk(mu_);
                ready_.push_back(&s);
                cv_.notify_all();
            }
        }
        finish();
    }
    Source* src_ = nullptr;
    size_t window_ = 0, file_off_ = 0;
    std::vector<Span> slots_;
    std::vector<int> free_;
    std::deque<Span*> ready_;
    std::thread th_;
    std::mutex mu_, join_mu_;
    std::condition_variable cv_;
    bool stop_ = false, done_ = false;
    Err err_;
};

// ------------------------------------------------------------------------------------------
// Record table of a span, as fasta_scan.cu produces it on the device.
struct RecordIndex {
    std::vector<int64_t> start;   // every record of the span (spans end at a header line or at EOF)
    std::vector<int32_t> length;
    std::vector<size_t> header;   // position of each header's '>'
};

// ------------------------------------------------------------------------------------------
// persistent formatting threads (the writer used to create and join nt threads per slice)
class ThreadPool {
public:
    explicit ThreadPool(int n)
    {
        for (int i = 0; i < n; i++) th_.emplace_back([this, i] { loop(i); });
    }
    int size() const { return (int)th_.size(); }
    // fn(part) for part in [0, nparts), nparts <= size(); returns when all are done
    void run(int nparts, const std::function<void(int)>& fn)
    {
        std::unique_lock<std::mutex> lk(mu_);
        fn_ = &fn; nparts_ = nparts; pending_ = nparts; gen_++;
        cv_.notify_all();
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~ThreadPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
            cv_.notify_all();
        }
        for (auto& t : th_) t.join();
    }

private:
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
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
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int nparts_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool quit_ = false;
};

// ------------------------------------------------------------------------------------------
// .cfrk text writer
// "bin:" labels as fixed 8-byte records (bins < 65536 -> at most "65535:"): one 64-bit store per token;
// and the text of a row whose counts are all one digit ("0:0 1:0 2:0 ... "), with the place of every digit
struct BinLabels {
    std::vector<uint64_t> rec;
    std::vector<uint8_t> len;
    std::string tmpl;
    std::vector<uint32_t> pos;
    explicit BinLabels(size_t bins) : rec(bins), len(bins), pos(bins)
    {
        char tmp[16];
        for (size_t b = 0; b < bins; b++) {
            memset(tmp, 0, sizeof tmp);
            len[b] = (uint8_t)snprintf(tmp, sizeof tmp, "%zu:", b);
            memcpy(&rec[b], tmp, 8);
            tmpl.append(tmp, len[b]);
            pos[b] = (uint32_t)tmpl.size();
            tmpl.append("0 ");
        }
    }
};

// output scratch that is NOT value-initialised (a std::vector would zero-fill the worst-case size
// of every slice: measured as most of the writer's time) and is kept between slices
struct RawBuf {
    char* data = nullptr;
    size_t size = 0, cap = 0;
    bool reserve(size_t n)
    {
        if (n > cap) {
            free(data);
            data = static_cast<char*>(malloc(n));
            cap = data ? n : 0;
        }
        size = 0;
        return data != nullptr;
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
inline int ndigits(uint64_t u)
{
    int n = 1;
    while (u >= 10) { u /= 10; n++; }
    return n;
}

// One row -> text at p.  EXACT = false may store up to 6 bytes past the end of the row's text (the 8-byte label
// records); EXACT = true never does (rows at the end of a thread's piece of the mapped file).
template <bool EXACT>
inline char* put_token(char* p, const BinLabels& lab, size_t b, int32_t v)
{
    if (EXACT) memcpy(p, &lab.rec[b], lab.len[b]); else memcpy(p, &lab.rec[b], 8);
    p = put_int(p + lab.len[b], v);
    *p++ = ' ';
    return p;
}
char* format_row(const int32_t* row, size_t bins, const BinLabels& lab, bool sparse, bool exact, char* p)
{
    if (!sparse) {
        // every count in 0..9 (the usual row): the row's text is the template with the digits patched in
        uint32_t over = 0;
        for (size_t b = 0; b < bins; b++) over |= (uint32_t)row[b] | (uint32_t)(9 - row[b]);
        if (!(over >> 31)) {
            memcpy(p, lab.tmpl.data(), lab.tmpl.size());
            const uint32_t* pos = lab.pos.data();
            for (size_t b = 0; b < bins; b++) p[pos[b]] = (char)('0' + row[b]);
            return p + lab.tmpl.size();
        }
        if (exact) for (size_t b = 0; b < bins; b++) p = put_token<true>(p, lab, b, row[b]);
        else for (size_t b = 0; b < bins; b++) p = put_token<false>(p, lab, b, row[b]);
        return p;
    }
    // non-zero bins: index list without a data-dependent branch (43 % of the bins of a 150-bp read at k = 4 are
    // zero: the branchy loop mispredicts on every other bin), then the tokens
    static thread_local std::vector<uint16_t> idx_buf;
    if (idx_buf.size() < bins + 1) idx_buf.resize(bins + 1);
    uint16_t* idx = idx_buf.data();
    size_t m = 0;
    for (size_t b = 0; b < bins; b++) { idx[m] = (uint16_t)b; m += row[b] != 0; }
    if (exact) for (size_t i = 0; i < m; i++) p = put_token<true>(p, lab, idx[i], row[idx[i]]);
    else for (size_t i = 0; i < m; i++) p = put_token<false>(p, lab, idx[i], row[idx[i]]);
    return p;
}
// bytes of format_row()'s text (the host twin of row_text_bytes_kernel)
int32_t row_text_bytes(const int32_t* row, size_t bins, const BinLabels& lab, bool sparse)
{
    int64_t n = 0;
    for (size_t b = 0; b < bins; b++) {
        const int32_t v = row[b];
        if (sparse && v == 0) continue;
        n += lab.len[b] + 1 + (v < 0 ? 1 + ndigits((uint64_t)(-(int64_t)v)) : ndigits((uint64_t)v));
    }
    return (int32_t)n;
}
template <typename KeyT>
inline char* format_pairs(const KeyT* kk, const uint32_t* cc, int32_t n, char* p)
{
    for (int32_t i = 0; i < n; i++) {
        char tmp[24];
        int l = 0;
        uint64_t u = (uint64_t)kk[i];
        do { tmp[l++] = (char)('0' + u % 10); u /= 10; } while (u);
        while (l) *p++ = tmp[--l];
        *p++ = ':';
        p = put_int(p, (int32_t)cc[i]);
        *p++ = ' ';
    }
    return p;
}

class CfrkWriter {
public:
    bool open(const char* path, int k, int nt, bool sparse, Err& err)
    {
        // fopen(path, "w"), src/main.cu:34.  Read-write so that the file can be mapped; a target that cannot be
        // opened that way (a write-only pipe) is written sequentially.
        fd_ = ::open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
        const bool rdwr = fd_ >= 0;
        if (fd_ < 0) fd_ = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd_ < 0) { err.code = CFRK_EIO; err.msg = std::string("cannot open output ") + path; return false; }
        struct stat st;
        seekable_ = lseek(fd_, 0, SEEK_CUR) != (off_t)-1;
        const char* ev = getenv("CFRK_WRITER");          // "pwrite": private buffers + pwrite (round 1) for A/B runs
        mapped_ = rdwr && seekable_ && fstat(fd_, &st) == 0 && S_ISREG(st.st_mode) && !(ev && strcmp(ev, "pwrite") == 0);
        bins_ = k <= CFRK_CLI_DENSE_MAX_K ? (size_t)1 << (2 * k) : 0;
        if (bins_) labels_.reset(new BinLabels(bins_));
        nt_ = std::max(1, std::min(nt, 64));
        pool_.reset(new ThreadPool(nt_));
        parts_ = std::vector<RawBuf>(nt_);
        sparse_ = sparse;
        return true;
    }
    bool sized() const { return mapped_; }   // text sizes of the rows wanted (they are computed on the GPU)
    size_t bins() const { return bins_; }
    const BinLabels& labels() const { return *labels_; }
    ThreadPool& pool() { return *pool_; }
    int threads() const { return nt_; }

    // text_bytes[r] = bytes of row r's tokens (row_text_bytes_kernel), or null
    bool write_rows(const int32_t* rows, size_t nrows, const int32_t* text_bytes, Err& err)
    {
        if (!nrows) return true;
        if (mapped_ && text_bytes) {
            const int m = write_mapped(nrows, text_bytes, [&](size_t r, char* p, bool exact) {
                return format_row(rows + r * bins_, bins_, *labels_, sparse_, exact, p); }, err);
            if (m >= 0) return m != 0;
        }
        const int nt = (int)std::min<size_t>((size_t)nt_, nrows);
        std::vector<char> good(nt, 1);
        const bool first_file = first_;
        pool_->run(nt, [&](int t) {
            const size_t a = nrows * t / nt, b = nrows * (t + 1) / nt;
            RawBuf& out = parts_[t];
            // worst case per token: 8-byte label store + 11 digits + space
            if (!out.reserve((b - a) * (bins_ * 20 + 1) + 16)) { good[t] = 0; return; }
            char* p = out.data;
            for (size_t r = a; r < b; r++) {
                if (!(first_file && r == 0)) *p++ = '\n';
                p = format_row(rows + r * bins_, bins_, *labels_, sparse_, false, p);
            }
            out.size = (size_t)(p - out.data);
        });
        first_ = false;
        for (char g : good) if (!g) { err.code = CFRK_ENOMEM; err.msg = "out of memory for the text of a row slice"; return false; }
        return flush_parts(nt, err);
    }
    // rows given as (key, count) pairs: the same "bin:count " tokens, non-zero bins only
    template <typename KeyT>
    bool write_pairs(const int64_t* row_begin, const int32_t* row_count, const KeyT* keys,
                     const uint32_t* counts, size_t nrows, const int32_t* text_bytes, Err& err)
    {
        if (!nrows) return true;
        if (mapped_ && text_bytes) {
            const int m = write_mapped(nrows, text_bytes, [&](size_t r, char* p, bool) {
                return format_pairs<KeyT>(keys + row_begin[r], counts + row_begin[r], row_count[r], p); }, err);
            if (m >= 0) return m != 0;
        }
        const int nt = (int)std::min<size_t>((size_t)nt_, nrows);
        std::vector<char> good(nt, 1);
        const bool first_file = first_;
        pool_->run(nt, [&](int t) {
            const size_t a = nrows * t / nt, b = nrows * (t + 1) / nt;
            size_t npairs = 0;
            for (size_t r = a; r < b; r++) npairs += (size_t)row_count[r];
            RawBuf& out = parts_[t];
            if (!out.reserve(npairs * 33 + (b - a) + 1)) { good[t] = 0; return; }   // 20 + 1 + 10 + 1 per token, '\n' per row
            char* p = out.data;
            for (size_t r = a; r < b; r++) {
                if (!(first_file && r == 0)) *p++ = '\n';
                p = format_pairs<KeyT>(keys + row_begin[r], counts + row_begin[r], row_count[r], p);
            }
            out.size = (size_t)(p - out.data);
        });
        first_ = false;
        for (char g : good) if (!g) { err.code = CFRK_ENOMEM; err.msg = "out of memory for the text of a row slice"; return false; }
        return flush_parts(nt, err);
    }
    ~CfrkWriter() { if (fd_ >= 0) ::close(fd_); }

private:
    // Rows whose text sizes are known: the file grows by exactly that much, the new piece is mapped and the pool
    // threads format STRAIGHT INTO IT -- no private buffer, no copy, and the page allocation of the new pages runs on
    // all threads, where pwrite()s into one file take turns on the inode lock (tools/host/pwrite_scaling.c: 2.1-2.6 GB/s
    // into tmpfs whatever the thread count; the mapping scales with the threads).
    // 1: written, 0: failed (err), -1: this file cannot be mapped -- the caller (and every later slice) takes the buffers
    template <typename RowFn>
    int write_mapped(size_t nrows, const int32_t* text_bytes, RowFn&& row_fn, Err& err)
    {
        row_off_.resize(nrows + 1);
        int64_t o = 0;
        for (size_t r = 0; r < nrows; r++) {
            row_off_[r] = o;
            if (text_bytes[r] < 0) { err.code = CFRK_EIO; err.msg = "internal: negative row text size"; return 0; }
            o += (int64_t)text_bytes[r] + ((first_ && r == 0) ? 0 : 1);
        }
        row_off_[nrows] = o;
        const int64_t total = o;
        const bool first_file = first_;
        if (total == 0) { first_ = false; return 1; }
        struct statvfs vfs;
        if (fstatvfs(fd_, &vfs) == 0 && vfs.f_blocks > 0 && (uint64_t)vfs.f_bavail * vfs.f_frsize < (uint64_t)total) {
            err.code = CFRK_EIO; err.msg = "no space left for the output"; return 0;   // a mapped write would fault instead
        }
        if (ftruncate(fd_, off_ + (off_t)total) != 0) { err.code = CFRK_EIO; err.msg = std::string("cannot grow the output: ") + strerror(errno); return 0; }
        const off_t pg = (off_t)sysconf(_SC_PAGESIZE);
        const off_t map_at = off_ & ~(pg - 1);
        const size_t map_len = (size_t)(off_ + (off_t)total - map_at);
        void* m = mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, map_at);
        if (m == MAP_FAILED) {          // a file system without shared writable mappings: back to buffers + pwrite
            if (ftruncate(fd_, off_) != 0) { err.code = CFRK_EIO; err.msg = std::string("cannot shrink the output: ") + strerror(errno); return 0; }
            mapped_ = false;
            return -1;
        }
        first_ = false;
        char* base = static_cast<char*>(m) + (off_ - map_at);
        // pieces of about equal text, cut at rows
        const int nt = (int)std::min<size_t>((size_t)nt_, nrows);
        std::vector<size_t> cut(nt + 1);
        for (int t = 0; t <= nt; t++)
            cut[t] = t == nt ? nrows : (size_t)(std::lower_bound(row_off_.begin(), row_off_.begin() + nrows, total * t / nt) - row_off_.begin());
        std::vector<char> good(nt, 1);
        pool_->run(nt, [&](int t) {
            const size_t a = cut[t], b = cut[t + 1];
            if (a >= b) return;
            char* p = base + row_off_[a];
            char* const end = base + row_off_[b];
            for (size_t r = a; r < b; r++) {
                if (!(first_file && r == 0)) *p++ = '\n';
                p = row_fn(r, p, end - (p + text_bytes[r]) < 8);
                if (p != base + row_off_[r + 1]) { good[t] = 0; return; }   // the GPU's size and the formatter disagree: stop
            }
        });
        munmap(m, map_len);
        off_ += (off_t)total;
        for (char g : good) if (!g) { err.code = CFRK_EIO; err.msg = "internal: row text size mismatch"; return 0; }
        return 1;
    }
    // formatted parts -> file, in order.  Regular files: every part is written at its own offset
    // by its own pool thread (pwrite); pipes (the Swift stdout form): sequential write.
    bool flush_parts(int nparts, Err& err)
    {
        bool ok = true;
        if (seekable_) {
            std::vector<off_t> at(nparts);
            for (int i = 0; i < nparts; i++) { at[i] = off_; off_ += (off_t)parts_[i].size; }
            std::vector<char> good(nparts, 1);
            pool_->run(nparts, [&](int i) {
                size_t done = 0;
                while (done < parts_[i].size) {
                    ssize_t w = pwrite(fd_, parts_[i].data + done, parts_[i].size - done, at[i] + (off_t)done);
                    if (w <= 0) { good[i] = 0; break; }
                    done += (size_t)w;
                }
            });
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
    std::vector<int64_t> row_off_;
    std::unique_ptr<ThreadPool> pool_;
    int fd_ = -1;
    bool seekable_ = false, mapped_ = false;
    off_t off_ = 0;
    size_t bins_ = 0;
    std::unique_ptr<BinLabels> labels_;
    int nt_ = 1;
    bool sparse_ = false, first_ = true;
};

// ------------------------------------------------------------------------------------------
// Spans are processed by several workers but their effects are ordered: a value handed from span i
// to span i+1 (index of the first read), and a turn (the writer).
class Sequencer {
public:
    // blocks until span `i` may go; false if the run was aborted
    bool wait_turn(int64_t i)
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return abort_ || turn_ == i; });
        return !abort_;
    }
    void end_turn(int64_t i)
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (turn_ == i) turn_ = i + 1;
        cv_.notify_all();
    }
    bool get_reads_before(int64_t i, int64_t* v)
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return abort_ || known_ >= i; });
        if (abort_) return false;
        *v = reads_[(size_t)(i % kRing)];
        return true;
    }
    void set_reads_before(int64_t i, int64_t v)     // called in span order (span i-1's worker, after its scan)
    {
        std::lock_guard<std::mutex> lk(mu_);
        reads_[(size_t)(i % kRing)] = v;
        known_ = i;
        cv_.notify_all();
    }
    void reset(int64_t first_read)
    {
        std::lock_guard<std::mutex> lk(mu_);
        turn_ = 0; known_ = 0; reads_[0] = first_read; abort_ = false;
    }
    void abort()
    {
        std::lock_guard<std::mutex> lk(mu_);
        abort_ = true;
        cv_.notify_all();
    }

private:
    static constexpr int kRing = 256;   // far more than the spans in flight (<= 2 per GPU + 1)
    std::mutex mu_;
    std::condition_variable cv_;
    int64_t turn_ = 0, known_ = 0;
    int64_t reads_[kRing] = {0};
    bool abort_ = false;
};

// ------------------------------------------------------------------------------------------
// dense rows -> (bin, count) pairs on the device (--sparse, k >= 5): one warp per row
__global__ void row_nnz_kernel(const int32_t* __restrict__ rows, int64_t nrows, int bins, int64_t* __restrict__ nnz)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r <= nrows; r += warps) {
        if (r == nrows) { if (lane == 0) nnz[r] = 0; continue; }
        const int4* row = reinterpret_cast<const int4*>(rows + r * bins);
        int c = 0;
        for (int i = lane; i < bins / 4; i += 32) {
            const int4 v = row[i];
            c += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) nnz[r] = c;
    }
}

__global__ void row_compact_kernel(const int32_t* __restrict__ rows, int64_t nrows, int bins, const int64_t* __restrict__ begin,
                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ counts, int32_t* __restrict__ row_count)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const int32_t* row = rows + r * bins;
        int64_t at = begin[r];
        for (int i0 = 0; i0 < bins; i0 += 32) {
            const int v = row[i0 + lane];
            const uint32_t m = __ballot_sync(0xffffffffu, v != 0);
            if (v != 0) {
                const int64_t p = at + __popc(m & ((1u << lane) - 1u));
                keys[p] = (uint32_t)(i0 + lane);
                counts[p] = (uint32_t)v;
            }
            at += __popc(m);
        }
        if (lane == 0) row_count[r] = (int32_t)(at - begin[r]);
    }
}

// ------------------------------------------------------------------------------------------
// Bytes of the text of every row ("bin:count " tokens, format_row / format_pairs on the host), so that the writer
// knows where every row goes in the file before a byte is formatted.  One warp per row.
__device__ __forceinline__ int dec_digits(uint64_t v)
{
    int n = 1;
    uint64_t p = 10;
    while (v >= p && n < 20) { n++; p *= 10; }   // 20 digits: p would overflow, the loop has ended
    return n;
}
__device__ __forceinline__ int dec_digits32(uint32_t v)
{
    return 1 + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
           (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}
__global__ void row_text_bytes_kernel(const int32_t* __restrict__ rows, int64_t nrows, int bins, int sparse,
                                      int32_t* __restrict__ text_bytes)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const int32_t* row = rows + r * bins;
        int c = 0;
        for (int b = lane; b < bins; b += 32) {
            const int32_t v = row[b];
            const int tok = dec_digits32((uint32_t)b) + 2 + (v < 0 ? 1 + dec_digits32((uint32_t)(-(int64_t)v)) : dec_digits32((uint32_t)v));
            c += (sparse && v == 0) ? 0 : tok;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) text_bytes[r] = c;
    }
}
template <typename KeyT>
__global__ void pairs_text_bytes_kernel(const int64_t* __restrict__ row_begin, const int32_t* __restrict__ row_count,
                                        const KeyT* __restrict__ keys, const uint32_t* __restrict__ counts, int64_t nrows,
                                        int32_t* __restrict__ text_bytes)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
        const int64_t at = row_begin[r];
        const int n = row_count[r];
        int c = 0;
        for (int i = lane; i < n; i += 32) c += dec_digits((uint64_t)keys[at + i]) + 2 + dec_digits32(counts[at + i]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) text_bytes[r] = c;
    }
}

// ------------------------------------------------------------------------------------------
// GPU side of one worker: device input buffer + row ring.
struct Pipeline {
    static constexpr size_t kSlotBytes = (size_t)16 << 20;
    static size_t slot_bytes_limit()      // CFRK_ROW_SLOT_BYTES: size of a row ring slot (measurements)
    {
        static const size_t v = [] {
            const char* ev = getenv("CFRK_ROW_SLOT_BYTES");
            const long long x = ev ? atoll(ev) : 0;
            return x > 0 ? (size_t)x : kSlotBytes;
        }();
        return v;
    }
    int device = 0;
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
    int32_t* d_rows[2] = {}; int32_t* h_rows[2] = {}; size_t cap_rows = 0, cap_hrows = 0;
    // pooled scratch of the sparse outputs (device + pinned mirrors), grown on demand
    void* d_pool[5] = {}; size_t cap_pool[5] = {};     // [4]: text bytes of the rows
    void* h_pool[5] = {}; size_t cap_hpool[5] = {};
    char* h_hdr = nullptr; size_t cap_hhdr = 0;     // pinned staging of the record table
    Sequencer* seq = nullptr;
    bool fastq = false;                              // 4-line FASTQ records (fasta_scan.cu launch_fastq_scan)

    bool init(int dev, Sequencer* s, Err& err)
    {
        device = dev; seq = s;
        RF_CU(cudaSetDevice(dev));
        cfrk::keep_pool_memory(dev);
        RF_CU(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
        RF_CU(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            RF_CU(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
            RF_CU(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
        }
        return true;
    }
    bool dev_pool(int i, size_t bytes, Err& err)
    {
        if (bytes <= cap_pool[i]) return true;
        cudaFree(d_pool[i]); d_pool[i] = nullptr; cap_pool[i] = 0;
        const size_t want = bytes + bytes / 4 + 256;
        RF_CU(cudaMalloc(&d_pool[i], want));
        cap_pool[i] = want;
        return true;
    }
    bool host_pool(int i, size_t bytes, Err& err)
    {
        if (bytes <= cap_hpool[i]) return true;
        if (h_pool[i]) cudaFreeHost(h_pool[i]);
        h_pool[i] = nullptr; cap_hpool[i] = 0;
        const size_t want = bytes + bytes / 4 + 256;
        RF_CU(cudaMallocHost(&h_pool[i], want));
        cap_hpool[i] = want;
        return true;
    }
    bool reserve(size_t in_bytes, size_t nreads, size_t row_bytes, Err& err, bool host_rows = true)
    {
        if (in_bytes && in_bytes + CFRK_PAD > cap_in) {
            cudaFree(d_in); d_in = nullptr;
            cap_in = in_bytes + CFRK_PAD + in_bytes / 8;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_in), cap_in));
        }
        if (nreads > cap_reads) {
            cudaFree(d_start); cudaFree(d_length); cudaFree(d_header);
            d_start = nullptr; d_length = nullptr; d_header = nullptr;
            cap_reads = nreads + nreads / 4 + 64;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_start), cap_reads * 8));
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_length), cap_reads * 4));
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_header), cap_reads * 16));   // FASTQ scan: second half is scratch
        }
        const size_t need = row_bytes;
        if (row_bytes && need > cap_rows) {
            for (int i = 0; i < 2; i++) { cudaFree(d_rows[i]); d_rows[i] = nullptr; }
            cap_rows = 0;
            for (int i = 0; i < 2; i++) RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_rows[i]), need));
            cap_rows = need;
        }
        if (row_bytes && host_rows && need > cap_hrows) {      // compacted rows never reach the host as rows
            for (int i = 0; i < 2; i++) { if (h_rows[i]) cudaFreeHost(h_rows[i]); h_rows[i] = nullptr; }
            cap_hrows = 0;
            for (int i = 0; i < 2; i++) RF_CU(cudaMallocHost(reinterpret_cast<void**>(&h_rows[i]), need));
            cap_hrows = need;
        }
        return true;
    }
    // Span of raw file bytes -> HBM, record table built THERE (fasta_scan.cu); the host gets the
    // (small) table back for its bookkeeping and never looks at a base.  Spans end at a header
    // line or at the end of the input, so every record of the span is complete.
    bool upload_and_scan(const char* h_in, size_t in_bytes, RecordIndex& ri, Err& err)
    {
        ri.start.clear(); ri.length.clear(); ri.header.clear();
        n_headers = 0;
        if (in_bytes == 0) return true;
        if (!reserve(in_bytes, std::max<size_t>(cap_reads, in_bytes / 64 + 64), 0, err)) return false;
        RF_CU(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, compute));
        RF_CU(cudaMemsetAsync(d_in + in_bytes, 0, CFRK_PAD, compute));
        int64_t out[2];
        for (;;) {
            cudaError_t e = fastq ? cfrk::launch_fastq_scan(reinterpret_cast<const uint8_t*>(d_in), (int64_t)in_bytes, d_header, d_start,
                                                            d_length, (int64_t)cap_reads, out, compute)
                                  : cfrk::launch_fasta_scan(reinterpret_cast<const uint8_t*>(d_in), (int64_t)in_bytes, 1,
                                                            d_header, d_start, d_length, (int64_t)cap_reads, out, compute);
            if (e != cudaSuccess) { err.code = CFRK_ECUDA; err.msg = std::string("record scan: ") + cudaGetErrorString(e); return false; }
            if (out[1] != 4) break;
            if (!reserve(in_bytes, (size_t)out[0], 0, err)) return false;   // more records than guessed: grow, rescan
        }
        if (fastq && out[1] != 0) {
            err.code = CFRK_EFORMAT;
            err.msg = out[1] == 1 ? "FASTQ: a record does not begin with '@' (4 lines per record expected)"
                    : out[1] == 2 ? "FASTQ: the third line of a record does not begin with '+' (wrapped FASTQ is not supported)"
                    : out[1] == 3 ? "FASTQ: read longer than 2^31-1 bytes" : "FASTQ: the input does not end with a whole 4-line record";
            return false;
        }
        if (out[1] == 1) { err.code = CFRK_EFORMAT; err.msg = "'>' inside a sequence line (grep -c over-counts nS in the reference, src/fastaIO.h:16)"; return false; }
        if (out[1] == 2) { err.code = CFRK_EFORMAT; err.msg = "sequence text before the first '>' header (undefined in the reference, src/fastaIO.h:49-52)"; return false; }
        if (out[1] == 3) { err.code = CFRK_EFORMAT; err.msg = "record longer than 2^31-1 bytes (length is int in the reference, src/tipos.h:26)"; return false; }
        const size_t nh = (size_t)out[0];
        n_headers = nh;
        cur_bases = d_in; cur_start = d_start; cur_length = d_length;
        ri.header.resize(nh); ri.start.resize(nh); ri.length.resize(nh);
        if (nh) {
            // pinned staging: a pageable destination makes these copies synchronous and slow
            const size_t need = nh * 20;
            if (need > cap_hhdr) {
                if (h_hdr) cudaFreeHost(h_hdr);
                h_hdr = nullptr; cap_hhdr = 0;
                RF_CU(cudaMallocHost(reinterpret_cast<void**>(&h_hdr), need + need / 4));
                cap_hhdr = need + need / 4;
            }
            RF_CU(cudaMemcpyAsync(h_hdr, d_header, nh * 8, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaMemcpyAsync(h_hdr + nh * 8, d_start, nh * 8, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaMemcpyAsync(h_hdr + nh * 16, d_length, nh * 4, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaStreamSynchronize(compute));
            const int64_t* hh = reinterpret_cast<const int64_t*>(h_hdr);
            for (size_t i = 0; i < nh; i++) ri.header[i] = (size_t)hh[i];
            memcpy(ri.start.data(), h_hdr + nh * 8, nh * 8);
            memcpy(ri.length.data(), h_hdr + nh * 16, nh * 4);
        }
        return true;
    }
    // Exact mode: k-mers do not stop at line ends.  Replace the span by its unwrapped copy.
    bool unwrap_scanned(size_t in_bytes, size_t nreads, Err& err)
    {
        if (nreads == 0) return true;
        if (fastq) return true;     // a FASTQ read is one line: (start, length) already is the unwrapped read
        if (in_bytes + CFRK_PAD > cap_packed) {
            cudaFree(d_packed); d_packed = nullptr;
            cap_packed = in_bytes + CFRK_PAD + in_bytes / 8;
            RF_CU(cudaMalloc(reinterpret_cast<void**>(&d_packed), cap_packed));
        }
        if (nreads + 1 > cap_reads2) {
            cudaFree(d_start2); cudaFree(d_length2); d_start2 = nullptr; d_length2 = nullptr;
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
    bool count_scanned(const Span& sp, const RecordIndex& ri, size_t nrows, int k, int mode, bool sparse,
                       int64_t chunk_size, int64_t index_base, CfrkWriter& w, Err& err)
    {
        size_t in_bytes = sp.n;
        if (nrows > 0 && mode == CFRK_MODE_COMPAT && std::find(ri.length.begin(), ri.length.end(), 0) != ri.length.end()) {
            // An EMPTY read makes the reference walk over the bytes that FOLLOW it in its batch layout
            // (kmer_device.cuh read_extent).  Raw file bytes have header lines there, so for such a
            // (rare) span the records are first compacted into the reference layout: text, one
            // separator, next text, ...
            size_t total = 0;
            for (int32_t l : ri.length) total += (size_t)l + 1;
            std::vector<char> pack_buf(total + CFRK_PAD);
            std::vector<int64_t> pstart(ri.start.size());
            size_t wpos = 0;
            for (size_t i = 0; i < ri.start.size(); i++) {
                pstart[i] = (int64_t)wpos;
                memcpy(pack_buf.data() + wpos, sp.buf + ri.start[i], (size_t)ri.length[i]);
                wpos += (size_t)ri.length[i];
                pack_buf[wpos++] = '\n';
            }
            in_bytes = wpos;
            if (!reserve(in_bytes, ri.start.size(), 0, err)) return false;
            RF_CU(cudaMemcpyAsync(d_in, pack_buf.data(), in_bytes, cudaMemcpyHostToDevice, compute));
            RF_CU(cudaMemsetAsync(d_in + in_bytes, 0, CFRK_PAD, compute));
            RF_CU(cudaMemcpyAsync(d_start, pstart.data(), pstart.size() * 8, cudaMemcpyHostToDevice, compute));
            RF_CU(cudaStreamSynchronize(compute));   // pageable sources
            cur_bases = d_in; cur_start = d_start; cur_length = d_length;
        }
        if (nrows > 0 && mode == CFRK_MODE_EXACT && !unwrap_scanned(in_bytes, ri.start.size(), err)) return false;
        return count_rows(sp.index, in_bytes, ri.start.size(), nrows, k, mode, sparse, chunk_size, index_base, w, err);
    }
    // k > 8: sparse rows (exact semantics) of the span that upload_and_scan() left in HBM
    bool count_scanned_sparse(const Span& sp, const RecordIndex& ri, size_t nrows, int k, CfrkWriter& w, Err& err)
    {
        const size_t nreads = ri.start.size();
        const int32_t* tb = nullptr;     // text bytes per row, when the writer maps the file
        if (nrows > 0) {
            if (!unwrap_scanned(sp.n, nreads, err)) return false;   // k > 8 is exact mode only
            int64_t cap = 0;
            for (int32_t l : ri.length) cap += std::max(0, l + 1 - k + 1);   // unwrapped text <= raw length + 1
            cap = std::max<int64_t>(cap, 1);
            if (!dev_pool(0, (nreads + 1) * 8, err) || !dev_pool(1, nreads * 4, err) || !dev_pool(2, (size_t)cap * 8, err) ||
                !dev_pool(3, (size_t)cap * 4, err)) return false;
            int64_t total = 0;
            cudaError_t e = cfrk::launch_sparse(cur_bases, cfrk::FMT_ASCII, cur_start, cur_length, (int64_t)nreads, k,
                                                static_cast<int64_t*>(d_pool[0]), static_cast<int32_t*>(d_pool[1]), d_pool[2], 8,
                                                static_cast<uint32_t*>(d_pool[3]), cap, &total, compute);
            if (e != cudaSuccess) { err.code = CFRK_ECUDA; err.msg = std::string("sparse path: ") + cudaGetErrorString(e); return false; }
            if (!host_pool(0, (nrows + 1) * 8, err) || !host_pool(1, nrows * 4, err)) return false;
            RF_CU(cudaMemcpyAsync(h_pool[0], d_pool[0], (nrows + 1) * 8, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaMemcpyAsync(h_pool[1], d_pool[1], nrows * 4, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaStreamSynchronize(compute));
            const size_t used = (size_t)static_cast<int64_t*>(h_pool[0])[nrows];
            if (!host_pool(2, std::max<size_t>(used, 1) * 8, err) || !host_pool(3, std::max<size_t>(used, 1) * 4, err)) return false;
            RF_CU(cudaMemcpyAsync(h_pool[2], d_pool[2], used * 8, cudaMemcpyDeviceToHost, compute));
            RF_CU(cudaMemcpyAsync(h_pool[3], d_pool[3], used * 4, cudaMemcpyDeviceToHost, compute));
            if (w.sized()) {
                if (!dev_pool(4, nrows * 4, err) || !host_pool(4, nrows * 4, err)) return false;
                const unsigned g = (unsigned)std::min<size_t>((nrows + 7) / 8, 148 * 8);
                pairs_text_bytes_kernel<uint64_t><<<g, 256, 0, compute>>>(static_cast<int64_t*>(d_pool[0]), static_cast<int32_t*>(d_pool[1]),
                                                                          static_cast<uint64_t*>(d_pool[2]), static_cast<uint32_t*>(d_pool[3]),
                                                                          (int64_t)nrows, static_cast<int32_t*>(d_pool[4]));
                cfrk::count_launch();
                RF_CU(cudaGetLastError());
                RF_CU(cudaMemcpyAsync(h_pool[4], d_pool[4], nrows * 4, cudaMemcpyDeviceToHost, compute));
                tb = static_cast<int32_t*>(h_pool[4]);
            }
            RF_CU(cudaStreamSynchronize(compute));
        }
        if (!seq->wait_turn(sp.index)) { err.code = CFRK_EIO; err.msg = "aborted"; return false; }
        bool ok = true;
        if (nrows > 0)
            ok = w.write_pairs<uint64_t>(static_cast<int64_t*>(h_pool[0]), static_cast<int32_t*>(h_pool[1]),
                                         static_cast<uint64_t*>(h_pool[2]), static_cast<uint32_t*>(h_pool[3]), nrows, tb, err);
        seq->end_turn(sp.index);
        return ok;
    }
    // d_in / d_start / d_length hold the span: kernels slice by slice, rows through the ring to the
    // writer (entered when it is this span's turn; the kernels of the first slices run before that).
    bool count_rows(int64_t span_index, size_t in_bytes, size_t nreads, size_t nrows, int k, int mode, bool sparse,
                    int64_t chunk_size, int64_t index_base, CfrkWriter& w, Err& err)
    {
        bool have_turn = false;
        const bool ok = count_rows_inner(span_index, in_bytes, nreads, nrows, k, mode, sparse, chunk_size, index_base, w, err,
                                         have_turn);
        if (ok && !have_turn && !seq->wait_turn(span_index)) { err.code = CFRK_EIO; err.msg = "aborted"; return false; }
        if (ok) seq->end_turn(span_index);
        return ok;     // on failure the run is aborted by the caller: nobody waits for this turn
    }
    bool count_rows_inner(int64_t span_index, size_t in_bytes, size_t nreads, size_t nrows, int k, int mode, bool sparse,
                          int64_t chunk_size, int64_t index_base, CfrkWriter& w, Err& err, bool& have_turn)
    {
        if (nrows == 0) return true;
        const size_t bins = (size_t)1 << (2 * k), row_bytes = bins * 4;
        const bool compact = sparse && k >= 5;   // (bin, count) pairs instead of dense rows over PCIe
        // ring slots: as large as the span needs, at most kSlotBytes -- pinned host memory is what makes large slots
        // expensive; compacted rows stay on the device, and their slices end with a host round trip (the number of
        // pairs): large device slots, no host rows
        const size_t slot_limit = compact ? std::max(slot_bytes_limit(), (size_t)128 << 20) : slot_bytes_limit();
        if (!reserve(0, 0, std::max(row_bytes, std::min(slot_limit, nrows * row_bytes)), err, !compact)) return false;
        const size_t rpt = (size_t)cfrk::dense_reads_per_tile(k);
        size_t slice = std::max<size_t>(1, (compact ? cap_rows : std::min(cap_rows, cap_hrows)) / row_bytes);
        slice = std::max(rpt, slice / rpt * rpt);
        slice = std::min(slice, (nrows + rpt - 1) / rpt * rpt);
        const size_t nslices = (nrows + slice - 1) / slice;
        // compact slot layout: begin[slice+1] int64 | row_count[slice] int32 | keys | counts
        const size_t per_row = std::min<size_t>(bins, (size_t)cfrk::kRefBlockThreads + 2);
        const size_t tab_bytes = ((slice + 1) * 8 + slice * 4 + 15) & ~(size_t)15;
        const size_t pair_cap = slice * per_row;
        const size_t slot_bytes = tab_bytes + 2 * pair_cap * 4;
        size_t scan_bytes = 0;
        if (compact) {
            cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int64_t*)nullptr, (int64_t*)nullptr, (int64_t)slice + 1, compute);
            if (!dev_pool(0, 2 * slot_bytes, err) || !dev_pool(1, scan_bytes + 16, err) || !host_pool(0, 2 * slot_bytes, err)) return false;
        }
        const bool sized = w.sized();
        if (sized && (!dev_pool(4, 2 * slice * 4, err) || !host_pool(4, 2 * slice * 4, err))) return false;
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
                int32_t* d_tb = static_cast<int32_t*>(d_pool[4]) + (size_t)slot * slice;
                int32_t* h_tb = static_cast<int32_t*>(h_pool[4]) + (size_t)slot * slice;
                if (sized) {
                    const unsigned g = (unsigned)std::min<size_t>((r1 - r0 + 7) / 8, 148 * 8);
                    row_text_bytes_kernel<<<g, 256, 0, compute>>>(d_rows[slot], (int64_t)(r1 - r0), (int)bins, sparse ? 1 : 0, d_tb);
                    cfrk::count_launch();
                }
                if (compact) {
                    char* db = static_cast<char*>(d_pool[0]) + (size_t)slot * slot_bytes;
                    char* hb = static_cast<char*>(h_pool[0]) + (size_t)slot * slot_bytes;
                    int64_t* nnz = reinterpret_cast<int64_t*>(db);
                    int32_t* rcnt = reinterpret_cast<int32_t*>(db + (slice + 1) * 8);
                    uint32_t* ck = reinterpret_cast<uint32_t*>(db + tab_bytes);
                    uint32_t* cc = ck + pair_cap;
                    const int64_t n = (int64_t)(r1 - r0);
                    const unsigned g = (unsigned)std::min<int64_t>((n + 8) / 8 + 1, 148 * 8);
                    row_nnz_kernel<<<g, 256, 0, compute>>>(d_rows[slot], n, (int)bins, nnz);
                    cfrk::count_launch();
                    cub::DeviceScan::ExclusiveSum(d_pool[1], scan_bytes, nnz, nnz, n + 1, compute);
                    row_compact_kernel<<<g, 256, 0, compute>>>(d_rows[slot], n, (int)bins, nnz, ck, cc, rcnt);
                    cfrk::count_launch();
                    RF_CU(cudaGetLastError());
                    RF_CU(cudaEventRecord(done[slot], compute));
                    RF_CU(cudaStreamWaitEvent(copy, done[slot], 0));
                    // the pairs are dense in [0, total): the table first, then exactly the pairs
                    RF_CU(cudaMemcpyAsync(hb, db, tab_bytes, cudaMemcpyDeviceToHost, copy));
                    RF_CU(cudaStreamSynchronize(copy));
                    const size_t used = (size_t)reinterpret_cast<int64_t*>(hb)[n];
                    if (used > pair_cap) { err.code = CFRK_ECUDA; err.msg = "sparse compaction overflow"; return false; }
                    RF_CU(cudaMemcpyAsync(hb + tab_bytes, ck, used * 4, cudaMemcpyDeviceToHost, copy));
                    RF_CU(cudaMemcpyAsync(hb + tab_bytes + pair_cap * 4, cc, used * 4, cudaMemcpyDeviceToHost, copy));
                    if (sized) RF_CU(cudaMemcpyAsync(h_tb, d_tb, (r1 - r0) * 4, cudaMemcpyDeviceToHost, copy));
                    RF_CU(cudaEventRecord(drained[slot], copy));
                } else {
                    RF_CU(cudaEventRecord(done[slot], compute));
                    RF_CU(cudaStreamWaitEvent(copy, done[slot], 0));
                    RF_CU(cudaMemcpyAsync(h_rows[slot], d_rows[slot], (r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, copy));
                    if (sized) RF_CU(cudaMemcpyAsync(h_tb, d_tb, (r1 - r0) * 4, cudaMemcpyDeviceToHost, copy));
                    RF_CU(cudaEventRecord(drained[slot], copy));
                }
            }
            if (s >= 1) {  // format slice s-1 while slice s is on the GPU
                const size_t q = s - 1;
                const int slot = (int)(q & 1);
                const size_t r0 = q * slice, r1 = std::min(nrows, r0 + slice);
                if (!have_turn) {
                    if (!seq->wait_turn(span_index)) { err.code = CFRK_EIO; err.msg = "aborted"; return false; }
                    have_turn = true;
                }
                RF_CU(cudaEventSynchronize(drained[slot]));
                const int32_t* tb = sized ? static_cast<int32_t*>(h_pool[4]) + (size_t)slot * slice : nullptr;
                bool ok;
                if (compact) {
                    char* hb = static_cast<char*>(h_pool[0]) + (size_t)slot * slot_bytes;
                    ok = w.write_pairs<uint32_t>(reinterpret_cast<int64_t*>(hb), reinterpret_cast<int32_t*>(hb + (slice + 1) * 8),
                                                 reinterpret_cast<uint32_t*>(hb + tab_bytes),
                                                 reinterpret_cast<uint32_t*>(hb + tab_bytes + pair_cap * 4), r1 - r0, tb, err);
                } else {
                    ok = w.write_rows(h_rows[slot], r1 - r0, tb, err);
                }
                if (!ok) return false;
            }
        }
        RF_CU(cudaStreamSynchronize(compute));
        return true;
    }
    ~Pipeline()
    {
        cudaSetDevice(device);
        if (compute) cudaStreamSynchronize(compute);
        if (copy) cudaStreamSynchronize(copy);
        cudaFree(d_in); cudaFree(d_start); cudaFree(d_length); cudaFree(d_header);
        cudaFree(d_packed); cudaFree(d_start2); cudaFree(d_length2);
        for (int i = 0; i < 2; i++) {
            cudaFree(d_rows[i]);
            if (h_rows[i]) cudaFreeHost(h_rows[i]);
            if (done[i]) cudaEventDestroy(done[i]);
            if (drained[i]) cudaEventDestroy(drained[i]);
        }
        for (int i = 0; i < 5; i++) {
            cudaFree(d_pool[i]);
            if (h_pool[i]) cudaFreeHost(h_pool[i]);
        }
        if (h_hdr) cudaFreeHost(h_hdr);
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
        cudaGetLastError();
    }
};

// ------------------------------------------------------------------------------------------
struct RunCfg {
    int k, nt, mode, flags;
    int64_t chunk_size;
    std::vector<int> devices;
};

struct PassResult {
    int64_t reads = 0;          // records seen by the pass
    int64_t tail_off = -1;      // file offset of the header of the last read that opens a chunk
    int64_t tail_index = -1;
};

// One pass over the input from `file_off`: every span is uploaded and scanned; with `rows` its reads are
// counted and written.  first_read = global index of the first read of the pass.
bool run_pass(Source& src, size_t file_off, int64_t first_read, bool rows, const RunCfg& cfg, CfrkWriter* w, Trace& tr,
              PassResult& res, Err& err)
{
    if (!src.restart(file_off, err)) return false;
    // 16 MiB streaming window (measured on 2 M x 150 bp, k = 4, all rows: 64 MiB windows + 128 MiB row slots 1033 ms,
    // 32 + 32: 769 ms, 16 + 16: 506 ms -- pinning costs 0.35-0.7 ms/MiB and the first span's rows wait for it);
    // small inputs get small pinned buffers: test-sized inputs should start instantly
    size_t window = (size_t)16 << 20;
    if (const char* ev = getenv("CFRK_WINDOW_BYTES")) window = std::max<size_t>(4096, (size_t)atoll(ev));   // tests: many spans from small files
    const size_t left = src.size_hint() > file_off ? src.size_hint() - file_off : 0;
    if (left < window) window = std::max<size_t>((size_t)1 << 16, (left + 4095) & ~(size_t)4095);
    const int nworkers = (int)cfg.devices.size() * 2;
    const size_t spans_guess = left / window + 1;
    const int nslots = (int)std::min<size_t>((size_t)nworkers + 1, spans_guess + 1);
    SpanReader rd;
    Sequencer seq;
    seq.reset(first_read);
    // the workers start first: their streams, device buffers and pinned row ring are set up while the reader pins its
    // buffers and reads the first window
    std::mutex go_mu;
    std::condition_variable go_cv;
    int go = 0;     // 1: the reader runs, 2: it could not be opened

    std::mutex res_mu;
    Err first_err;
    std::atomic<bool> failed{false};
    auto fail = [&](const Err& e) {
        {
            std::lock_guard<std::mutex> lk(res_mu);
            if (!failed.exchange(true)) first_err = e;
        }
        seq.abort();
        rd.stop();
    };
    const int nthreads = (int)std::min<size_t>((size_t)nworkers, std::max<size_t>(1, spans_guess));
    std::vector<std::thread> th;
    // a worker that runs out of spans keeps its buffers until all are done: freeing device and pinned memory takes
    // the driver's locks for tens of milliseconds and stalls the worker that still has the last span
    // (measured, k = 6 --sparse: 760 ms for the last span instead of 35)
    std::mutex fin_mu;
    std::condition_variable fin_cv;
    int finished = 0;
    struct Finish {
        std::mutex& mu; std::condition_variable& cv; int& n; int all;
        ~Finish()
        {
            std::unique_lock<std::mutex> lk(mu);
            n++;
            cv.notify_all();
            cv.wait(lk, [&] { return n >= all; });
        }
    };
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([&, t] {
            Err e;
            Pipeline gpu;
            Finish fin{fin_mu, fin_cv, finished, nthreads};     // destroyed before gpu: waits for the other workers
            bool ready = gpu.init(cfg.devices[(size_t)t % cfg.devices.size()], &seq, e);
            if (ready && left >= ((size_t)1 << 20)) {
                const size_t row_bytes = rows && cfg.k <= CFRK_CLI_DENSE_MAX_K ? (size_t)4 << (2 * cfg.k) : 0;
                const bool compact = (cfg.flags & CFRK_RUN_SPARSE) && cfg.k >= 5;      // count_rows_inner sizes those itself
                ready = gpu.reserve(window + window / 8, window / 64 + 64,
                                    row_bytes && !compact ? std::max(row_bytes, Pipeline::slot_bytes_limit()) : 0, e);
            }
            {
                std::unique_lock<std::mutex> lk(go_mu);
                go_cv.wait(lk, [&] { return go != 0; });
                if (go == 2) return;
            }
            if (!ready) { fail(e); return; }
            gpu.fastq = src.is_fastq();
            RecordIndex ri;
            for (;;) {
                if (failed.load()) return;
                Span* sp = rd.next();
                if (!sp) return;
                bool ok = gpu.upload_and_scan(sp->buf, sp->n, ri, e);
                size_t nrows = 0;
                int64_t before = 0;
                if (ok) {
                    nrows = (size_t)(std::lower_bound(ri.header.begin(), ri.header.end(), sp->rows_end) - ri.header.begin());
                    ok = seq.get_reads_before(sp->index, &before);
                    if (ok) seq.set_reads_before(sp->index + 1, before + (int64_t)nrows);
                    else { e.code = CFRK_EIO; e.msg = "aborted"; }
                }
                tr.mark("span scanned (index, rows)", (size_t)sp->index, nrows);
                if (ok) {
                    std::lock_guard<std::mutex> lk(res_mu);
                    res.reads = std::max(res.reads, before + (int64_t)nrows - first_read);
                    // the last read of this span that opens a reference chunk
                    if (nrows > 0) {
                        const int64_t last = before + (int64_t)nrows - 1;
                        const int64_t opener = last - last % cfg.chunk_size;
                        if (opener >= before && opener > res.tail_index) {
                            res.tail_index = opener;
                            res.tail_off = (int64_t)(sp->file_off + ri.header[(size_t)(opener - before)]);
                        }
                    }
                }
                if (ok && rows) {
                    ok = cfg.k > CFRK_CLI_DENSE_MAX_K
                             ? gpu.count_scanned_sparse(*sp, ri, nrows, cfg.k, *w, e)
                             : gpu.count_scanned(*sp, ri, nrows, cfg.k, cfg.mode, (cfg.flags & CFRK_RUN_SPARSE) != 0, cfg.chunk_size, before, *w, e);
                    tr.mark("span rows written (index)", (size_t)sp->index);
                }
                rd.release(sp);
                if (!ok) { fail(e); return; }
            }
        });
    }
    const bool opened = rd.open(&src, window, std::max(2, nslots), err);
    if (opened) {
        tr.mark("reader open (pinned buffers, window)", (size_t)nslots, window);
        rd.start(file_off);
    }
    {
        std::lock_guard<std::mutex> lk(go_mu);
        go = opened ? 1 : 2;
    }
    go_cv.notify_all();
    for (auto& x : th) x.join();
    if (!opened) return false;
    rd.stop();
    if (failed.load()) { err = first_err; return false; }
    if (rd.error().code != CFRK_OK) { err = rd.error(); return false; }
    return true;
}

bool run_file(const char* fasta, const char* out_path, const RunCfg& cfg, Err& err)
{
    RF_CU(cudaSetDevice(cfg.devices[0]));
    RF_CU(cudaFree(nullptr));          // CUDA context creation (0.3-1 s per process) is outside the trace, as in round 1
    Trace tr;
    Source src;
    if (!src.open(fasta, err) || !src.detect_fastq(err)) return false;
    CfrkWriter w;
    if (!w.open(out_path, cfg.k, cfg.nt, cfg.flags & CFRK_RUN_SPARSE, err)) return false;
    tr.mark("opened");
    PassResult res;
    if (cfg.flags & CFRK_RUN_ALL_ROWS) return run_pass(src, 0, 0, true, cfg, &w, tr, res, err);
    // reference: only the remainder chunk reaches the file; nothing when nS % chunkSize == 0
    if (!run_pass(src, 0, 0, false, cfg, nullptr, tr, res, err)) return false;
    tr.mark("pass 1 done (reads, tail offset)", (size_t)res.reads, (size_t)std::max<int64_t>(res.tail_off, 0));
    const int64_t nS = res.reads;
    if (nS % cfg.chunk_size == 0 || res.tail_off < 0) return true;
    PassResult res2;
    return run_pass(src, (size_t)res.tail_off, res.tail_index, true, cfg, &w, tr, res2, err);
}

}  // namespace

extern "C" int cfrk_run_file_multi(const char* fasta_path, const char* out_path, int k, int nt, int64_t chunk_size,
                                   int flags, const int* devices, int n_devices)
{
    Err err;
    if (!fasta_path || !out_path) { err.code = CFRK_EINVAL; err.msg = "null path"; }
    else if (k < 1 || k > CFRK_SPARSE_MAX_K) { err.code = CFRK_EINVAL; err.msg = "k must be in 1..31"; }
    else if (k > CFRK_CLI_DENSE_MAX_K && (flags & (CFRK_RUN_SPARSE | CFRK_RUN_EXACT)) != (CFRK_RUN_SPARSE | CFRK_RUN_EXACT)) {
        err.code = CFRK_EINVAL;
        err.msg = "k > 8: rows have 4^k bins; pass --sparse --exact (non-zero bins only, intended semantics)";
    }
    else if (chunk_size <= 0) { err.code = CFRK_EINVAL; err.msg = "chunkSize must be positive"; }
    else if (n_devices < 1 || n_devices > 64 || !devices) { err.code = CFRK_EINVAL; err.msg = "need 1..64 devices"; }
    else {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
        RunCfg cfg;
        cfg.k = k; cfg.nt = nt; cfg.flags = flags; cfg.chunk_size = chunk_size;
        cfg.mode = (flags & CFRK_RUN_EXACT) ? CFRK_MODE_EXACT : CFRK_MODE_COMPAT;
        bool ok = true;
        for (int i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= ndev) ok = false;
            cfg.devices.push_back(devices[i]);
        }
        if (!ok) { err.code = CFRK_ECUDA; err.msg = "no such CUDA device (this library has no CPU fallback)"; }
        else run_file(fasta_path, out_path, cfg, err);
    }
    if (err.code != CFRK_OK) cfrk::set_last_error(err.msg);
    return err.code;
}

// PrintFreq (src/main.cu:26-63) as a function: dense rows -> .cfrk text.  Host only (no CUDA call).
extern "C" int cfrk_write_rows(const char* out_path, const int32_t* rows, int64_t n_rows, int k, int nt, int flags)
{
    Err err;
    if (!out_path || (!rows && n_rows > 0)) { err.code = CFRK_EINVAL; err.msg = "null argument"; }
    else if (k < 1 || k > CFRK_CLI_DENSE_MAX_K || n_rows < 0) { err.code = CFRK_EINVAL; err.msg = "k must be in 1..8, n_rows >= 0"; }
    else {
        CfrkWriter w;
        const size_t bins = (size_t)1 << (2 * k);
        // slices as the pipeline hands them over: at most kSlotBytes of rows each
        const size_t slice = std::max<size_t>(1, Pipeline::slot_bytes_limit() / (bins * 4));
        if (w.open(out_path, k, nt, flags & CFRK_RUN_SPARSE, err)) {
            std::vector<int32_t> tb;
            for (size_t r = 0; r < (size_t)n_rows && err.code == CFRK_OK; r += slice) {
                const size_t n = std::min(slice, (size_t)n_rows - r);
                const auto t0 = std::chrono::steady_clock::now();
                if (w.sized()) {      // the pipeline gets these from row_text_bytes_kernel
                    tb.resize(n);
                    const int nth = (int)std::min<size_t>((size_t)w.threads(), n);
                    w.pool().run(nth, [&](int t) {
                        for (size_t i = n * t / nth; i < n * (t + 1) / nth; i++)
                            tb[i] = row_text_bytes(rows + (r + i) * bins, bins, w.labels(), flags & CFRK_RUN_SPARSE);
                    });
                }
                const auto t1 = std::chrono::steady_clock::now();
                w.write_rows(rows + r * bins, n, w.sized() ? tb.data() : nullptr, err);
                if (getenv("CFRK_TRACE"))
                    fprintf(stderr, "[cfrk write_rows] %zu rows: text sizes %.1f ms, format + write %.1f ms\n", n,
                            std::chrono::duration<double, std::milli>(t1 - t0).count(),
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
            }
        }
    }
    if (err.code != CFRK_OK) cfrk::set_last_error(err.msg);
    return err.code;
}

extern "C" int cfrk_run_file(const char* fasta_path, const char* out_path, int k, int nt, int64_t chunk_size,
                             int flags, int device)
{
    return cfrk_run_file_multi(fasta_path, out_path, k, nt, chunk_size, flags, &device, 1);
}
