// fasta_scan.cu -- record boundary detection on the GPU.
//
// Replaces, for the file pipeline, popen("grep -c '>'") + the getline loop of the reference
// reader (src/fastaIO.h:12-69): the raw file bytes are already in HBM for the count kernel, so
// the record table is derived there and the host never looks at a base.
//   pass 1  every thread inspects 16 bytes: '>' at a line start = header; a '>' inside a header
//           line is text; a '>' inside a sequence line is where the reference is undefined (grep
//           over-counts nS) -> error flag.  Headers are counted per 4 KiB chunk.
//   scan    exclusive sum of the chunk counts (cub::DeviceScan: plumbing).
//   pass 2  header positions written in file order.
//   pass 3  one thread per record: walk to the end of the header line; the record text is
//           everything up to the next header (or the end of the span), and
//           length = max(0, text - 1)  (len = strlen - 1, src/fastaIO.h:53,65).
#include "kernels.h"

#include <cub/device/device_scan.cuh>

namespace cfrk {

extern void count_launch();

constexpr int kScanThreads = 256;
constexpr int kScanChunk = kScanThreads * 16;

__device__ __forceinline__ uint32_t header_mask16(const uint8_t* __restrict__ buf, int64_t n, int64_t p0, int* err)
{
    // bit j set <=> byte p0+j is '>' starting a line
    uint32_t m = 0;
    if (p0 >= n) return 0;
    const int cnt = (int)min((int64_t)16, n - p0);
    const uint4 v = *reinterpret_cast<const uint4*>(buf + p0);   // buffer is padded to 16 bytes
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint8_t prev = p0 == 0 ? (uint8_t)'\n' : buf[p0 - 1];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint8_t c = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
        if (j < cnt && c == '>') {
            if (prev == '\n') {
                m |= 1u << j;
            } else {
                // '>' inside a line.  Inside a HEADER line (">s1 A>G variant") the reference is well
                // defined: grep -c counts lines, and the parser only tests line[0] (src/fastaIO.h:16,
                // 49).  Inside a sequence line grep over-counts nS: undefined.  Walk back to the start
                // of the line (header lines are short; any other case is the error path).
                int64_t q = p0 + j - 1;
                while (q > 0 && buf[q - 1] != '\n') q--;
                if (buf[q] != '>') *err = 1;
            }
        }
        prev = c;
    }
    return m;
}

__global__ void __launch_bounds__(kScanThreads) scan_count_kernel(const uint8_t* __restrict__ buf, int64_t n,
                                                                 int64_t* __restrict__ chunk_count, int* __restrict__ err)
{
    const int64_t p0 = ((int64_t)blockIdx.x * kScanThreads + threadIdx.x) * 16;
    int e = 0;
    const int c = __popc(header_mask16(buf, n, p0, &e));
    if (e) *err = 1;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0 && buf[0] != '>') *err = 2;  // text before the first header
    __shared__ int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int wsum = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_sum, wsum);
    __syncthreads();
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = s_sum;
}

__global__ void __launch_bounds__(kScanThreads) scan_write_kernel(const uint8_t* __restrict__ buf, int64_t n,
                                                                 const int64_t* __restrict__ chunk_off,
                                                                 int64_t* __restrict__ header, int64_t cap)
{
    const int64_t p0 = ((int64_t)blockIdx.x * kScanThreads + threadIdx.x) * 16;
    int e = 0;
    const uint32_t m = header_mask16(buf, n, p0, &e);
    const int c = __popc(m);
    // exclusive scan of c over the CTA
    __shared__ int s_warp[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) before += w < warp ? s_warp[w] : 0;
    int64_t pos = chunk_off[blockIdx.x] + before + inc - c;
    uint32_t mm = m;
    while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        if (pos < cap) header[pos] = p0 + j;
        pos++;
    }
}

__global__ void records_kernel(const uint8_t* __restrict__ buf, int64_t n, const int64_t* __restrict__ header,
                               int64_t n_headers, int final_span, int64_t* __restrict__ start,
                               int32_t* __restrict__ length, int* __restrict__ err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t complete = final_span ? n_headers : n_headers - 1;
    if (i >= complete) return;
    const int64_t h = header[i];
    const int64_t end = i + 1 < n_headers ? header[i + 1] : n;
    int64_t s = h;
    while (s < end && buf[s] != '\n') s++;   // header lines are short
    s = s < end ? s + 1 : end;
    const int64_t text = end - s;
    if (text > 2147483647ll) { *err = 3; return; }
    start[i] = s;
    length[i] = text > 0 ? (int32_t)(text - 1) : 0;
}

// ------------------------------------------------------------------------------------------
// Exact mode at file level reads FASTA the intended way: line terminators are not bases and the
// last base is kept (the reference keeps the newlines and drops the last byte, src/fastaIO.h:49-67).
// One warp per record copies the record text without '\n' / '\r' into a packed buffer.
__global__ void unwrap_count_kernel(const uint8_t* __restrict__ buf, int64_t n, const int64_t* __restrict__ header,
                                    int64_t n_headers, const int64_t* __restrict__ start, int64_t nrec,
                                    int64_t* __restrict__ kept)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r <= nrec; r += warps) {
        if (r == nrec) { if (lane == 0) kept[r] = 0; continue; }
        const int64_t s = start[r], e = r + 1 < n_headers ? header[r + 1] : n;
        int c = 0;
        for (int64_t i = s + lane; i < e; i += 32) {
            const uint8_t b = buf[i];
            c += (b != '\n' && b != '\r');
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) kept[r] = c;
    }
}

__global__ void unwrap_copy_kernel(const uint8_t* __restrict__ buf, int64_t n, const int64_t* __restrict__ header,
                                   int64_t n_headers, const int64_t* __restrict__ start, int64_t nrec,
                                   const int64_t* __restrict__ new_start, uint8_t* __restrict__ out,
                                   int32_t* __restrict__ new_length)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrec; r += warps) {
        const int64_t s = start[r], e = r + 1 < n_headers ? header[r + 1] : n;
        uint8_t* dst = out + new_start[r];
        int64_t base = 0;
        for (int64_t i0 = s; i0 < e; i0 += 32) {
            const int64_t i = i0 + lane;
            const uint8_t b = i < e ? buf[i] : (uint8_t)'\n';
            const bool keep = b != '\n' && b != '\r';
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) dst[base + __popc(m & ((1u << lane) - 1u))] = b;
            base += __popc(m);
        }
        if (lane == 0) new_length[r] = (int32_t)base;
    }
}

// records [0, nrec) of a scanned span -> packed bases + (new_start, new_length); d_new_start has
// nrec + 1 entries (the last = packed size).  Asynchronous on st.
cudaError_t launch_unwrap(const uint8_t* d_buf, int64_t n, const int64_t* d_header, int64_t n_headers,
                          const int64_t* d_start, int64_t nrec, uint8_t* d_out, int64_t* d_new_start,
                          int32_t* d_new_length, cudaStream_t st)
{
    if (nrec <= 0) return cudaSuccess;
    const int64_t blocks = (nrec + 1 + 7) / 8;
    const unsigned grid = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
    unwrap_count_kernel<<<grid, 256, 0, st>>>(d_buf, n, d_header, n_headers, d_start, nrec, d_new_start);
    count_launch();
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_new_start, d_new_start, nrec + 1, st);
    if (e != cudaSuccess) return e;
    void* tmp = nullptr;
    if ((e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st)) != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_new_start, d_new_start, nrec + 1, st);
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return e;
    unwrap_copy_kernel<<<grid, 256, 0, st>>>(d_buf, n, d_header, n_headers, d_start, nrec, d_new_start, d_out, d_new_length);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// FASTQ (4 lines per record: '@' header, sequence, '+' line, qualities; the reference cannot read it --
// it looks for '>' -- but the SRR datasets of swift/roda.sh are distributed in this form).  Record i =
// lines 4i .. 4i+3 of the span; the read is line 4i+1 without its line terminator, so (start, length)
// into the raw bytes again, and the '\n' behind it is the separator the kernels expect.
//   pass 1  newlines per 4 KiB chunk;  scan;  pass 2  one thread per 16 bytes: for every newline, its
//           line number decides what begins behind it (4i: header '@', 4i+1: read, 4i+2: '+').
__global__ void __launch_bounds__(kScanThreads) fastq_count_kernel(const uint8_t* __restrict__ buf, int64_t n,
                                                                  int64_t* __restrict__ chunk_count)
{
    const int64_t p0 = ((int64_t)blockIdx.x * kScanThreads + threadIdx.x) * 16;
    int c = 0;
    if (p0 < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(buf + p0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int cnt = (int)min((int64_t)16, n - p0);
#pragma unroll
        for (int j = 0; j < 16; j++) c += (j < cnt && (uint8_t)(w[j >> 2] >> (8 * (j & 3))) == '\n');
    }
    __shared__ int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int wsum = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_sum, wsum);
    __syncthreads();
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = s_sum;
}

// line L (0-based) begins at byte 0 (L = 0) or behind the L-th newline.  header[i] = begin of line 4i,
// start[i] = begin of line 4i+1, length[i] = (begin of line 4i+2) - 1 - start[i]; err: line 4i does not
// begin with '@' (1), line 4i+2 does not begin with '+' (2).
__global__ void __launch_bounds__(kScanThreads) fastq_write_kernel(const uint8_t* __restrict__ buf, int64_t n,
                                                                  const int64_t* __restrict__ chunk_off,
                                                                  int64_t* __restrict__ header, int64_t* __restrict__ start,
                                                                  int32_t* __restrict__ length, int64_t cap, int* __restrict__ err)
{
    const int64_t p0 = ((int64_t)blockIdx.x * kScanThreads + threadIdx.x) * 16;
    uint32_t m = 0;
    if (p0 < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(buf + p0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int cnt = (int)min((int64_t)16, n - p0);
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (j < cnt && (uint8_t)(w[j >> 2] >> (8 * (j & 3))) == '\n') m |= 1u << j;
    }
    const int c = __popc(m);
    __shared__ int s_warp[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) before += w < warp ? s_warp[w] : 0;
    int64_t nl = chunk_off[blockIdx.x] + before + inc - c;   // newlines before this thread's bytes
    if (p0 == 0 && n > 0) {                                  // line 0 begins the span
        if (buf[0] != '@') *err = 1;
        if (cap > 0) header[0] = 0;
    }
    uint32_t mm = m;
    while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int64_t line = nl + 1;                         // the line that begins behind this newline
        const int64_t q = p0 + j + 1, rec = line >> 2;
        nl++;
        if (rec >= cap) continue;
        switch ((int)(line & 3)) {
        case 0: if (q < n) { header[rec] = q; if (buf[q] != '@') *err = 1; } break;
        case 1: start[rec] = q; break;
        case 2: {
            // the read ended one byte before q: its '\n' (a '\r' before it stays a base, like CRLF FASTA)
            // start[rec] is written by another thread: store the END here, the length is fixed up below
            length[rec] = (int32_t)0;
            header[cap + rec] = q - 1;                       // scratch: end of the read (second half of header[])
            if (q < n && buf[q] != '+') *err = 2;
            break;
        }
        default: break;
        }
    }
}

__global__ void fastq_length_kernel(const int64_t* __restrict__ header, const int64_t* __restrict__ start, int64_t nrec,
                                    int64_t cap, int32_t* __restrict__ length, int* __restrict__ err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    const int64_t len = header[cap + i] - start[i];
    if (len > 2147483647ll) { *err = 3; return; }
    length[i] = (int32_t)(len > 0 ? len : 0);
}

// FASTQ span (begins with a '@' header line, ends behind a whole record or at the end of the input) ->
// header / start / length of its records.  d_header must hold 2 * cap entries (the second half is scratch).
// h_out[0] = records, h_out[1] = error (0 ok, 1 a record does not begin with '@', 2 no '+' line, 3 read too
// long, 4 more records than cap, 5 the span does not end with a whole record).  Synchronises the stream.
cudaError_t launch_fastq_scan(const uint8_t* d_buf, int64_t n, int64_t* d_header, int64_t* d_start, int32_t* d_length,
                              int64_t cap, int64_t* h_out, cudaStream_t st)
{
    h_out[0] = 0; h_out[1] = 0;
    if (n <= 0) return cudaSuccess;
    cudaError_t e;
    const int64_t nchunks = (n + kScanChunk - 1) / kScanChunk;
    int64_t* chunk = nullptr;
    int* d_err = nullptr;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&chunk), (size_t)(nchunks + 1) * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&d_err), 4, st)) != cudaSuccess) return e;
    cudaMemsetAsync(d_err, 0, 4, st);
    cudaMemsetAsync(chunk + nchunks, 0, 8, st);
    fastq_count_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(d_buf, n, chunk);
    count_launch();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, chunk, chunk, nchunks + 1, st);
    void* tmp = nullptr;
    if ((e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st)) != cudaSuccess) return e;
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, chunk, chunk, nchunks + 1, st)) != cudaSuccess) return e;
    int64_t newlines = 0;
    uint8_t last = 0;
    cudaMemcpyAsync(&newlines, chunk + nchunks, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last, d_buf + n - 1, 1, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    const int64_t lines = newlines + (last == '\n' ? 0 : 1);      // a last line without '\n' still counts
    int err = 0;
    const int64_t nrec = lines / 4;
    if (lines % 4 != 0) {
        err = 5;
    } else if (nrec > cap) {
        err = 4;
    } else if (nrec > 0) {
        fastq_write_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(d_buf, n, chunk, d_header, d_start, d_length, cap, d_err);
        count_launch();
        if (last != '\n') {
            // the quality line of the last record has no terminator: nothing begins behind it; fine
        }
        fastq_length_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(d_header, d_start, nrec, cap, d_length, d_err);
        count_launch();
        cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    }
    cudaFreeAsync(tmp, st);
    cudaFreeAsync(chunk, st);
    cudaFreeAsync(d_err, st);
    h_out[0] = err == 4 ? nrec : (err ? 0 : nrec);
    h_out[1] = err;
    return cudaGetLastError();
}

// d_header: cap entries; d_start/d_length: cap entries; h_out[0] = number of headers in the span,
// h_out[1] = error (0 ok, 1 '>' inside a sequence line, 2 text before the first header, 3 record too long,
// 4 more headers than cap).  Synchronises the stream.
cudaError_t launch_fasta_scan(const uint8_t* d_buf, int64_t n, int final_span, int64_t* d_header, int64_t* d_start,
                              int32_t* d_length, int64_t cap, int64_t* h_out, cudaStream_t st)
{
    h_out[0] = 0; h_out[1] = 0;
    if (n <= 0) return cudaSuccess;
    cudaError_t e;
    const int64_t nchunks = (n + kScanChunk - 1) / kScanChunk;
    int64_t* chunk = nullptr;
    int* d_err = nullptr;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&chunk), (size_t)(nchunks + 1) * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&d_err), 4, st)) != cudaSuccess) return e;
    cudaMemsetAsync(d_err, 0, 4, st);
    cudaMemsetAsync(chunk + nchunks, 0, 8, st);
    scan_count_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(d_buf, n, chunk, d_err);
    count_launch();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, chunk, chunk, nchunks + 1, st);
    void* tmp = nullptr;
    if ((e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st)) != cudaSuccess) return e;
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, chunk, chunk, nchunks + 1, st)) != cudaSuccess) return e;
    int64_t n_headers = 0;
    int err = 0;
    cudaMemcpyAsync(&n_headers, chunk + nchunks, 8, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (n_headers > cap) {
        err = 4;
    } else {
        scan_write_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(d_buf, n, chunk, d_header, cap);
        count_launch();
        if (n_headers > 0) {
            records_kernel<<<(unsigned)((n_headers + 255) / 256), 256, 0, st>>>(d_buf, n, d_header, n_headers, final_span,
                                                                                d_start, d_length, d_err);
            count_launch();
        }
        cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    }
    cudaFreeAsync(tmp, st);
    cudaFreeAsync(chunk, st);
    cudaFreeAsync(d_err, st);
    h_out[0] = n_headers;
    h_out[1] = err;
    return cudaGetLastError();
}

}  // namespace cfrk
