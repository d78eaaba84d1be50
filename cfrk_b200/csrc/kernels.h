// kernels.h -- launchers of the sm_100a kernels (kernels.cu); internal to libcfrk_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cfrk {

// Rows [read_begin, read_end) of a batch of nS reads -> out[(i-read_begin)*4^k ..]; k in 1..8.
// compat: read i opens a reference chunk (its spill is dropped) iff i == 0 when chunk_size == 0,
// else iff (index_base + i) % chunk_size == 0.
cudaError_t launch_dense(const void* bases, int fmt, const int64_t* start, const int32_t* length,
                         int64_t nN, int64_t nS, int64_t read_begin, int64_t read_end, int k, int mode,
                         int64_t chunk_size, int64_t index_base, int32_t* out, cudaStream_t st,
                         const uint16_t* packed_valid = nullptr);   // fmt 2 (packed): bases = uint32 codes
int dense_reads_per_tile(int k);

// hist[4^k] += exact-mode k-mer counts of the whole batch (nN bytes of bases); k in 1..15.
// k = 9..13 on large batches goes through the partitioned path (hist_split.cu).
cudaError_t launch_global_hist(const void* bases, int fmt, const int64_t* start, const int32_t* length,
                               int64_t nN, int64_t nS, int k, uint32_t* hist, cudaStream_t st);
bool hist_split_applies(int k, int64_t nN);
void keep_pool_memory(int dev);   // scratch of the stream-ordered pool stays mapped between calls (sparse.cu)
cudaError_t launch_global_hist_split(const void* bases, int fmt, const int64_t* start, const int32_t* length, int64_t nN,
                                     int64_t nS, int k, uint32_t* hist, cudaStream_t st);

// n bytes of bases -> ceil(n/16) words of 2-bit codes + 16-bit validity masks.
cudaError_t launch_encode_2bit(const void* bases, int fmt, int64_t n, uint32_t* codes, uint16_t* valid,
                               cudaStream_t st);

// Sparse per-read rows (exact semantics), k in 1..31; key_bytes 4 (k <= 16) or 8.  Synchronises
// `st` internally (needs the total window count on the host).  cudaErrorInvalidValue = capacity.
cudaError_t launch_sparse(const void* bases, int fmt, const int64_t* start, const int32_t* length, int64_t nS, int k,
                          int64_t* row_begin, int32_t* row_count, void* keys, int key_bytes, uint32_t* counts,
                          int64_t capacity, int64_t* total_windows, cudaStream_t st,
                          const uint16_t* packed_valid = nullptr);   // fmt 2 (packed): bases = uint32 codes

// FASTA record table of a span of raw file bytes (16-byte aligned, padded): header positions,
// and (start, length) of every record whose end is known (all when final_span, else all but the
// last).  h_out[0] = headers found, h_out[1] = error code (fasta_scan.cu).  Synchronises `st`.
cudaError_t launch_fasta_scan(const uint8_t* d_buf, int64_t n, int final_span, int64_t* d_header, int64_t* d_start,
                              int32_t* d_length, int64_t cap, int64_t* h_out, cudaStream_t st);

// The same for a 4-line FASTQ span; d_header holds 2 * cap entries.  h_out[1]: 0 ok, 1 no '@', 2 no '+',
// 3 read too long, 4 cap too small, 5 the span does not end with a whole record.
cudaError_t launch_fastq_scan(const uint8_t* d_buf, int64_t n, int64_t* d_header, int64_t* d_start, int32_t* d_length,
                              int64_t cap, int64_t* h_out, cudaStream_t st);

// Exact mode at file level: record text without line terminators, last base kept (fasta_scan.cu).
cudaError_t launch_unwrap(const uint8_t* d_buf, int64_t n, const int64_t* d_header, int64_t n_headers,
                          const int64_t* d_start, int64_t nrec, uint8_t* d_out, int64_t* d_new_start,
                          int32_t* d_new_length, cudaStream_t st);

// Dense rows in HBM -> (bin, count) pairs of their non-zero bins, row r at keys/counts[off[r] ..] (row_pairs.cu):
// what the host-buffer operator ships over PCIe for the rows that host threads expand.
cudaError_t launch_rows_to_pairs(const int32_t* rows, int64_t nrows, int bins, const int64_t* off, uint32_t* keys,
                                 uint32_t* counts, int32_t* row_count, cudaStream_t st);

// three word-wise copies into mapped pinned host memory, done by SMs (the copy engine is busy with dense rows)
cudaError_t launch_pairs_to_host(const uint32_t* s0, uint32_t* d0, int64_t n0, const uint32_t* s1, uint32_t* d1, int64_t n1,
                                 const uint32_t* s2, uint32_t* d2, int64_t n2, cudaStream_t st);

uint64_t launch_count();

}  // namespace cfrk
