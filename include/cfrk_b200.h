/*
 * cfrk_b200.h -- C ABI of the B200-native CFRK hot path (libcfrk_b200.so).
 *
 * Drop-in boundary for the reference's in-process operator
 *     void kmer_main(struct read *rd, lint nN, lint nS, int k, ushort device);
 *                                                   (reference src/kmer.cuh:6,
 *                                                    src/kmer_main.cu:20-128)
 * and for the stages either side of it (reader src/fastaIO.h:24-148, writer
 * src/main.cu:26-62, chunk/tail driver src/main.cu:232-305).
 *
 * Plain pointers and sizes only; no C++ or torch types.  Every function returns
 * CFRK_OK (0) or a negative CFRK_E* code; cfrk_last_error() gives the text for
 * the calling thread.  There is no CPU fallback: without a usable CUDA device
 * every compute entry point fails with CFRK_ECUDA.
 *
 * INTEGRATION.md shows the two-line change that makes the reference's own
 * main.cu call this library.
 */
#ifndef CFRK_B200_H
#define CFRK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFRK_OK        0
#define CFRK_EINVAL   -1  /* bad argument (k out of range, null pointer, ...)   */
#define CFRK_ECUDA    -2  /* CUDA runtime error, text in cfrk_last_error()      */
#define CFRK_ENOMEM   -3  /* host or device allocation failed                   */
#define CFRK_EIO      -4  /* cannot open / read / write a file                  */
#define CFRK_EFORMAT  -5  /* FASTA input on which the reference is undefined    */

/* semantics */
#define CFRK_MODE_COMPAT 0 /* bit-exact with the reference, quirks included:
                              positions t < min(len-1,1024) are visited, a window holding
                              a non-ACGT byte or the terminator adds 1 to the LAST bin of
                              the PREVIOUS read (src/kmer_kernel.cu:83-88)              */
#define CFRK_MODE_EXACT  1 /* every one of the len-k+1 windows; invalid windows skipped */

/* layout of the bases buffer */
#define CFRK_FMT_CODES 0   /* reference layout (src/tipos.h:23-30, src/fastaIO.h:123-139):
                              one int8 per base, A/C/G/T = 0/1/2/3, anything else -1, one
                              -1 terminator after each read                             */
#define CFRK_FMT_ASCII 1   /* FASTA letters as read from the file (upper or lower case);
                              one separator byte (any non-ACGT byte) after each read    */

#define CFRK_DENSE_MAX_K 12  /* dense rows, 4^k int32 per read (the reference's float32 index
                                is exact up to k = 12); k <= 8 is the tuned range */
#define CFRK_CLI_DENSE_MAX_K 8 /* dense TEXT rows of the cfrk command                    */
#define CFRK_HIST_MAX_K  15  /* whole-dataset histogram, 4^k uint32 in HBM              */
#define CFRK_SPARSE_MAX_K 31 /* sparse per-read rows, uint64 keys                       */
#define CFRK_PAD         16  /* device bases buffers must be readable up to the next
                                16-byte boundary and 16-byte aligned                    */

const char *cfrk_version(void);
const char *cfrk_last_error(void);
int         cfrk_device_count(void);

/* Number of kernels this library has launched since load (all threads). */
uint64_t    cfrk_launch_count(void);

/*
 * Cached resources.  Per CONCURRENT calling thread of cfrk_count_dense_host / kmer_main the library
 * keeps one context: the device copy of the largest batch seen, up to 2 x 256 MiB of row ring,
 * 2 streams, 4 events (a thread that exits hands its context to the next new thread, so thread
 * churn does not add up).  Per stream: a few KiB of launch scratch.  cfrk_release() frees the idle
 * contexts, the calling thread's own, the scratch and the cached pinned buffers; call it when no
 * other thread is inside the library.  Nothing is freed at process exit.
 */
int         cfrk_release(void);

/*
 * Give back a rows buffer that kmer_main allocated (rd->Freq).  The reference pins a fresh buffer
 * per call and never frees it (src/kmer_main.cu:115); here the buffer returns to a pinned arena
 * and the next call of that size reuses it.  Optional: a caller that never frees behaves exactly
 * like the reference's.  Any other pinned pointer is passed to cudaFreeHost.
 */
void        cfrk_free_host(void *p);

/* ---- the operator: replaces kmer_main (src/kmer_main.cu:20-128) --------------------- */

/*
 * Host buffers in, host rows out; synchronous, like kmer_main.
 *   bases[nN], start[nS] (byte offset of each read in bases), length[nS]: caller-owned,
 *   pinned or pageable.  freq_out[nS * 4^k]: caller-owned (pinned memory makes the
 *   device->host copy asynchronous and full speed).
 * Replaces: the 5 cudaMalloc + 3 H2D + 4 launches + cudaMallocHost + D2H + 5 cudaFree of
 * src/kmer_main.cu:59-124.  Unlike the reference it returns errors instead of printing
 * them, uses 64-bit indexing throughout (no nS*4^k < 2^31 limit) and never
 * writes outside freq_out (the reference stores Freq[-1] for read 0).
 */
int cfrk_count_dense_host(const void *bases, int fmt, const int64_t *start, const int32_t *length,
                          int64_t nN, int64_t nS, int k, int mode, int device,
                          int32_t *freq_out);

/*
 * Host threads of cfrk_count_dense_host / kmer_main (the reference's `nt`, src/main.cu:235, is parsed
 * and unused: its OpenMP pragmas are compiled without -fopenmp).  For rows of >= 4 KiB the operator
 * splits a batch: one share of the rows is written into the caller's buffer by the GPU's DMA engine,
 * the rest is counted on the GPU as well, compacted there to (bin, count) pairs, and n host threads expand
 * the pairs into the rows (zeros + counts, streaming stores).  Nothing is counted on the host.  n = 0: DMA only;
 * n < 0: back to the default.
 * Default: min(hardware threads, 16), or the environment variable CFRK_HOST_THREADS.
 */
void cfrk_set_host_threads(int n);

/*
 * Device-resident variant: every pointer is device memory on `device`, `stream` is a
 * cudaStream_t (NULL = the legacy default stream).  Asynchronous.  d_bases must honour
 * CFRK_PAD; d_freq must be 16-byte aligned.  Reads [read_begin, read_end) of the batch
 * are counted into d_freq[(i-read_begin)*4^k ...] (any range; row tiles are laid out from
 * read_begin).  Splitting a batch into ranges -- row rings, multi-GPU shards -- gives exactly
 * the rows of one call: the last tile of a range scans the first read of the next one for its
 * spill.  cfrk_dense_reads_per_tile(k) is a good granularity for range sizes.
 * compat mode: the reference drops the spill of the first read of every kmer_main call
 * (one call per chunk, src/main.cu:222,294,300).  chunk_size == 0: the batch is one such
 * call.  chunk_size > 0: read i opens a chunk iff (first_read_index + i) % chunk_size == 0,
 * so one launch reproduces many reference calls.
 */
int cfrk_count_dense_device(const void *d_bases, int fmt, const int64_t *d_start,
                            const int32_t *d_length, int64_t nN, int64_t nS,
                            int64_t read_begin, int64_t read_end, int k, int mode,
                            int64_t chunk_size, int64_t first_read_index,
                            int32_t *d_freq, void *stream);
int cfrk_dense_reads_per_tile(int k);

/*
 * The same operator on PACKED reads: d_codes / d_valid are what cfrk_encode_2bit_device wrote
 * (16 bases per uint32, first base in the top bits; validity bit 15-j for base 16w+j), start[]
 * and length[] still count bases.  This is the batch layout that replaces the reference's
 * 1-byte-per-base `struct read` (src/tipos.h:23-30): encode once, count for any number of k.
 */
int cfrk_count_dense_packed_device(const uint32_t *d_codes, const uint16_t *d_valid,
                                   const int64_t *d_start, const int32_t *d_length, int64_t nN,
                                   int64_t nS, int64_t read_begin, int64_t read_end, int k,
                                   int mode, int64_t chunk_size, int64_t first_read_index,
                                   int32_t *d_freq, void *stream);

/* ---- stages of the north-star pipeline exposed on their own ------------------------- */

/*
 * Base -> 2-bit encode with ambiguous-base masking (replaces the per-base switch of
 * src/fastaIO.h:123-139 and the int8 batch layout of src/tipos.h:23-30).
 * 16 bases per uint32 word, first base in the two most significant bits; valid[w] bit
 * (15-j) is set iff base 16w+j is A/C/G/T.  n = number of bytes; d_codes has ceil(n/16)
 * words, d_valid ceil(n/16) uint16.
 */
int cfrk_encode_2bit_device(const void *d_bases, int fmt, int64_t n, uint32_t *d_codes,
                            uint16_t *d_valid, void *stream);

/*
 * Whole-dataset k-mer histogram (exact semantics), accumulated INTO d_hist[4^k] uint32
 * (caller zeroes it; several calls / several GPUs add up, reduce across ranks with NCCL).
 */
int cfrk_global_hist_device(const void *d_bases, int fmt, const int64_t *d_start,
                            const int32_t *d_length, int64_t nN, int64_t nS, int k,
                            uint32_t *d_hist, void *stream);

/*
 * The exchange step of the whole-dataset histogram on the GPUs of ONE box (north_star config 5; the reference has
 * no multi-GPU reduction -- its devCount pthreads share one GPU, src/main.cu:208-230): every rank's table
 * becomes the element-wise sum of all tables, in place, by one kernel per GPU over NVLink peer memory (two shots:
 * rank r sums slice r of every table, then writes it into every table).  To be called by every rank with the same
 * epoch (1, 2, 3, ... per call), on the stream that produced its table.
 *   peer_tables[i]  address of rank i's 4^k uint32 table as THIS rank addresses it (own table included) --
 *                   symmetric / peer-mapped allocations made by the host side (torch symmetric memory, cuMem + IPC)
 *   peer_flags[i]   the same for a CFRK_HIST_REDUCE_FLAG_BYTES buffer per rank, zeroed once before the first call;
 *                   word CFRK_HIST_REDUCE_STATUS_WORD of the own buffer is non-zero if a peer did not show up
 *                   within ~2 s (tables then undefined)
 *   multicast_table NULL, or the multicast address of the tables (NVLS: the switch adds; sums must stay < 2^32)
 */
#define CFRK_HIST_REDUCE_MAX_RANKS   8
#define CFRK_HIST_REDUCE_FLAG_BYTES  65536
#define CFRK_HIST_REDUCE_STATUS_WORD 8192
int cfrk_hist_allreduce_device(void *const *peer_tables, void *const *peer_flags, int rank, int world,
                               int64_t n_bins, uint32_t epoch, void *multicast_table, void *stream);

/*
 * Sparse per-read counts, exact semantics, k = 1..31 (BASELINE configs 3 and 4; no reference
 * counterpart: its dense rows end at k = 8 with the default chunk).
 * Row r = the sorted distinct k-mers of read r and their multiplicities, stored at
 *   d_keys/d_counts[d_row_begin[r] .. d_row_begin[r] + d_row_count[r])
 * with d_row_begin[r] = sum_{j<r} max(0, length[j]-k+1) computed by the call (nS+1 entries; the
 * last one is the total number of windows, also returned in *total_windows).  capacity = number
 * of pairs d_keys/d_counts can hold; CFRK_EINVAL if the total exceeds it.  key_bytes = 4
 * (uint32 keys, k <= 16) or 8 (uint64 keys).  Key = sum code_i * 4^(k-1-i), like the dense bin.
 * Synchronises the stream (it needs the total on the host) and allocates scratch with
 * cudaMallocAsync.
 */
int cfrk_count_sparse_device(const void *d_bases, int fmt, const int64_t *d_start,
                             const int32_t *d_length, int64_t nN, int64_t nS, int k, int key_bytes,
                             int64_t *d_row_begin, int32_t *d_row_count, void *d_keys,
                             uint32_t *d_counts, int64_t capacity, int64_t *total_windows,
                             void *stream);

/* The same on PACKED reads (the output of cfrk_encode_2bit_device: encode once, count for any number of k). */
int cfrk_count_sparse_packed_device(const uint32_t *d_codes, const uint16_t *d_valid, const int64_t *d_start,
                                    const int32_t *d_length, int64_t nN, int64_t nS, int k, int key_bytes,
                                    int64_t *d_row_begin, int32_t *d_row_count, void *d_keys,
                                    uint32_t *d_counts, int64_t capacity, int64_t *total_windows,
                                    void *stream);

/*
 * FASTA record table on the GPU (replaces popen("grep -c") + the getline loop of
 * src/fastaIO.h:12-69 for bytes that are already in HBM).  d_bytes: n raw file bytes, 16-byte
 * aligned and readable to the next 16-byte boundary; the span must begin with a header line.
 * Writes the position of every header's '>' to d_header, and (start, length) -- reference
 * semantics: text = all lines after the header up to the next header, length = text - 1 -- of
 * every record whose end is known: all of them if is_final, else all but the last.
 * *n_headers = headers found (records described = n_headers or n_headers - 1).
 * CFRK_EFORMAT where the reference is undefined ('>' inside a sequence line, text before the
 * first header; a '>' inside a header line is plain text, as in the reference), CFRK_EINVAL if capacity is too small.  Synchronises the stream.
 */
int cfrk_scan_fasta_device(const void *d_bytes, int64_t n, int is_final, int64_t *d_header,
                           int64_t *d_start, int32_t *d_length, int64_t capacity,
                           int64_t *n_headers, void *stream);

/* ---- file level: replaces main() of src/main.cu:232-305 ----------------------------- */

#define CFRK_RUN_ALL_ROWS   1  /* print every read (chunk by chunk) instead of only the
                                  last nS mod chunkSize reads (src/main.cu:303-305)     */
#define CFRK_RUN_EXACT      2  /* CFRK_MODE_EXACT instead of compat, and the file is read the
                                  intended way: line terminators are not bases (k-mers span
                                  wrapped lines) and the last base is kept                 */
#define CFRK_RUN_SPARSE     4  /* write only non-zero bins (the filter commented out at
                                  src/main.cu:51,56)                                    */

/*
 * cfrk <fasta> <out> <k> [nt] [chunkSize] as a function: pinned multi-buffered FASTA
 * streamer (plain or gzip; records of any size) -> GPU record scan + count -> multi-threaded .cfrk
 * writer.  nt = host writer threads.  CFRK_EIO on a read error or a file truncated under the run.
 * k <= 8: dense rows (reference format).  k = 9..31 needs CFRK_RUN_SPARSE | CFRK_RUN_EXACT:
 * each row lists "kmer_index:count " for the k-mers that occur, in increasing index order.
 */
int cfrk_run_file(const char *fasta_path, const char *out_path, int k, int nt,
                  int64_t chunk_size, int flags, int device);

/*
 * The same over several GPUs of one box (the reference fans its chunks out over devCount pthreads --
 * all on one GPU -- and prints them in order, src/main.cu:208-230,277-289,303-305).  The file is cut
 * into spans at header lines; two host threads per listed device take spans in file order (each with
 * its own streams, device buffers and pinned row ring); per-read rows are gathered on the host and
 * reach the output in read order, byte-identical to the one-GPU run.  No collective: reads are
 * independent, and the compat spill across a span boundary is covered by each span's lookahead.
 * The input may be gzip-compressed (detected by its magic bytes).
 */
int cfrk_run_file_multi(const char *fasta_path, const char *out_path, int k, int nt,
                        int64_t chunk_size, int flags, const int *devices, int n_devices);

/*
 * PrintFreq (src/main.cu:26-63) as a function: n_rows dense rows of 4^k int32 bins -> .cfrk text
 * ("bin:count " tokens, rows separated by '\n', no trailing newline; out_path is truncated as by
 * fopen(.., "w")), formatted and written by nt host threads.  flags: CFRK_RUN_SPARSE skips zero bins.
 * Host only.  k = 1..8.
 */
int cfrk_write_rows(const char *out_path, const int32_t *rows, int64_t n_rows, int k, int nt, int flags);

#ifdef __cplusplus
}
#endif
#endif /* CFRK_B200_H */
