/*
 * cfrk_oracle.c -- CPU restatement of CFRK's per-read k-mer counting path.
 *
 * TEST INFRASTRUCTURE ONLY (see cfrk_oracle.h).  Parity status: PINNED against
 * the reference's goldens and against the reference's own code run on the CPU
 * (tests/test_oracle_golden.py: goldens through the stand-in inputs, the sha256 manifest of the reference's own
 * sources compiled for the CPU, and a live comparison with oracle/_ref/cfrk_ref_cpu).
 *
 * Written from the behavioural spec in SURVEY.md 8(c); each function cites the
 * reference lines it restates.  Deliberately the *slow obvious* algorithm
 * (O(k) per window, one pass per read) for oracle_count_compat/exact so that it
 * shares nothing with the rolling/bit-parallel CUDA path it checks; the fast
 * multithreaded variant (the reported CPU baseline) is checked against it.
 */
#define _GNU_SOURCE
#include "cfrk_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define REF_BLOCK_THREADS 1024 /* maxThreadsDim[0]; src/kmer_main.cu:82 */

static inline int64_t four_pow(int k) { return (int64_t)1 << (2 * k); }

/* src/fastaIO.h:123-139 */
int8_t oracle_encode_base(unsigned char c)
{
    switch (c) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    default:            return -1;
    }
}

void oracle_free_reads(oracle_reads *r)
{
    if (!r) return;
    free(r->data); free(r->length); free(r->start);
    memset(r, 0, sizeof *r);
}

/*
 * src/fastaIO.h:12-22  nS = `grep -c ">"`  (lines containing '>' anywhere)
 * src/fastaIO.h:38-69  a record starts at a line whose first byte is '>'; every other
 *                      line, newline included, is strcat'ed to the record text;
 *                      len = strlen(text) - 1
 * src/fastaIO.h:114-141 per-base switch for j < len
 * src/fastaIO.h:74-102 concatenation with one -1 terminator per read, start[] offsets
 */
int oracle_parse_fasta_mem(const char *buf, size_t n, oracle_reads *out)
{
    return oracle_parse_fasta_mem_ex(buf, n, 0, out);
}

int oracle_parse_fasta_mem_ex(const char *buf, size_t n, int unwrap, oracle_reads *out)
{
    memset(out, 0, sizeof *out);
    /* pass 1: count lines with '>' (grep) and header lines, bound the text size */
    int64_t grep_count = 0, headers = 0;
    size_t pos = 0;
    int seen_header = 0;
    while (pos < n) {
        const char *nl = memchr(buf + pos, '\n', n - pos);
        size_t end = nl ? (size_t)(nl - buf) + 1 : n; /* getline keeps the '\n' */
        int has_gt = memchr(buf + pos, '>', end - pos) != NULL;
        if (has_gt) grep_count++;
        if (buf[pos] == '>') { headers++; seen_header = 1; }
        else {
            if (!seen_header) return -2;           /* seq[-1] in the reference: undefined */
            if (has_gt) return -3;                 /* nS over-counted: undefined */
        }
        pos = end;
    }
    if (grep_count != headers) return -3;
    int64_t nS = headers;
    out->nS = nS;
    out->length = (int32_t *)calloc((size_t)(nS > 0 ? nS : 1), sizeof(int32_t));
    out->start  = (int64_t *)calloc((size_t)(nS > 0 ? nS : 1), sizeof(int64_t));
    out->data   = (int8_t *)malloc(n + (size_t)nS + 16);
    if (!out->length || !out->start || !out->data) { oracle_free_reads(out); return -4; }

    /* pass 2: build records */
    int64_t rec = -1, w = 0;
    int64_t text_begin = 0; /* where the current record's text starts in out->data */
#define CLOSE_RECORD()                                                                   \
    do {                                                                                 \
        /* len = strlen(text) - 1, never below 0 (a header with no sequence line is      \
         * undefined in the reference; defined here as a read of length 0) */            \
        int64_t tlen = w - text_begin;                                                   \
        int64_t len = unwrap ? tlen : (tlen > 0 ? tlen - 1 : 0);                         \
        w = text_begin + len;                                                            \
        out->length[rec] = (int32_t)len;                                                 \
        out->start[rec] = text_begin;                                                    \
        out->data[w++] = -1; /* terminator, src/fastaIO.h:96 */                          \
    } while (0)
    pos = 0;
    while (pos < n) {
        const char *nl = memchr(buf + pos, '\n', n - pos);
        size_t end = nl ? (size_t)(nl - buf) + 1 : n;
        if (buf[pos] == '>') {
            if (rec >= 0) CLOSE_RECORD();
            rec++;
            text_begin = w;
        } else {
            /* strcat copies up to the first NUL of the line */
            const char *z = memchr(buf + pos, '\0', end - pos);
            size_t cpy = z ? (size_t)(z - (buf + pos)) : end - pos;
            for (size_t j = 0; j < cpy; j++) {
                if (unwrap && (buf[pos + j] == '\n' || buf[pos + j] == '\r')) continue;
                out->data[w++] = oracle_encode_base((unsigned char)buf[pos + j]);
            }
        }
        pos = end;
    }
    if (rec >= 0) CLOSE_RECORD();
#undef CLOSE_RECORD
    out->nN = w;
    return 0;
}

int oracle_parse_fasta(const char *path, oracle_reads *out) { return oracle_parse_fasta_ex(path, 0, out); }

int oracle_parse_fasta_ex(const char *path, int unwrap, oracle_reads *out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1; /* src/fastaIO.h:36 exit(EXIT_FAILURE) */
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); return -4; }
    size_t got = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    int rc = oracle_parse_fasta_mem_ex(buf, got, unwrap, out);
    free(buf);
    return rc;
}

/*
 * Index of the window starting at byte p: src/kmer_kernel.cu:30-47.
 * -1 at the first -1 byte, else sum code_i * 4^(k-1-i).  (The reference
 * accumulates in float32, exact for k <= 12; integers here.)
 */
static inline int64_t window_index(const int8_t *data, int64_t nN, int64_t p, int k)
{
    int64_t idx = 0;
    for (int i = 0; i < k; i++) {
        int64_t q = p + i;
        int c = (q < nN) ? data[q] : -1; /* the reference never reads past a terminator */
        if (c == -1) return -1;
        idx += (int64_t)c * four_pow(k - 1 - i);
    }
    return idx;
}

/*
 * src/kmer_kernel.cu:73-90 with the launch shape of src/kmer_main.cu:82-88,111:
 * block i <-> read i, thread t < length[i]-1 (and t < 1024) adds 1 at
 * Freq[4^k*i + Index[start[i]+t]]; Index==-1 therefore lands on the last bin of
 * read i-1, and for i==0 on Freq[-1] (lost).
 */
void oracle_count_compat(const int8_t *data, const int64_t *start, const int32_t *length,
                         int64_t nN, int64_t nS, int k, int32_t *freq)
{
    const int64_t fourk = four_pow(k);
    memset(freq, 0, (size_t)(nS * fourk) * sizeof(int32_t)); /* SetMatrix, src/kmer_main.cu:108 */
    for (int64_t i = 0; i < nS; i++) {
        /* `threadIdx.x < length[i]-1` compares unsigned (src/kmer_kernel.cu:85): for an EMPTY
         * read the bound is (unsigned)-1 and all 1024 threads pass, so the block walks over
         * the bytes that follow (terminators and later reads) and counts them into row i.
         * Positions at or beyond nN read Index[] out of bounds in the reference (garbage);
         * they are not counted here. */
        int64_t visited = length[i] == 0 ? REF_BLOCK_THREADS : (int64_t)length[i] - 1;
        if (visited > REF_BLOCK_THREADS) visited = REF_BLOCK_THREADS;
        for (int64_t t = 0; t < visited; t++) {
            if (start[i] + t >= nN) break;
            int64_t pos = fourk * i + window_index(data, nN, start[i] + t, k);
            if (pos >= 0) freq[pos] += 1;
        }
    }
}

void oracle_count_exact(const int8_t *data, const int64_t *start, const int32_t *length,
                        int64_t nN, int64_t nS, int k, int32_t *freq)
{
    const int64_t fourk = four_pow(k);
    memset(freq, 0, (size_t)(nS * fourk) * sizeof(int32_t));
    for (int64_t i = 0; i < nS; i++) {
        for (int64_t t = 0; t + k <= (int64_t)length[i]; t++) {
            int64_t idx = window_index(data, nN, start[i] + t, k);
            if (idx >= 0) freq[fourk * i + idx] += 1;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Fast multithreaded counter (CPU baseline).  Rolling 2-bit index.           */

typedef struct {
    const int8_t *data; const int64_t *start; const int32_t *length;
    int64_t nN, nS; int k, mode, ascii;
    int64_t r0, r1;
    int32_t *freq; uint64_t *hist; /* one of the two */
} mt_job;

static inline int code_of(const mt_job *j, int64_t p)
{
    int8_t b = j->data[p];
    return j->ascii ? oracle_encode_base((unsigned char)b) : (b == -1 ? -1 : (b & 3));
}

/* number of counted window-end positions for a read; see DESIGN.md "compat algebra" */
static void scan_read(const mt_job *j, int64_t i, int32_t *row, uint64_t *hist, int64_t *invalid_out)
{
    const int k = j->k;
    const uint64_t mask = (k >= 32) ? ~0ull : (((uint64_t)1 << (2 * k)) - 1);
    int64_t len = j->length[i];
    int64_t visited;
    if (j->mode == ORACLE_MODE_COMPAT) {
        if (len == 0) {   /* empty read: the unsigned compare lets all 1024 threads through */
            len = j->nN - j->start[i];   /* ... over whatever follows in the buffer */
            visited = len;
        } else {
            visited = len - 1;
        }
        if (visited > REF_BLOCK_THREADS) visited = REF_BLOCK_THREADS;
        if (visited < 0) visited = 0;
    } else {
        visited = len - k + 1; if (visited < 0) visited = 0;
    }
    int64_t tend = visited + k - 1; if (tend > len) tend = len; /* last byte whose window counts */
    if (visited == 0) tend = 0;
    int64_t valid = 0;
    uint64_t idx = 0; int run = 0;
    const int64_t s = j->start[i];
    for (int64_t t = 0; t < tend; t++) {
        int c = code_of(j, s + t);
        if (c < 0) { run = 0; idx = 0; }
        else { idx = ((idx << 2) | (uint64_t)c) & mask; if (run < k) run++; }
        if (run >= k) { /* window starting at t-k+1 < visited by construction of tend */
            valid++;
            if (row) row[idx] += 1; else hist[idx] += 1;
        }
    }
    if (invalid_out) *invalid_out = visited - valid;
}

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    const int64_t fourk = four_pow(j->k);
    for (int64_t i = j->r0; i < j->r1; i++) {
        int32_t *row = j->freq + fourk * i;
        int64_t inv = 0;
        /* the owner of read i zeroes row i before anybody spills into it: the spill into
         * row i comes from read i+1, handled after row i by the same thread (or, at the
         * end of the range, by this thread through the halo below) */
        if (i == j->r0) memset(row, 0, (size_t)fourk * sizeof(int32_t));
        if (i + 1 < j->r1) memset(row + fourk, 0, (size_t)fourk * sizeof(int32_t));
        scan_read(j, i, row, NULL, &inv);
        if (j->mode == ORACLE_MODE_COMPAT && i > j->r0 && inv > 0)
            row[-1] += (int32_t)inv; /* last bin of read i-1 */
    }
    /* halo: spill of the first read of the next range into our last row */
    if (j->mode == ORACLE_MODE_COMPAT && j->r1 < j->nS && j->r1 > j->r0) {
        int64_t inv = 0;
        mt_job tmp = *j;
        static __thread int32_t *scratch = NULL; static __thread int64_t scratch_n = 0;
        if (scratch_n < fourk) { free(scratch); scratch = (int32_t *)malloc((size_t)fourk * 4); scratch_n = fourk; }
        memset(scratch, 0, (size_t)fourk * 4);
        scan_read(&tmp, j->r1, scratch, NULL, &inv);
        if (inv > 0) j->freq[fourk * j->r1 - 1] += (int32_t)inv;
    }
    return NULL;
}

static int clamp_threads(int nthreads, int64_t nS)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > nS) nthreads = nS > 0 ? (int)nS : 1;
    return nthreads;
}

void oracle_count_fast_mt(const int8_t *data, const int64_t *start, const int32_t *length,
                          int64_t nN, int64_t nS, int k, int mode, int ascii,
                          int nthreads, int32_t *freq)
{
    if (nS <= 0) return;
    nthreads = clamp_threads(nthreads, nS);
    pthread_t th[256]; mt_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        mt_job *j = &jobs[t];
        j->data = data; j->start = start; j->length = length; j->nN = nN; j->nS = nS;
        j->k = k; j->mode = mode; j->ascii = ascii; j->freq = freq; j->hist = NULL;
        j->r0 = nS * t / nthreads; j->r1 = nS * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, mt_worker, j);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

static void *hist_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    for (int64_t i = j->r0; i < j->r1; i++) scan_read(j, i, NULL, j->hist, NULL);
    return NULL;
}

void oracle_global_hist(const int8_t *data, const int64_t *start, const int32_t *length,
                        int64_t nN, int64_t nS, int k, int ascii, int nthreads, uint64_t *hist)
{
    const int64_t fourk = four_pow(k);
    memset(hist, 0, (size_t)fourk * sizeof(uint64_t));
    if (nS <= 0) return;
    nthreads = clamp_threads(nthreads, nS);
    pthread_t th[256]; mt_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        mt_job *j = &jobs[t];
        j->data = data; j->start = start; j->length = length; j->nN = nN; j->nS = nS;
        j->k = k; j->mode = ORACLE_MODE_EXACT; j->ascii = ascii; j->freq = NULL;
        j->hist = t == 0 ? hist : (uint64_t *)calloc((size_t)fourk, sizeof(uint64_t));
        j->r0 = nS * t / nthreads; j->r1 = nS * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, hist_worker, j);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    for (int t = 1; t < nthreads; t++) {
        for (int64_t b = 0; b < fourk; b++) hist[b] += jobs[t].hist[b];
        free(jobs[t].hist);
    }
}

/* ------------------------------------------------------------------------- */

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

int64_t oracle_count_sparse(const int8_t *data, const int64_t *start, const int32_t *length,
                            int64_t nN, int64_t nS, int k, int ascii,
                            int64_t *row_ptr, uint64_t *keys, uint32_t *counts, int64_t cap)
{
    (void)nN;
    const uint64_t mask = (k >= 32) ? ~0ull : (((uint64_t)1 << (2 * k)) - 1);
    int64_t out = 0, tmp_cap = 0;
    uint64_t *tmp = NULL;
    for (int64_t i = 0; i < nS; i++) {
        row_ptr[i] = out;
        int64_t len = length[i], m = 0;
        if (len > tmp_cap) { free(tmp); tmp = (uint64_t *)malloc((size_t)len * 8); tmp_cap = len; }
        uint64_t idx = 0; int run = 0;
        for (int64_t t = 0; t < len; t++) {
            int8_t b = data[start[i] + t];
            int c = ascii ? oracle_encode_base((unsigned char)b) : (b == -1 ? -1 : (b & 3));
            if (c < 0) { run = 0; idx = 0; continue; }
            idx = ((idx << 2) | (uint64_t)c) & mask; if (run < k) run++;
            if (run >= k) tmp[m++] = idx;
        }
        qsort(tmp, (size_t)m, sizeof(uint64_t), cmp_u64);
        for (int64_t a = 0; a < m;) {
            int64_t b = a; while (b < m && tmp[b] == tmp[a]) b++;
            if (out >= cap) { free(tmp); return -1; }
            keys[out] = tmp[a]; counts[out] = (uint32_t)(b - a); out++;
            a = b;
        }
    }
    row_ptr[nS] = out;
    free(tmp);
    return out;
}

/* ------------------------------------------------------------------------- */

/* src/main.cu:36-60: "%d:%d " per bin, "\n" before every row but the first, nothing at EOF */
static int write_rows(FILE *f, const int32_t *freq, int64_t rows, int k, int *first)
{
    const int64_t fourk = four_pow(k);
    for (int64_t r = 0; r < rows; r++) {
        if (!*first) fputc('\n', f);
        *first = 0;
        for (int64_t b = 0; b < fourk; b++)
            fprintf(f, "%d:%d ", (int)b, freq[r * fourk + b]);
    }
    return 0;
}

int oracle_write_cfrk(const char *path, const int32_t *freq, int64_t rows, int k)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    int first = 1;
    write_rows(f, freq, rows, k, &first);
    fclose(f);
    return 0;
}

/*
 * src/main.cu:270-305.  nChunk = nS / chunkSize full chunks are computed and printed
 * into the output file, which is then re-opened with "w" for the remainder chunk
 * (src/main.cu:34,303,305): only reads [nChunk*chunkSize, nS) survive.  Each chunk is a
 * separate kmer_main call, so the spill of its first read is dropped.
 */
int oracle_run_cli(const char *fasta, const char *out, int k, int64_t chunk_size,
                   int mode, int all_rows)
{
    if (chunk_size <= 0 || k < 1 || k > 12) return -5;
    oracle_reads rd;
    /* exact mode reads the file the intended way: lines unwrapped, last base kept */
    int rc = oracle_parse_fasta_ex(fasta, mode == ORACLE_MODE_EXACT, &rd);
    if (rc) return rc;
    const int64_t fourk = four_pow(k);
    FILE *f = fopen(out, "w");
    if (!f) { oracle_free_reads(&rd); return -1; }
    int first = 1;
    int64_t first_chunk = all_rows ? 0 : rd.nS / chunk_size;
    for (int64_t c = first_chunk; c * chunk_size < rd.nS; c++) {
        int64_t r0 = c * chunk_size, r1 = r0 + chunk_size;
        if (r1 > rd.nS) r1 = rd.nS;
        int64_t n = r1 - r0;
        /* chunk-local copy, as SelectChunk / SelectChunkRemain build it (src/main.cu:110-206) */
        int64_t b0 = rd.start[r0];
        int64_t b1 = rd.start[r1 - 1] + rd.length[r1 - 1] + 1;
        int64_t *st = (int64_t *)malloc((size_t)n * 8);
        for (int64_t i = 0; i < n; i++) st[i] = rd.start[r0 + i] - b0;
        int32_t *freq = (int32_t *)malloc((size_t)(n * fourk) * 4);
        if (mode == ORACLE_MODE_COMPAT)
            oracle_count_compat(rd.data + b0, st, rd.length + r0, b1 - b0, n, k, freq);
        else
            oracle_count_exact(rd.data + b0, st, rd.length + r0, b1 - b0, n, k, freq);
        write_rows(f, freq, n, k, &first);
        free(freq); free(st);
    }
    fclose(f);
    oracle_free_reads(&rd);
    return 0;
}
