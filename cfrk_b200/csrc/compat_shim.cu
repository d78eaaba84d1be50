// compat_shim.cu -- the reference's in-process operator symbol, backed by the new hot path.
//
//     void kmer_main(struct read *rd, lint nN, lint nS, int k, ushort device);
//                                    (reference src/kmer.cuh:6, mangled _Z9kmer_mainP4readllit)
//
// Linking the reference's unmodified main.cu + fastaIO.h against libcfrk_b200.so instead of its
// own kmer_main.cu/kmer_kernel.cu gives "reference driver, new hot path" (INTEGRATION.md).
// The struct below is the reference's batch layout (src/tipos.h:23-30); it is the ABI.
#include "../../include/cfrk_b200.h"
#include "internal.h"

#include <cuda_runtime.h>
#include <cstdio>

typedef long int lint;

struct read {
    char* data;         // nN codes {0,1,2,3,-1}, one -1 terminator per read (pinned, caller-owned)
    int* length;        // nS
    lint* start;        // nS, chunk-local byte offsets
    int* Freq;          // OUT: allocated here with cudaMallocHost, as src/kmer_main.cu:115 does
    struct read* next;  // unused
};

void kmer_main(struct read* rd, lint nN, lint nS, int k, unsigned short device)
{
    static_assert(sizeof(lint) == sizeof(int64_t), "LP64 expected");
    const size_t bytes = (size_t)nS * ((size_t)1 << (2 * k)) * sizeof(int);
    cudaSetDevice(device);
    // pinned like the reference's (src/kmer_main.cu:115), but from the arena: a caller that hands the
    // rows back with cfrk_free_host() pays the pinning once, not per call
    rd->Freq = static_cast<int*>(cfrk::pinned_alloc(bytes ? bytes : 4));
    if (!rd->Freq) {
        // same channel as the reference: report on stdout and carry on (src/kmer_main.cu:115)
        printf("\n[Error 9] %s\n", "cudaMallocHost failed");
        return;
    }
    int rc = cfrk_count_dense_host(rd->data, CFRK_FMT_CODES, reinterpret_cast<const int64_t*>(rd->start),
                                   rd->length, nN, nS, k, CFRK_MODE_COMPAT, device, rd->Freq);
    if (rc != CFRK_OK) printf("\n[Error cfrk_b200 %d] %s\n", rc, cfrk_last_error());
}
