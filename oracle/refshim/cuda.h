/*
 * oracle/refshim/cuda.h -- a stand-in for the CUDA runtime, just large enough to
 * compile the reference's own host + kernel sources for the CPU (SURVEY.md 8c,
 * "oracle #1").  TEST INFRASTRUCTURE ONLY; never part of the product.
 *
 * Kernels become plain functions; a launch `K<<<g,b>>>(args)` is rewritten by
 * oracle/Makefile (sed, on the fly, nothing is written back) to
 * `LAUNCH(K, g, b, args)`, which runs every (block, thread) pair in order.
 * Device allocations are zero-filled mmaps with slack on both sides, so the
 * reference's Freq[-1] store and its Index over-read stay harmless, as on a GPU.
 */
#ifndef CFRK_REFSHIM_CUDA_H
#define CFRK_REFSHIM_CUDA_H

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define __global__
#define __device__
#define __host__

struct shim_dim3 { unsigned x, y, z; };
inline shim_dim3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };

struct cudaDeviceProp {
    char name[256];
    size_t totalGlobalMem;
    int maxGridSize[3];
    int maxThreadsDim[3];
    int warpSize;
    int maxThreadsPerMultiProcessor;
};

static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    strcpy(p->name, "cpu-shim");
    p->totalGlobalMem = (size_t)180 << 30;
    p->maxGridSize[0] = 2147483647; p->maxGridSize[1] = p->maxGridSize[2] = 65535;
    p->maxThreadsDim[0] = p->maxThreadsDim[1] = 1024; p->maxThreadsDim[2] = 64;
    p->warpSize = 32;
    p->maxThreadsPerMultiProcessor = 2048;
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceReset() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "shim"; }

#define SHIM_SLACK 8192
static inline cudaError_t shim_alloc(void **p, size_t n)
{
    size_t total = n + 2 * SHIM_SLACK;
    void *m = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) { *p = NULL; return 2; }
    *p = (char *)m + SHIM_SLACK;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n) { return shim_alloc(p, n); }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return shim_alloc(p, n); }
static inline cudaError_t cudaFree(void *) { return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{ memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, int = 0)
{ memcpy(d, s, n); return cudaSuccess; }

static inline int atomicAdd(int *a, int v) { int o = *a; *a = o + v; return o; }

#define LAUNCH(kern, g, b, ...)                                                         \
    do {                                                                                \
        unsigned long shim_g = (unsigned long)(g), shim_b = (unsigned long)(b);         \
        gridDim.x = (unsigned)shim_g; blockDim.x = (unsigned)shim_b;                    \
        for (unsigned long shim_i = 0; shim_i < shim_g; shim_i++)                       \
            for (unsigned long shim_j = 0; shim_j < shim_b; shim_j++) {                 \
                blockIdx.x = (unsigned)shim_i; threadIdx.x = (unsigned)shim_j;          \
                kern(__VA_ARGS__);                                                      \
            }                                                                           \
    } while (0)

#endif
