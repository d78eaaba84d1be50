// ref_helper.cu -- OUR code, linked next to the reference's unmodified objects in
// oracle/_ref/libcfrk_ref_gpu.so.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py --impl reference).
//   ref_free_host      releases what the reference's kmer_main allocates and never frees
//                      (rd->Freq, src/kmer_main.cu:115).
//   ref_kernels_time   SURVEY 8(d) baseline 1b: the reference's own four kernel launches
//                      (src/kmer_main.cu:107-111) on DEVICE-RESIDENT inputs, with the launch shapes
//                      of src/kmer_main.cu:66-100, timed with CUDA events -- the reference's kernels
//                      without its per-call cudaMalloc/cudaMallocHost/copies.  The kernels are the
//                      reference's objects (kmer_kernel.o); only the launch arithmetic is restated.
#include <cuda_runtime.h>
#include <math.h>
#include "kmer.cuh"   // the reference's own declarations (-I $(REF))

extern "C" int ref_free_host(void* p) { return (int)cudaFreeHost(p); }
extern "C" int ref_device_sync(void) { return (int)cudaDeviceSynchronize(); }

// d_Seq/d_start/d_length: device copies of struct read's data/start/length (src/tipos.h:23-30).
// Returns 0 and the average milliseconds of one SetMatrix+SetMatrix+ComputeIndex+ComputeFreqNew
// sequence over `reps` repetitions (after one untimed repetition), or a CUDA error code.
extern "C" int ref_kernels_time(char* d_Seq, lint* d_start, int* d_length, lint nN, lint nS, int k, int reps, float* ms_out)
{
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    const lint maxGridSize = prop.maxGridSize[0];       // src/kmer_main.cu:30-33
    const int maxThreadDim = prop.maxThreadsDim[0];
    const int fourk = POW(k);
    const lint nF64 = nS * (lint)fourk;
    if (nF64 >= 2147483647LL) return -1;                // int nF overflows in the reference (src/kmer_main.cu:90)
    const int nF = (int)nF64;
    int *d_Index = 0, *d_Freq_alloc = 0;
    cudaError_t e;
    // + 2048 ints: ComputeFreqNew loads Index[start[i] + threadIdx.x] for all 1024 threads (src/kmer_kernel.cu:83)
    if ((e = cudaMalloc((void**)&d_Index, ((size_t)nN + 2048) * sizeof(int))) != cudaSuccess) return (int)e;
    // 256 bytes in front of Freq: the reference's Freq[-1] store for the first read (src/kmer_kernel.cu:84-87)
    if ((e = cudaMalloc((void**)&d_Freq_alloc, (size_t)nF * sizeof(int) + 256)) != cudaSuccess) { cudaFree(d_Index); return (int)e; }
    int* d_Freq = d_Freq_alloc + 64;
    // launch shapes, src/kmer_main.cu:66-100
    int block[4], grid[4];
    ushort offset[4] = {1, 1, 1, 1};
    block[0] = maxThreadDim;
    grid[0] = (int)floor((double)(nN / block[0])) + 1;
    if (grid[0] > maxGridSize) { grid[0] = (int)maxGridSize; offset[0] = (ushort)((nN / ((lint)grid[0] * block[0])) + 1); }
    block[2] = maxThreadDim;
    grid[2] = (int)nS;
    if (nS > maxGridSize) { grid[2] = (int)maxGridSize; offset[2] = (ushort)((nS / grid[2]) + 1); }
    block[3] = maxThreadDim;
    grid[3] = (nF / 1024) + 1;
    if (grid[3] > maxGridSize) { grid[3] = (int)maxGridSize; offset[3] = (ushort)((nF / ((lint)grid[3] * block[3])) + 1); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int r = 0; r <= reps; r++) {
        if (r == 1) cudaEventRecord(e0, cudaStreamPerThread);
        SetMatrix<<<grid[0], block[0], 0, cudaStreamPerThread>>>(d_Index, offset[0], -1, (int)nN);
        SetMatrix<<<grid[3], block[3], 0, cudaStreamPerThread>>>(d_Freq, offset[3], 0, nF);
        ComputeIndex<<<grid[0], block[0], 0, cudaStreamPerThread>>>(d_Seq, d_Index, k, nN, offset[0]);
        ComputeFreqNew<<<grid[2], block[2], 0, cudaStreamPerThread>>>(d_Index, d_Freq, d_start, d_length, offset[2], fourk, nS);
    }
    cudaEventRecord(e1, cudaStreamPerThread);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = reps > 0 ? ms / reps : 0.f;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_Index);
    cudaFree(d_Freq_alloc);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}
