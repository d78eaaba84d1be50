// ref_helper.cu -- OUR code, linked next to the reference's unmodified objects in
// oracle/_ref/libcfrk_ref_gpu.so so that a harness can release what the reference's kmer_main
// allocates and never frees (rd->Freq, src/kmer_main.cu:115).  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>
extern "C" int ref_free_host(void* p) { return (int)cudaFreeHost(p); }
extern "C" int ref_device_sync(void) { return (int)cudaDeviceSynchronize(); }
