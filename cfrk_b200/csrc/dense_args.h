// dense_args.h -- argument block shared by the dense per-read kernels (kernels.cu, dense_lane.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cfrk {

struct DenseArgs {
    const uint8_t* bases;
    const uint16_t* valid;   // FMT_PACKED only: validity masks (bases = the uint32 codes)
    const int64_t* start;
    const int32_t* length;
    int64_t nS;          // reads in the batch (halo / spill scope)
    int64_t nN;          // bytes in the bases buffer
    int64_t read_begin;  // rows [read_begin, read_end) are produced by this launch
    int64_t read_end;
    uint32_t* out;       // row read_begin at out[0]
    int mode;
    int64_t num_tiles;
    int64_t chunk_size;   // compat: reads with (index_base + i) % chunk_size == 0 start a reference
    int64_t index_base;   //         chunk (their spill is dropped); 0 = only read 0 does
    int flags;            // bit 0: big-row path zeroes with plain stores instead of TMA (A/B switch)
    uint32_t* handoff;    // warp tiles, compat: one word per tile boundary (zeroed), or null
};

// handoff[t] != 0: the first read of tile t+1 had that many data-dependent invalid windows -> last bin
// of the last row of tile t (kernels.cu)
cudaError_t launch_spill_fixup(const uint32_t* handoff, int64_t ntiles, uint32_t* out, int64_t tile_bins, cudaStream_t st);

// small per-(device, stream) scratch that survives between launches (kernels.cu)
cudaError_t stream_scratch(cudaStream_t st, size_t bytes, void** out);
void count_launch();

// k = 1..4: lane-per-read tiles with TMA-staged input (dense_lane.cu)
cudaError_t launch_dense_lane(int k, int fmt, const DenseArgs& a, cudaStream_t st);
int dense_lane_reads_per_tile(int k);

}  // namespace cfrk
