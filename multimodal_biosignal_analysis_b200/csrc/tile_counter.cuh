// Per-launch device counters shared by the persistent K1 kernels (fft_tma.cu owns the pool).
#pragma once
#include <cuda_runtime.h>

namespace cmc {

// Two unsigned words in zero-initialised device memory that the kernel using them leaves at zero again.
//   fft_segments_tma_pipe_kernel: next = tiles claimed beyond the static first wave, done = workers that have left;
//   dft_hann_tc_kernel:           next = store-phase epilogue warps that have finished, done = CTAs that have left.
struct TileCounter {
    unsigned next;
    unsigned done;
};

// A counter no other in-flight launch can share: one slot per (device, stream) for eager launches (launches of one
// stream are ordered), a fresh slot for every launch recorded during stream capture.  nullptr when the pool is
// exhausted (callers then take a path that needs no counter).
TileCounter* tile_counter_for(int dev, cudaStream_t st);

}  // namespace cmc
