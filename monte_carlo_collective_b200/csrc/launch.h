// Kernel launchers, one translation unit per kernel family so that the library builds in parallel
// (each family is a few dozen template instantiations).
#pragma once
#include <cuda_runtime.h>

#include "anneal.cuh"

namespace mcq {

extern int g_smem_optin;   // sharedMemPerBlockOptin of the device (mcq_api.cu, set by mcq_create)

// wide.cuh geometry the host needs for its shared-memory budget
constexpr int BETA_PAD = 64;                  // floats behind the schedule table (fast.cuh reads past the last step of a launch)
constexpr int WIDE_THREADS = 256;             // widest CTA (one per SM on the largest boards); 128 and 64 where more CTAs fit
constexpr int WIDE_JCAP = 62;                 // journal of state elements changed since the last best-state snapshot
constexpr int WIDE_XCH_BYTES = 384;           // exchange words (3 per warp), journal count, journal, published moves (3 words per warp)

cudaError_t launch_anneal(int G, const KArgs &a, bool replay, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t launch_spec(int lpc, const KArgs &a, bool replay, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t launch_wide(int threads, const KArgs &a, int grid, size_t smem, cudaStream_t s);
// fast.cuh: production runs without early stop, history none / uint16 / statistics (fast_serves says whether)
bool fast_serves(int lpc, const KArgs &a, bool replay);
cudaError_t launch_fast(int lpc, const KArgs &a, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t launch_gslab_build(const KArgs &a, int n_chains, cudaStream_t s);

}  // namespace mcq
