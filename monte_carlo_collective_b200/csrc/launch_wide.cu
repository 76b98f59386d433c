// launch_wide.cu: see launch.h
#include "../../include/mcq.h"
#include "launch.h"
#include "wide.cuh"

namespace mcq {

template <bool FULL, bool EARLY, int NT>
static cudaError_t launch_wide_one(const KArgs &a, int grid, size_t smem, cudaStream_t s) {
    auto k = wide_kernel<FULL, EARLY, NT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    if (e != cudaSuccess) return e;
    k<<<grid, NT, smem, s>>>(a);
    return cudaGetLastError();
}

template <int NT>
static cudaError_t launch_wide_nt(const KArgs &a, int grid, size_t smem, cudaStream_t s) {
    if (a.full) return launch_wide_one<true, false, NT>(a, grid, smem, s);
    return a.patience >= 0 ? launch_wide_one<false, true, NT>(a, grid, smem, s) : launch_wide_one<false, false, NT>(a, grid, smem, s);
}

cudaError_t launch_wide(int threads, const KArgs &a, int grid, size_t smem, cudaStream_t s) {
    switch (threads) {
        case 32: return launch_wide_nt<32>(a, grid, smem, s);
        case 64: return launch_wide_nt<64>(a, grid, smem, s);
        case 128: return launch_wide_nt<128>(a, grid, smem, s);
        default: return launch_wide_nt<256>(a, grid, smem, s);
    }
}


}  // namespace mcq
