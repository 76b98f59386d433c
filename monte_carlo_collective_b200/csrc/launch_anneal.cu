// launch_anneal.cu: see launch.h
#include "../../include/mcq.h"
#include "launch.h"

namespace mcq {

template <int G, bool FULL, bool REPLAY>
static cudaError_t launch_one(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    auto k = anneal_kernel<G, FULL, REPLAY>;
    // always the device maximum: the attribute is per function and per device, so concurrent host threads
    // (one engine each) must not race different values into it
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    if (e != cudaSuccess) return e;
    k<<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}

template <int G>
static cudaError_t launch_g(const KArgs &a, bool replay, int grid, int block, size_t smem, cudaStream_t s) {
    if (a.full) return replay ? launch_one<G, true, true>(a, grid, block, smem, s) : launch_one<G, true, false>(a, grid, block, smem, s);
    return replay ? launch_one<G, false, true>(a, grid, block, smem, s) : launch_one<G, false, false>(a, grid, block, smem, s);
}

cudaError_t launch_anneal(int G, const KArgs &a, bool replay, int grid, int block, size_t smem, cudaStream_t s) {
    switch (G) {
        case 1: return launch_g<1>(a, replay, grid, block, smem, s);
        case 4: return launch_g<4>(a, replay, grid, block, smem, s);
        case 8: return launch_g<8>(a, replay, grid, block, smem, s);
        case 16: return launch_g<16>(a, replay, grid, block, smem, s);
        default: return launch_g<32>(a, replay, grid, block, smem, s);
    }
}


cudaError_t launch_gslab_build(const KArgs &a, int n_chains, cudaStream_t s) {
    gslab_build_kernel<<<n_chains, 32, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mcq
