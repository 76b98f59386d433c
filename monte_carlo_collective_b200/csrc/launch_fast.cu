// launch_fast.cu: see launch.h
#include <stdlib.h>

#include "../../include/mcq.h"
#include "launch.h"
#include "fast.cuh"

namespace mcq {

template <bool FULL, int NR, int LPC, int CN, int HK>
static cudaError_t launch_fast_one(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    auto k = fast_kernel<FULL, NR, LPC, CN, HK>;
    // always the device maximum: the attribute is per function and per device, so concurrent host threads
    // (one engine each) must not race different values into it
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    if (e != cudaSuccess) return e;
    k<<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}

template <bool FULL, int NR, int LPC, int CN>
static cudaError_t launch_fast_hk(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    if (a.hist_kind == MCQ_HIST_U16) return launch_fast_one<FULL, NR, LPC, CN, 1>(a, grid, block, smem, s);
    if (a.dsum_e) return launch_fast_one<FULL, NR, LPC, CN, 3>(a, grid, block, smem, s);
    return launch_fast_one<FULL, NR, LPC, CN, 0>(a, grid, block, smem, s);
}

// the board sizes the reference's experiments use are compiled in altogether (N, Q = N^2 and the slab geometry become
// immediates); other sizes take the geometry from the arguments and only the neighbour-row length is compiled in
template <bool FULL, int LPC, int CN>
static cudaError_t launch_fast_cn(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    constexpr int NR = spec_layout(FULL, CN, CN * CN).rounds;
    return launch_fast_hk<FULL, NR, LPC, CN>(a, grid, block, smem, s);
}

template <bool FULL, int LPC>
static cudaError_t launch_fast_n(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    if (a.Q == a.N * a.N && !getenv("MCQ_NO_FIXED_N")) {
        switch (a.N) {
            case 8: return launch_fast_cn<FULL, LPC, 8>(a, grid, block, smem, s);
            case 9: return launch_fast_cn<FULL, LPC, 9>(a, grid, block, smem, s);
            case 10: return launch_fast_cn<FULL, LPC, 10>(a, grid, block, smem, s);
            case 11: return launch_fast_cn<FULL, LPC, 11>(a, grid, block, smem, s);
            case 12: return launch_fast_cn<FULL, LPC, 12>(a, grid, block, smem, s);
            case 13: return launch_fast_cn<FULL, LPC, 13>(a, grid, block, smem, s);
            case 14: return launch_fast_cn<FULL, LPC, 14>(a, grid, block, smem, s);
            case 15: return launch_fast_cn<FULL, LPC, 15>(a, grid, block, smem, s);
            case 16: return launch_fast_cn<FULL, LPC, 16>(a, grid, block, smem, s);
            case 20: if (!FULL) return launch_fast_cn<false, LPC, 20>(a, grid, block, smem, s); break;
            default: break;
        }
    }
    if (LPC == 32) {
        switch (a.sl.rounds) {
            case 1: return launch_fast_hk<FULL, 1, 32, 0>(a, grid, block, smem, s);
            case 2: return launch_fast_hk<FULL, 2, 32, 0>(a, grid, block, smem, s);
            case 3: return launch_fast_hk<FULL, 3, 32, 0>(a, grid, block, smem, s);
            case 4: return launch_fast_hk<FULL, 4, 32, 0>(a, grid, block, smem, s);
            case 5: return launch_fast_hk<FULL, 5, 32, 0>(a, grid, block, smem, s);
            case 6: return launch_fast_hk<FULL, 6, 32, 0>(a, grid, block, smem, s);
            case 7: return launch_fast_hk<FULL, 7, 32, 0>(a, grid, block, smem, s);
            default: return launch_fast_hk<FULL, 8, 32, 0>(a, grid, block, smem, s);
        }
    }
    return cudaErrorInvalidValue;   // fast_serves() said no
}

static bool fixed_n(const KArgs &a) {
    if (a.Q != a.N * a.N || getenv("MCQ_NO_FIXED_N")) return false;
    return (a.N >= 8 && a.N <= 16) || (!a.full && a.N == 20);
}

bool fast_serves(int lpc, const KArgs &a, bool replay) {
    if (replay || a.patience >= 0 || getenv("MCQ_NO_FAST")) return false;
    if (a.dsum_e && (!a.stat_rows32 || a.hist_kind != MCQ_HIST_NONE)) return false;   // statistics: 32-bit row offsets, no history beside them
    if (a.hist_kind != MCQ_HIST_NONE && a.hist_kind != MCQ_HIST_U16) return false;
    return lpc == 32 || fixed_n(a);                                                   // two chains per warp: compiled-in sizes only
}

cudaError_t launch_fast(int lpc, const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    if (lpc == 16) return a.full ? launch_fast_n<true, 16>(a, grid, block, smem, s) : launch_fast_n<false, 16>(a, grid, block, smem, s);
    return a.full ? launch_fast_n<true, 32>(a, grid, block, smem, s) : launch_fast_n<false, 32>(a, grid, block, smem, s);
}

}  // namespace mcq
