// launch_spec.cu: see launch.h
#include "../../include/mcq.h"
#include "launch.h"
#include "spec.cuh"

namespace mcq {

template <bool FULL, bool REPLAY, bool EARLY, int NR, int LPC, int CN = 0, int HK = -1>
static cudaError_t launch_spec_one(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    auto k = spec_kernel<FULL, REPLAY, EARLY, NR, LPC, CN, HK>;
    // always the device maximum: the attribute is per function and per device, so concurrent host threads
    // (one engine each) must not race different values into it
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    if (e != cudaSuccess) return e;
    k<<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}

// the board sizes the reference's experiments use are compiled in altogether (N, Q = N^2 and the slab geometry
// become immediates: ~25 fewer instructions per round); other sizes take the geometry from the arguments
template <bool FULL, int CN>
static cudaError_t launch_spec_cn(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    constexpr int NR = spec_layout(FULL, CN, CN * CN).rounds;
    // ... and the per-step output with it: nothing, the statistics in difference form, or a uint16 history
    // (callers check fixed_n_serves() first: other combinations run the kernels that take the geometry at run time)
    if (a.hist_kind == MCQ_HIST_U16) return launch_spec_one<FULL, false, false, NR, 32, CN, 1>(a, grid, block, smem, s);
    if (a.dsum_e) return launch_spec_one<FULL, false, false, NR, 32, CN, 3>(a, grid, block, smem, s);   // (row offsets are 32-bit there)
    return launch_spec_one<FULL, false, false, NR, 32, CN, 0>(a, grid, block, smem, s);
}
static inline bool fixed_n_serves(const KArgs &a) {
    if (a.dsum_e && !a.stat_rows32) return false;
    return a.hist_kind == MCQ_HIST_NONE || (a.hist_kind == MCQ_HIST_U16 && !a.dsum_e);
}

// production kernels have the neighbour-row length compiled in; replay / early-stop ones take it at run time
template <bool FULL, int LPC>
static cudaError_t launch_spec_nr(const KArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
    if (LPC == 32 && a.Q == a.N * a.N && fixed_n_serves(a) && !getenv("MCQ_NO_FIXED_N")) {
        switch (a.N) {
            case 8: return launch_spec_cn<FULL, 8>(a, grid, block, smem, s);
            case 9: return launch_spec_cn<FULL, 9>(a, grid, block, smem, s);
            case 10: return launch_spec_cn<FULL, 10>(a, grid, block, smem, s);
            case 11: return launch_spec_cn<FULL, 11>(a, grid, block, smem, s);
            case 12: return launch_spec_cn<FULL, 12>(a, grid, block, smem, s);
            case 13: return launch_spec_cn<FULL, 13>(a, grid, block, smem, s);
            case 14: return launch_spec_cn<FULL, 14>(a, grid, block, smem, s);
            case 15: return launch_spec_cn<FULL, 15>(a, grid, block, smem, s);
            case 16: return launch_spec_cn<FULL, 16>(a, grid, block, smem, s);
            case 20: if (!FULL) return launch_spec_cn<false, 20>(a, grid, block, smem, s); break;
            default: break;
        }
    }
    switch (a.sl.rounds) {
        case 1: return launch_spec_one<FULL, false, false, 1, LPC>(a, grid, block, smem, s);
        case 2: return launch_spec_one<FULL, false, false, 2, LPC>(a, grid, block, smem, s);
        case 3: return launch_spec_one<FULL, false, false, 3, LPC>(a, grid, block, smem, s);
        case 4: return launch_spec_one<FULL, false, false, 4, LPC>(a, grid, block, smem, s);
        case 5: return launch_spec_one<FULL, false, false, 5, LPC>(a, grid, block, smem, s);
        case 6: return launch_spec_one<FULL, false, false, 6, LPC>(a, grid, block, smem, s);
        case 7: return launch_spec_one<FULL, false, false, 7, LPC>(a, grid, block, smem, s);
        default: return launch_spec_one<FULL, false, false, 8, LPC>(a, grid, block, smem, s);
    }
}

template <int LPC>
static cudaError_t launch_spec_lpc(const KArgs &a, bool replay, int grid, int block, size_t smem, cudaStream_t s) {
    if (a.full) return replay ? launch_spec_one<true, true, false, 0, LPC>(a, grid, block, smem, s) : launch_spec_nr<true, LPC>(a, grid, block, smem, s);
    if (a.patience >= 0) return replay ? launch_spec_one<false, true, true, 0, LPC>(a, grid, block, smem, s) : launch_spec_one<false, false, true, 0, LPC>(a, grid, block, smem, s);
    return replay ? launch_spec_one<false, true, false, 0, LPC>(a, grid, block, smem, s) : launch_spec_nr<false, LPC>(a, grid, block, smem, s);
}

cudaError_t launch_spec(int lpc, const KArgs &a, bool replay, int grid, int block, size_t smem, cudaStream_t s) {
    return lpc == 16 ? launch_spec_lpc<16>(a, replay, grid, block, smem, s) : launch_spec_lpc<32>(a, replay, grid, block, smem, s);
}


}  // namespace mcq
