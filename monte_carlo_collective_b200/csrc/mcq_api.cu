// libmcq: C ABI + host orchestration for the B200 annealing engine (see include/mcq.h).
//
// Host side of the hot path: what run_experiment (experiments.py:475-573) does with a process
// pool -- start n_runs chains, collect histories / best energies / accept lists -- is done here
// as a handful of kernel launches over a batch of chains resident in shared memory.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/mcq.h"
#include "anneal.cuh"
#include "geometry.cuh"
#include "launch.h"

namespace mcq {

static thread_local std::string g_err;
int g_smem_optin = 227 * 1024;   // sharedMemPerBlockOptin of the device (set by mcq_create)

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            char _b[512];                                                                           \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return fail(MCQ_ECUDA, _b);                                                             \
        }                                                                                           \
    } while (0)

// ------------------------------------------------------------------------------------------
// geometry shared by host and device
// ------------------------------------------------------------------------------------------
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// idx = x*i + y*j + z*k + w for the 13 line families (oracle/queens_numpy.py: line_ids).
// Board mode has no (i,j) family; its counters start at family 1.
static int make_coefs(int full, int N, int4 coef[NFAM], int4 csel[NFAM]) {
    const int W = 2 * N - 1, o = N - 1;
    // space-diagonal families: the hexagon of valid (a, b) pairs folded into N rows of 3N-2 (anneal.cuh: line_index)
    const int R = 3 * N - 3, fold = 3 * N * N - N - 1;
    const int sz_axis = N * N, sz_plane = N * W, sz_space = N * (3 * N - 2);
    int base[NFAM];
    int b = 0;
    for (int f = 0; f < NFAM; ++f) {
        csel[f] = make_int4(0, 0, 0, 0);
        if (f == 0 && !full) { base[f] = 0; continue; }
        base[f] = b;
        b += f < 3 ? sz_axis : f < 9 ? sz_plane : sz_space;
    }
    coef[0] = make_int4(N, 1, 0, base[0]);
    coef[1] = make_int4(N, 0, 1, base[1]);
    coef[2] = make_int4(0, N, 1, base[2]);
    // k is the fastest index in every family that depends on it: a board move changes k only, so the old and
    // the new cell's counters of a family are at most N-1 bytes apart (same line / sector in the global-memory variant)
    coef[3] = make_int4(N, -N, 1, base[3] + o * N);   // (i-j+o)*N + k
    coef[4] = make_int4(N, N, 1, base[4]);            // (i+j)*N + k
    coef[5] = make_int4(1, W, -1, base[5] + o);
    coef[6] = make_int4(1, W, 1, base[6]);
    coef[7] = make_int4(W, 1, -1, base[7] + o);
    coef[8] = make_int4(W, 1, 1, base[8]);
    // idx = a R + b + o (+ fold when a < 0) with (a, b) = (i-j, i-k), (i-j, i+k-o), (i+j-o, i-k), (i+j-o, i+k-o)
    coef[9] = make_int4(R + 1, -R, -1, base[9] + o);
    coef[10] = make_int4(R + 1, -R, 1, base[10]);
    coef[11] = make_int4(R + 1, R, -1, base[11] + o - R * o);
    coef[12] = make_int4(R + 1, R, 1, base[12] - R * o);
    csel[9] = csel[10] = make_int4(1, -1, fold, 0);
    csel[11] = csel[12] = make_int4(1, 1, fold, -o);
    return b;  // total counter bytes
}

static Layout make_layout(int full, int N, int Q, int G) {
    Layout L;
    int4 tmp[NFAM], tmp2[NFAM];
    L.n_cnt = round_up(make_coefs(full, N, tmp, tmp2), 4);
    L.pos32 = (full && N > 32) ? 1 : 0;
    L.off_state = L.n_cnt;
    int state_b = full ? Q * (L.pos32 ? 4 : 2) : N * N;
    L.off_occ = round_up(L.off_state + state_b, 4);
    int occ_b = full ? round_up((N * N * N + 31) / 32 * 4, 4) : 0;
    L.off_pkt = round_up(L.off_occ + occ_b, 16);
    L.pb = G < 8 ? G : 8;
    L.off_hst = L.off_pkt + L.pb * PKT_BYTES;
    L.off_jrn = L.off_hst + HBLK * 4;
    L.stride = round_up(L.off_jrn + (G == 1 ? (JRN + 2) * 2 : 0), 16);
    return L;
}

static SLayout make_spec_layout(int full, int N, int Q) { return spec_layout(full, N, Q); }

// board: uint8 entries hold at most 12*N; full_3d: uint16 entries, bounded by the neighbour-row length
// the kernels are compiled for (13*(N-1) <= 256) and the 16-bit cell / wide ids
static inline bool spec_eligible(int full, int N) { return full ? 13 * (N - 1) + 1 <= 256 : 12 * N <= 255; }

#ifndef MCQ_GMEM_MIN_CHAINS_PER_SM
#define MCQ_GMEM_MIN_CHAINS_PER_SM 8   // below this many shared-memory chains per SM the global-memory kernel takes over
#endif

static inline int state_bytes_of(int mode, int n, int q) { return mode == MCQ_MODE_FULL3D ? 3 * q : n * n; }

// ------------------------------------------------------------------------------------------
// small kernels: initial states, energies, delta probes, cross-replica statistics
// ------------------------------------------------------------------------------------------
__host__ __device__ static inline int gcd_int(int a, int b) {
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}

// One thread per chain.  Distribution-identical to mcmc_board.py:26-59 / mcmc.py:20-101 (the
// reference consumes MT19937; we consume Philox in the INIT domain of the chain's key).
__global__ void init_states_kernel(int full, int N, int Q, int init_mode, int n_chains,
                                   const unsigned long long *seeds, uint8_t *state, int state_bytes,
                                   uint32_t *occ_scratch, int occ_words) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const unsigned long long sd = seeds ? seeds[chain] : 0ull;
    const uint32_t k0 = (uint32_t)sd, k1 = (uint32_t)(sd >> 32);
    uint8_t *st = state + (size_t)chain * state_bytes;
    uint32_t ctr = 0;
    Philox4 r;
    int have = 0;
    auto next_word = [&]() -> uint32_t {
        if (have == 0) { r = chain_words(ctr++, k0, k1, PHILOX_STREAM_INIT); have = 4; }
        const uint32_t v = have == 4 ? r.x : have == 3 ? r.y : have == 2 ? r.z : r.w;
        --have;
        return v;
    };
    int M = N;  // Klarner core edge
    if (init_mode == MCQ_INIT_KLARNER && gcd_int(N, 210) != 1) {
        for (M = N - 1; M > 0; --M) if (gcd_int(M, 210) == 1) break;
    }
    if (!full) {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                int k;
                if (init_mode == MCQ_INIT_LATIN) k = (i + j) % N;
                else if (init_mode == MCQ_INIT_KLARNER && i < M && j < M) k = (3 * i + 5 * j) % M;
                else k = (int)__umulhi(next_word(), (uint32_t)N);
                st[i * N + j] = (uint8_t)k;
            }
        return;
    }
    uint32_t *occ = occ_scratch + (size_t)chain * occ_words;
    for (int w = 0; w < occ_words; ++w) occ[w] = 0u;
    int placed = 0;
    if (init_mode == MCQ_INIT_LATIN || init_mode == MCQ_INIT_KLARNER) {
        const int edge = init_mode == MCQ_INIT_LATIN ? N : M;
        for (int i = 0; i < edge; ++i)
            for (int j = 0; j < edge; ++j) {
                const int k = init_mode == MCQ_INIT_LATIN ? (i + j) % N : (3 * i + 5 * j) % M;
                st[3 * placed] = (uint8_t)i; st[3 * placed + 1] = (uint8_t)j; st[3 * placed + 2] = (uint8_t)k;
                const int cid = (i * N + j) * N + k;
                occ[cid >> 5] |= 1u << (cid & 31);
                ++placed;
            }
    }
    while (placed < Q) {  // uniformly random unused cells, in order (== choice(replace=False))
        uint32_t w = next_word();
        const int i = draw_digit(w, N), j = draw_digit(w, N), k = draw_digit(w, N);
        const int cid = (i * N + j) * N + k;
        if ((occ[cid >> 5] >> (cid & 31)) & 1u) continue;
        occ[cid >> 5] |= 1u << (cid & 31);
        st[3 * placed] = (uint8_t)i; st[3 * placed + 1] = (uint8_t)j; st[3 * placed + 2] = (uint8_t)k;
        ++placed;
    }
}

// Explicit states are validated like the reference's constructors do (heights in range,
// mcmc_board.py:64-65; cells in range and pairwise distinct, mcmc.py:113-118): one thread per state,
// duplicates found with a scratch occupancy bitset.  *bad counts the offending states.
__global__ void validate_states_kernel(int full, int N, int Q, int n_states, const uint8_t *state, int state_bytes,
                                       uint32_t *occ_scratch, int occ_words, unsigned *bad) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_states) return;
    const uint8_t *st = state + (size_t)b * state_bytes;
    bool ok = true;
    if (!full) {
        for (int c = 0; c < N * N; ++c) ok = ok && st[c] < N;
    } else {
        uint32_t *occ = occ_scratch + (size_t)b * occ_words;
        for (int w = 0; w < occ_words; ++w) occ[w] = 0u;
        for (int q = 0; q < Q && ok; ++q) {
            const int i = st[3 * q], j = st[3 * q + 1], k = st[3 * q + 2];
            ok = i < N && j < N && k < N;
            if (ok) {
                const int cid = (i * N + j) * N + k;
                ok = !((occ[cid >> 5] >> (cid & 31)) & 1u);
                occ[cid >> 5] |= 1u << (cid & 31);
            }
        }
    }
    if (!ok) atomicAdd(bad, 1u);
}

// one warp per state: E = sum over attack lines of C(count,2)
__global__ void __launch_bounds__(32) energy_kernel(const __grid_constant__ KArgs a, int *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    const int e = build_chain<32>(a, smem, a.state + (size_t)b * a.state_bytes, true, threadIdx.x);
    if (threadIdx.x == 0) out[b] = e;
}

// one warp per state: delta-E of each candidate move with the same LineEval the chain uses
__global__ void __launch_bounds__(32) delta_kernel(const __grid_constant__ KArgs a, int n_moves, const uint32_t *moves,
                                                    int *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x, g = threadIdx.x;
    build_chain<32>(a, smem, a.state + (size_t)b * a.state_bytes, true, g);
    LaneLines<32> L;
    L.init(a, g);
    const unsigned char *st = smem + a.lay.off_state;
    for (int m = 0; m < n_moves; ++m) {
        const uint32_t w = moves[(size_t)b * n_moves + m];
        int i0, j0, k0, i1, j1, k1;
        bool bad;
        if (a.full) {
            const int q = w & 0xfff;
            i1 = (w >> 12) & 63; j1 = (w >> 18) & 63; k1 = (w >> 24) & 63;
            bad = q >= a.Q || i1 >= a.N || j1 >= a.N || k1 >= a.N;
            unpack_pos(a.lay.pos32, load_pos(st, a.lay.pos32, bad ? 0 : q), i0, j0, k0);
        } else {
            i0 = i1 = w & 255; j0 = j1 = (w >> 8) & 255; k1 = (w >> 16) & 255;
            bad = i0 >= a.N || j0 >= a.N || k1 >= a.N;
            k0 = st[bad ? 0 : i0 * a.N + j0];
        }
        if (bad) {   // out-of-range move: flagged, never evaluated
            if (g == 0) out[(size_t)b * n_moves + m] = INT_MIN;
            continue;
        }
        LineEval<32> ev;
        const int d = group_sum<32>(ev.eval(L, smem, i0, j0, k0, i1, j1, k1));
        if (g == 0) out[(size_t)b * n_moves + m] = d;
    }
}

// Cross-replica statistics (experiments.py:593-595: mean and std of the energy over the runs, per step).  The chain
// kernels deposit DIFFERENCES: column h of a group's row receives what its chains add to sum E / sum E^2 / the
// live-chain count when they reach history index h (an accepted move, the initial energy at h = 0, an early
// stop).  This kernel integrates columns [h0, h1] of every row in place, with the value of column h0-1 as the
// carry (0 when h0 == 0), so that the rows hold the sums themselves.  One CTA per row, tiles of 4096 columns.
template <typename T>
__global__ void __launch_bounds__(1024) stat_integrate_kernel(T *rows, long long pitch, long long h0, long long h1) {
    __shared__ T warp_tot[32];
    __shared__ T carry_s;
    T *row = rows + (size_t)blockIdx.x * pitch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = h0 > 0 ? row[h0 - 1] : (T)0;
    __syncthreads();
    for (long long base = h0; base <= h1; base += 4096) {
        T v[4];
        T run = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long h = base + 4 * tid + e;
            v[e] = h <= h1 ? row[h] : (T)0;
            run += v[e];
            v[e] = run;
        }
        T inc = run;   // inclusive scan of the threads' totals: warp, then across warps
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            T w = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const T before = carry_s + (warp > 0 ? warp_tot[warp - 1] : (T)0) + (inc - run);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long h = base + 4 * tid + e;
            if (h <= h1) row[h] = before + v[e];
        }
        __syncthreads();
        if (tid == 1023) carry_s = before + run;
        __syncthreads();
    }
}

// device build of the generator, for known-answer tests of the compiled device code
__global__ void philox_kat_kernel(int n, const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

__global__ void philox2_kat_kernel(int n, const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // through board_step_words, so that both compiled forms (constant key / key in a register) are the ones tested
    const Philox2 r = board_step_words(ctr[2 * i], ctr[2 * i + 1], key[i] ^ CHAIN_KEY0);
    out[2 * i] = r.x; out[2 * i + 1] = r.z;
}

// float64 beta(step) of parametrised schedules (what metropolis_exact evaluates on demand)
__global__ void beta64_table_kernel(const SchedDev *sched, int n_groups, int n_steps, double *out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_groups * n_steps) return;
    const int g = (int)(idx / n_steps);
    out[idx] = sched_beta64(sched[g], n_steps, (int)(idx - (long long)g * n_steps));
}

// ------------------------------------------------------------------------------------------
// context: device, streams, grow-only scratch
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// events of one call, destroyed on every return path
struct EventBag {
    std::vector<cudaEvent_t> ev;
    cudaError_t make(cudaEvent_t *out, unsigned flags) {
        cudaError_t e = cudaEventCreateWithFlags(out, flags);
        if (e == cudaSuccess) ev.push_back(*out);
        return e;
    }
    ~EventBag() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};

enum BufId {
    B_SEEDS, B_GROUP, B_BETA, B_INIT, B_RMOVES, B_RUNIF, B_STATE, B_BEST, B_REC, B_OCC, B_HIST0, B_HIST1,
    B_ABITS, B_ACCH, B_BINS, B_STATE_IN, B_MOVES, B_OUT, B_SUME, B_SUME2, B_SCNT, B_GSLAB, B_SCHED, B_BETA64, B_NBUF
};

}  // namespace mcq

struct mcq_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t copy_stream;
    cudaStream_t sub_stream[4];
    cudaDeviceProp prop;
    mcq::DevBuf buf[mcq::B_NBUF];
    std::map<int, mcq::DevBuf> nbr;   // neighbour lists per (mode, N), built on first use
    std::map<int, mcq::DevBuf> geo;   // shared-line bits + wide ids per N (full_3d)
};

namespace mcq {

static int check_problem(int mode, int n, int q) {
    if (mode != MCQ_MODE_BOARD && mode != MCQ_MODE_FULL3D) return fail(MCQ_EINVAL, "mode must be MCQ_MODE_BOARD or MCQ_MODE_FULL3D");
    if (n < 2 || n > 64) return fail(MCQ_EINVAL, "N must be in [2, 64]");
    if (mode == MCQ_MODE_BOARD && q != n * n) return fail(MCQ_EINVAL, "board mode requires Q == N*N");
    if (q < 1 || q > 4096) return fail(MCQ_EINVAL, "Q must be in [1, 4096]");
    if (mode == MCQ_MODE_FULL3D && (long long)q >= (long long)n * n * n) return fail(MCQ_EINVAL, "full_3d requires Q < N^3 (an empty cell to move to)");
    return 0;
}

// copy `bytes` from a caller buffer (host or device) into scratch `b`; returns device pointer
static int stage_in(mcq_ctx *ctx, int b, const void *src, size_t bytes, int mem, cudaStream_t s, void **out) {
    if (mem == MCQ_MEM_DEVICE) { *out = const_cast<void *>(src); return 0; }
    if (ctx->buf[b].ensure(bytes)) return fail(MCQ_ENOMEM, "device allocation failed");
    CUDA_TRY(cudaMemcpyAsync(ctx->buf[b].p, src, bytes, cudaMemcpyHostToDevice, s));
    *out = ctx->buf[b].p;
    return 0;
}

static int copy_out(void *dst, const void *src_dev, size_t bytes, int mem, cudaStream_t s) {
    if (!dst) return 0;
    CUDA_TRY(cudaMemcpyAsync(dst, src_dev, bytes, mem == MCQ_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    return 0;
}

// returns MCQ_EINVAL when a state is malformed (the reference raises ValueError in its constructors)
static int validate_states(mcq_ctx *ctx, int full, int n, int q, int n_states, const uint8_t *d_states, int sbytes,
                           uint32_t *d_counter, cudaStream_t s) {
    const int occ_words = full ? (n * n * n + 31) / 32 : 0;
    if (full && ctx->buf[B_OCC].ensure((size_t)n_states * occ_words * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
    CUDA_TRY(cudaMemsetAsync(d_counter, 0, 4, s));
    validate_states_kernel<<<(n_states + 127) / 128, 128, 0, s>>>(full, n, q, n_states, d_states, sbytes,
                                                                  static_cast<uint32_t *>(ctx->buf[B_OCC].p), occ_words, d_counter);
    CUDA_TRY(cudaGetLastError());
    unsigned bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, d_counter, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaMemsetAsync(d_counter, 0, 4, s));
    if (bad) return fail(MCQ_EINVAL, full ? "initial state: a cell is out of range or two queens occupy the same (i,j,k) cell"
                                          : "initial state: all heights must be in [0, N-1]");
    return 0;
}

}  // namespace mcq

using namespace mcq;

extern "C" {

int mcq_abi_version(void) { return MCQ_ABI_VERSION; }

int mcq_sizeof_run_params(void) { return (int)sizeof(mcq_run_params); }

const char *mcq_last_error(void) { return g_err.c_str(); }

int mcq_device_count(int *count) {
    if (!count) return fail(MCQ_EINVAL, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return fail(MCQ_ECUDA, cudaGetErrorString(e)); }
    return 0;
}

int mcq_create(int device, mcq_ctx **out) {
    if (!out) return fail(MCQ_EINVAL, "out is NULL");
    *out = nullptr;
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(MCQ_EINVAL, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    mcq_ctx *c = new mcq_ctx();
    c->device = device;
    CUDA_TRY(cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major < 10) {
        delete c;
        return fail(MCQ_ECUDA, "libmcq is built for sm_100a (B200) only; this device is older");
    }
    g_smem_optin = (int)c->prop.sharedMemPerBlockOptin;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto &ss : c->sub_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
    *out = c;
    return 0;
}

int mcq_destroy(mcq_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    for (auto &b : ctx->buf) b.release();
    for (auto &kv : ctx->nbr) kv.second.release();
    for (auto &kv : ctx->geo) kv.second.release();
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    for (auto ss : ctx->sub_stream) cudaStreamDestroy(ss);
    delete ctx;
    return 0;
}

int mcq_device_info(mcq_ctx *ctx, int *sm_count, int *smem_per_sm, int *smem_per_block_optin, int *clock_khz,
                    char *name, int name_len) {
    if (!ctx) return fail(MCQ_EINVAL, "ctx is NULL");
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (smem_per_sm) *smem_per_sm = (int)ctx->prop.sharedMemPerMultiprocessor;
    if (smem_per_block_optin) *smem_per_block_optin = (int)ctx->prop.sharedMemPerBlockOptin;
    if (clock_khz) *clock_khz = ctx->prop.clockRate;
    if (name && name_len > 0) { strncpy(name, ctx->prop.name, name_len - 1); name[name_len - 1] = 0; }
    return 0;
}

int mcq_state_bytes(int mode, int n, int q) {
    if (check_problem(mode, n, q)) return MCQ_EINVAL;
    return state_bytes_of(mode, n, q);
}

int mcq_chain_smem_bytes(int mode, int n, int q, int lanes_per_chain) {
    if (check_problem(mode, n, q)) return MCQ_EINVAL;
    if (lanes_per_chain != 4 && lanes_per_chain != 8 && lanes_per_chain != 16 && lanes_per_chain != 32)
        return fail(MCQ_EINVAL, "lanes_per_chain must be 4, 8, 16 or 32");
    return make_layout(mode == MCQ_MODE_FULL3D, n, q, lanes_per_chain).stride;
}

int mcq_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(MCQ_EINVAL, "ptr is NULL");
    CUDA_TRY(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable));   // usable from every device of the box
    return 0;
}

int mcq_host_free(void *ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return 0;
}

void mcq_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    const Philox4 r = philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int mcq_philox4x32_10_device(mcq_ctx *ctx, int n, const uint32_t *counters, const uint32_t *keys, uint32_t *out) {
    if (!ctx || n < 0 || (n && (!counters || !keys || !out))) return fail(MCQ_EINVAL, "bad arguments");
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (ctx->buf[B_MOVES].ensure((size_t)n * 16) || ctx->buf[B_OUT].ensure((size_t)n * 16) || ctx->buf[B_SEEDS].ensure((size_t)n * 8))
        return fail(MCQ_ENOMEM, "device allocation failed");
    CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_MOVES].p, counters, (size_t)n * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_SEEDS].p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    philox_kat_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, static_cast<const uint32_t *>(ctx->buf[B_MOVES].p),
                                                      static_cast<const uint32_t *>(ctx->buf[B_SEEDS].p), static_cast<uint32_t *>(ctx->buf[B_OUT].p));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, ctx->buf[B_OUT].p, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

void mcq_philox2x32_10(const uint32_t counter[2], uint32_t key, uint32_t out[2]) {
    const Philox2 r = philox2x32_10(counter[0], counter[1], key);
    out[0] = r.x; out[1] = r.z;
}

int mcq_philox2x32_10_device(mcq_ctx *ctx, int n, const uint32_t *counters, const uint32_t *keys, uint32_t *out) {
    if (!ctx || n < 0 || (n && (!counters || !keys || !out))) return fail(MCQ_EINVAL, "bad arguments");
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (ctx->buf[B_MOVES].ensure((size_t)n * 8) || ctx->buf[B_OUT].ensure((size_t)n * 8) || ctx->buf[B_SEEDS].ensure((size_t)n * 4))
        return fail(MCQ_ENOMEM, "device allocation failed");
    CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_MOVES].p, counters, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_SEEDS].p, keys, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    philox2_kat_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, static_cast<const uint32_t *>(ctx->buf[B_MOVES].p),
                                                       static_cast<const uint32_t *>(ctx->buf[B_SEEDS].p), static_cast<uint32_t *>(ctx->buf[B_OUT].p));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, ctx->buf[B_OUT].p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

int mcq_beta_table(mcq_ctx *ctx, int n_groups, const mcq_schedule *schedules, int n_steps, double *out_beta, float *out_c) {
    if (!ctx || n_groups < 1 || n_steps < 0 || !schedules) return fail(MCQ_EINVAL, "bad arguments");
    for (int g = 0; g < n_groups; ++g)
        if (schedules[g].type < MCQ_SCHED_CONSTANT || schedules[g].type > MCQ_SCHED_SINUSOIDAL) return fail(MCQ_EINVAL, "unknown schedule type");
    if (n_steps == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const long long cells = (long long)n_groups * n_steps;
    if (ctx->buf[B_SCHED].ensure((size_t)n_groups * sizeof(SchedDev)) || ctx->buf[B_BETA64].ensure((size_t)cells * 8) ||
        ctx->buf[B_BETA].ensure((size_t)cells * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
    CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_SCHED].p, schedules, (size_t)n_groups * sizeof(SchedDev), cudaMemcpyHostToDevice, s));
    const SchedDev *sd = static_cast<const SchedDev *>(ctx->buf[B_SCHED].p);
    const unsigned grid = (unsigned)((cells + 255) / 256);
    if (out_beta) {
        beta64_table_kernel<<<grid, 256, 0, s>>>(sd, n_groups, n_steps, static_cast<double *>(ctx->buf[B_BETA64].p));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_beta, ctx->buf[B_BETA64].p, (size_t)cells * 8, cudaMemcpyDeviceToHost, s));
    }
    if (out_c) {
        beta_table_kernel<<<grid, 256, 0, s>>>(sd, nullptr, n_groups, n_steps, static_cast<float *>(ctx->buf[B_BETA].p));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_c, ctx->buf[B_BETA].p, (size_t)cells * 4, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

static int probe_common(mcq_ctx *ctx, int mode, int n, int q, int n_states, const uint8_t *states, int mem,
                        cudaStream_t s, KArgs &a) {
    if (!ctx) return fail(MCQ_EINVAL, "ctx is NULL");
    if (int rc = check_problem(mode, n, q)) return rc;
    if (n_states < 0 || (!states && n_states)) return fail(MCQ_EINVAL, "states is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    memset(&a, 0, sizeof a);
    a.full = mode == MCQ_MODE_FULL3D;
    a.N = n; a.Q = q; a.n_chains = n_states;
    a.lay = make_layout(a.full, n, q, 32);
    make_coefs(a.full, n, a.coef, a.csel);
    a.state_bytes = state_bytes_of(mode, n, q);
    if ((size_t)a.lay.stride > ctx->prop.sharedMemPerBlockOptin) return fail(MCQ_ENOMEM, "one chain does not fit in shared memory");
    void *d = nullptr;
    if (int rc = stage_in(ctx, B_STATE_IN, states, (size_t)n_states * a.state_bytes, mem, s, &d)) return rc;
    a.state = static_cast<uint8_t *>(d);
    if (n_states > 0) {
        if (ctx->buf[B_REC].ensure(64)) return fail(MCQ_ENOMEM, "device allocation failed");
        if (int rc = validate_states(ctx, a.full, n, q, n_states, a.state, a.state_bytes, static_cast<uint32_t *>(ctx->buf[B_REC].p), s)) return rc;
    }
    return 0;
}

int mcq_energy(mcq_ctx *ctx, int mode, int n, int q, int n_states, const uint8_t *states, int32_t *out_energy,
               int mem, void *stream) {
    if (!out_energy && n_states) return fail(MCQ_EINVAL, "out_energy is NULL");
    cudaStream_t s = stream ? (cudaStream_t)stream : (ctx ? ctx->stream : nullptr);
    KArgs a;
    if (int rc = probe_common(ctx, mode, n, q, n_states, states, mem, s, a)) return rc;
    if (n_states == 0) return 0;
    int *d_out = out_energy;
    if (mem == MCQ_MEM_HOST) {
        if (ctx->buf[B_OUT].ensure((size_t)n_states * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
        d_out = static_cast<int *>(ctx->buf[B_OUT].p);
    }
    CUDA_TRY(cudaFuncSetAttribute(energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    energy_kernel<<<n_states, 32, a.lay.stride, s>>>(a, d_out);
    CUDA_TRY(cudaGetLastError());
    if (mem == MCQ_MEM_HOST) CUDA_TRY(cudaMemcpyAsync(out_energy, d_out, (size_t)n_states * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

int mcq_delta_energy(mcq_ctx *ctx, int mode, int n, int q, int n_states, const uint8_t *states, int n_moves,
                     const uint32_t *moves, int32_t *out_delta, int mem, void *stream) {
    if ((!out_delta || !moves) && n_states && n_moves) return fail(MCQ_EINVAL, "moves/out_delta is NULL");
    cudaStream_t s = stream ? (cudaStream_t)stream : (ctx ? ctx->stream : nullptr);
    KArgs a;
    if (int rc = probe_common(ctx, mode, n, q, n_states, states, mem, s, a)) return rc;
    if (n_states == 0 || n_moves <= 0) return 0;
    const size_t cnt = (size_t)n_states * n_moves;
    void *d_moves = nullptr;
    if (int rc = stage_in(ctx, B_MOVES, moves, cnt * 4, mem, s, &d_moves)) return rc;
    int *d_out = out_delta;
    if (mem == MCQ_MEM_HOST) {
        if (ctx->buf[B_OUT].ensure(cnt * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
        d_out = static_cast<int *>(ctx->buf[B_OUT].p);
    }
    CUDA_TRY(cudaFuncSetAttribute(delta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    delta_kernel<<<n_states, 32, a.lay.stride, s>>>(a, n_moves, static_cast<const uint32_t *>(d_moves), d_out);
    CUDA_TRY(cudaGetLastError());
    if (mem == MCQ_MEM_HOST) CUDA_TRY(cudaMemcpyAsync(out_delta, d_out, cnt * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

int mcq_run(mcq_ctx *ctx, const mcq_run_params *p) {
    if (!ctx || !p) return fail(MCQ_EINVAL, "ctx/params is NULL");
    if (p->struct_size != sizeof(mcq_run_params)) return fail(MCQ_EINVAL, "mcq_run_params size mismatch (ABI)");
    if (int rc = check_problem(p->mode, p->n, p->q)) return rc;
    if (p->n_steps < 0) return fail(MCQ_EINVAL, "n_steps must be >= 0");
    if (p->n_chains < 0) return fail(MCQ_EINVAL, "n_chains must be >= 0");
    if (p->n_groups < 1) return fail(MCQ_EINVAL, "n_groups must be >= 1");
    if (p->init_mode < MCQ_INIT_RANDOM || p->init_mode > MCQ_INIT_EXPLICIT) return fail(MCQ_EINVAL, "unknown init_mode");
    if (p->init_mode == MCQ_INIT_EXPLICIT && !p->init_states && p->n_chains) return fail(MCQ_EINVAL, "init_states required for MCQ_INIT_EXPLICIT");
    if ((p->init_mode == MCQ_INIT_LATIN || p->init_mode == MCQ_INIT_KLARNER) && p->q != p->n * p->n)
        return fail(MCQ_EINVAL, "latin/klarner initialization assumes Q = N^2");
    const bool replay = p->replay_moves != nullptr;
    if (replay && (!p->replay_uniforms || !p->beta_f64)) return fail(MCQ_EINVAL, "replay needs replay_moves, replay_uniforms and beta_f64");
    if (!replay && !p->schedules && !p->beta_f64 && p->n_steps > 0) return fail(MCQ_EINVAL, "a production run needs `schedules` (parameters) or `beta_f64` (a tabulated schedule)");
    if (!replay && p->schedules)
        for (int g = 0; g < p->n_groups; ++g)
            if (p->schedules[g].type < MCQ_SCHED_CONSTANT || p->schedules[g].type > MCQ_SCHED_SINUSOIDAL) return fail(MCQ_EINVAL, "unknown schedule type");
    if (p->mem != MCQ_MEM_HOST && p->mem != MCQ_MEM_DEVICE) return fail(MCQ_EINVAL, "mem must be MCQ_MEM_HOST or MCQ_MEM_DEVICE");
    const bool want_hist = p->hist_dtype != MCQ_HIST_NONE && p->energy_history;
    const bool want_stats = p->stat_sum_e || p->stat_sum_e2;
    if (want_stats && !(p->stat_sum_e && p->stat_sum_e2)) return fail(MCQ_EINVAL, "stat_sum_e and stat_sum_e2 go together");
    if (p->hist_dtype < MCQ_HIST_NONE || p->hist_dtype > MCQ_HIST_I32) return fail(MCQ_EINVAL, "unknown hist_dtype");
    if (want_hist && p->hist_pitch < (int64_t)p->n_steps + 1) return fail(MCQ_EINVAL, "hist_pitch must be >= n_steps+1");
    const long long e_max = 13LL * p->q * (p->n - 1) / 2;
    if (want_hist && p->hist_dtype == MCQ_HIST_U16 && e_max >= 65536) return fail(MCQ_EINVAL, "uint16 history would overflow for this N; use MCQ_HIST_I32");
    if (p->n_bins < 0 || (p->n_bins > 0 && (!p->bin_starts || !p->accept_hist))) return fail(MCQ_EINVAL, "n_bins > 0 needs bin_starts and accept_hist");
    // checkpoint / resume: this call executes steps [t_start, t_stop) of the schedule
    const int t_start = p->start_step, t_stop = p->stop_step > 0 ? p->stop_step : p->n_steps;
    if (t_start < 0 || t_start > t_stop || t_stop > p->n_steps) return fail(MCQ_EINVAL, "need 0 <= start_step <= stop_step <= n_steps");
    if (t_start % HBLK != 0 || (t_stop % HBLK != 0 && t_stop != p->n_steps))
        return fail(MCQ_EINVAL, "start_step and stop_step must be multiples of 32 (stop_step may also be n_steps)");
    const bool resume = t_start > 0;
    if (resume && (p->init_mode != MCQ_INIT_EXPLICIT || !p->resume_record || !p->resume_best_state))
        return fail(MCQ_EINVAL, "start_step > 0 needs init_states (the states at start_step), resume_record and resume_best_state");
    if (p->n_chains == 0) { if (p->kernel_ms) *p->kernel_ms = 0.f; if (p->gpu_launches) *p->gpu_launches = 0; return 0; }

    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = p->stream ? (cudaStream_t)p->stream : ctx->stream;
    const int mem = p->mem;
    const int full = p->mode == MCQ_MODE_FULL3D;
    const int nc = p->n_chains, ns = p->n_steps;
    const int sbytes = state_bytes_of(p->mode, p->n, p->q);
    int launches = 0;

    // ---- geometry ----
    const size_t smem_block = ctx->prop.sharedMemPerBlockOptin;
    const size_t smem_sm = ctx->prop.sharedMemPerMultiprocessor;
    int G = p->lanes_per_chain;
    if (G != 0 && G != 4 && G != 8 && G != 16 && G != 32) return fail(MCQ_EINVAL, "lanes_per_chain must be 0, 4, 8, 16 or 32");
    if (p->algo < MCQ_ALGO_AUTO || p->algo > MCQ_ALGO_WIDE) return fail(MCQ_EINVAL, "unknown algo");
    if (p->algo == MCQ_ALGO_TABLE && !spec_eligible(full, p->n)) return fail(MCQ_EINVAL, "MCQ_ALGO_TABLE serves N <= 20 (full_3d) or N <= 21 (board)");
    // full_3d N = 19, 20: the 16-bit table leaves few chains per SM; the CTA-per-chain kernel is 11-24 % ahead (measured)
    const bool table_pays = spec_eligible(full, p->n) && !(full && p->n >= 19 && !replay);
    const bool use_spec = p->algo == MCQ_ALGO_TABLE || (p->algo == MCQ_ALGO_AUTO && G == 0 && table_pays);
    // Boards whose line counters leave room for only a few chains per SM run one thread per chain with the
    // counters in global memory (HBM-bound byte traffic instead of a latency-bound handful of warps).
    bool use_gmem = !use_spec && (p->algo == MCQ_ALGO_GMEM ||
        (p->algo == MCQ_ALGO_AUTO && G == 0 && (size_t)make_layout(full, p->n, p->q, 32).stride * MCQ_GMEM_MIN_CHAINS_PER_SM > smem_sm));
    // One CTA per chain on shared-memory line counters (wide.cuh); production runs only.  It is the default beyond
    // the conflict table's reach: 3-20x faster than a warp or a thread per chain when the chains are few, ahead on
    // long anneals at any count (a cold chain retires ~60 proposals per round), and -- since the folded
    // space-diagonal counters let an SM hold two N = 64 chains -- level with the global-memory kernel even on
    // short hot runs of very many chains (65536 chains x 1e5 steps at N = 64: 2.5e9 proposals/s either way).
    bool use_wide = !use_spec && p->algo == MCQ_ALGO_WIDE;
    if (use_wide && replay) return fail(MCQ_EINVAL, "MCQ_ALGO_WIDE does not replay recorded streams");
    {
        const Layout l1 = make_layout(full, p->n, p->q, 1);
        const size_t need = (size_t)l1.off_pkt + 2 * 32 * 16 + WIDE_XCH_BYTES + 32 * (full ? 4 : 2) * 4;   // narrowest CTA
        if (p->algo == MCQ_ALGO_AUTO && !use_spec && !replay && G == 0 && need <= smem_block) use_wide = true;
    }
    if (use_wide) use_gmem = false;
    if (use_gmem || use_wide) G = 1;
    if (G == 0) {
        G = 8;
        while (G < 32 && (size_t)make_layout(full, p->n, p->q, G).stride * (32 / G) > smem_block) G *= 2;
    }
    Layout lay = make_layout(full, p->n, p->q, G);
    // CTA-per-chain kernel: 8, 4, 2 or 1 warps per chain.  What counts first is how many chains an SM holds (shared
    // memory: counters + state + a ring of 2 * threads steps; registers: 144 per thread).  Among the widths that
    // reach that residency: one warp up to N = 34 (no block barrier at all), two warps beyond (measured with the
    // multi-commit rounds, board / full_3d: N = 28 13 % and N = 32 4 % / 5 % faster on one warp, N = 36 2 %, N = 40
    // 9 % / 12 % and N = 44..64 8-30 % faster on two; four and eight warps are slower everywhere).
    // warps_per_cta = 1, 2, 4, 8 overrides.
    const int w_best = lay.off_pkt, w_ring = lay.off_pkt;   // (the best state is kept in global memory: no shared copy)
    int wide_threads = WIDE_THREADS;
    auto wide_bytes = [&](int nt) { return (size_t)w_ring + 2 * (size_t)nt * 16 + WIDE_XCH_BYTES + (size_t)nt * (full ? 4 : 2) * 4; };   // + a record per thread (multi-commit rounds)
    if (use_wide) {
        const long long want = (nc + ctx->prop.multiProcessorCount - 1) / ctx->prop.multiProcessorCount;   // chains per SM on offer
        long long best_conc = 0;
        const int preferred = p->n <= 34 ? 32 : 64;
        auto residency = [&](int nt) -> long long {
            if (wide_bytes(nt) > smem_block) return 0;
            const long long by_smem = (long long)(smem_sm / (wide_bytes(nt) + 1024)), by_regs = 65536 / (144 * nt);
            return std::min(want, std::min<long long>(32, std::min(by_smem, by_regs)));
        };
        for (int nt = 32; nt <= 256; nt <<= 1) best_conc = std::max(best_conc, residency(nt));   // (non-increasing in nt)
        for (int nt = 32; nt <= preferred; nt <<= 1)
            if (residency(nt) == best_conc) wide_threads = nt;                                    // the widest up to the preferred one
        if (best_conc == 0) return fail(MCQ_ENOMEM, "the line counters of one chain do not fit in shared memory");
        if (p->warps_per_cta == 1 || p->warps_per_cta == 2 || p->warps_per_cta == 4 || p->warps_per_cta == 8) wide_threads = p->warps_per_cta * 32;
        if (const char *e = getenv("MCQ_WIDE_THREADS")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) wide_threads = v; }
        if (wide_bytes(wide_threads) > smem_block) return fail(MCQ_ENOMEM, "the line counters of one chain do not fit in shared memory at this CTA width");
    }
    const int w_xch = w_ring + 2 * wide_threads * 16;
    const size_t wide_smem = wide_bytes(wide_threads);
    if (!use_gmem && !use_wide && (size_t)lay.stride * (32 / G) > smem_block) return fail(MCQ_ENOMEM, "a warp's chains do not fit in shared memory; raise lanes_per_chain");
    int wpc = p->warps_per_cta;
    if (wpc < 0 || wpc > 8) return fail(MCQ_EINVAL, "warps_per_cta must be in [0, 8]");
    int best_w = 1, best_chains = 0, best_ctas = 1;
    for (int w = 8; w >= 1 && !use_gmem && !use_wide; --w) {
        if (wpc && w != wpc) continue;
        const int cpc_w = w * 32 / G;
        const size_t need = (size_t)cpc_w * lay.stride;
        if (need > smem_block) continue;
        int ctas = (int)std::min<size_t>(32, smem_sm / (need + 1024));
        ctas = std::min(ctas, 64 / w);
        if (ctas * cpc_w > best_chains) { best_chains = ctas * cpc_w; best_w = w; best_ctas = ctas; }
    }
    if (best_chains == 0 && !use_gmem && !use_wide) return fail(MCQ_ENOMEM, "requested warps_per_cta does not fit in shared memory");
    int cpc = use_gmem ? 128 : use_wide ? 1 : best_w * 32 / G;
    int block = use_gmem ? 128 : use_wide ? wide_threads : best_w * 32;
    int grid = (nc + cpc - 1) / cpc;
    size_t smem = use_gmem ? 0 : use_wide ? wide_smem : (size_t)cpc * lay.stride;
    const SLayout sl = make_spec_layout(full, p->n, p->q);
    int spec_lpc = 32;
    if (use_spec) {
        // Lane groups of 16 or 32 per chain.  Registers (64/thread) allow 32 warps per SM; shared memory may allow
        // fewer.  All CTAs of a launch take the same time, so a launch costs ceil(waves) full waves: pick the
        // (group width, resident CTAs per SM) pair with the best modelled rate x wave efficiency.
        // (fast_serves needs the launch arguments; the geometry decisions below only need to know which kernel it will be)
        KArgs probe;
        memset(&probe, 0, sizeof probe);
        probe.full = full; probe.N = p->n; probe.Q = p->q; probe.patience = (!full && p->early_stop_patience >= 0) ? p->early_stop_patience : -1;
        probe.hist_kind = want_hist ? p->hist_dtype : MCQ_HIST_NONE;
        probe.dsum_e = want_stats ? reinterpret_cast<unsigned long long *>(1) : nullptr;
        probe.stat_rows32 = (long long)p->n_groups * ((long long)ns + 1) < (1LL << 31);
        const int lpc_req = (p->algo == MCQ_ALGO_TABLE && p->lanes_per_chain == 16) ? 16 : 32;
        const bool fast = fast_serves(lpc_req, probe, replay);
        const int w_max = fast && lpc_req == 32 ? MCQ_FAST_WARPS : 4, w_def = w_max;
        const int minb = lpc_req != 32 ? 5 : fast ? MCQ_FAST_MINB : MCQ_SPEC_MINB;
        if (wpc > w_max) return fail(MCQ_EINVAL, "the conflict-table kernel does not run that many warps per CTA");
        const int w = wpc ? wpc : w_def;
        const int sms = ctx->prop.multiProcessorCount;
        auto cta_smem = [&](int lpc) { return (size_t)sl.cta_bytes + (size_t)w * (32 / lpc) * sl.stride; };
        auto max_ctas = [&](int lpc) {   // residency limit: registers, warps, shared memory
            const size_t need = cta_smem(lpc);
            if (need > smem_block) return 0;
            // registers: the kernel is compiled for MCQ_SPEC_MINB CTAs of 4 warps per SM
            return (int)std::min<size_t>((size_t)(minb * w_max / w), smem_sm / (round_up((int)need, 1024) + 1024));
        };
        // relative throughput of one SM vs resident warps (measured, N=12 full_3d, single full wave)
        auto rate = [&](int lpc, int warps) {
            static const float r32[] = {0.0f, 0.40f, 0.80f, 0.967f, 1.0f};      // at 0, 8, 16, 24, 32 warps
            static const float r16[] = {0.0f, 0.43f, 0.86f, 1.09f, 1.185f};
            const float *r = lpc == 16 ? r16 : r32;
            const float x = std::min(32, std::max(0, warps)) / 8.0f;
            const int i = std::min(3, (int)x);
            return r[i] + (r[i + 1] - r[i]) * (x - i);
        };
        // One group of 32 lanes per chain unless asked otherwise: 16-lane groups retire ~12 % fewer
        // instructions per proposal but two slabs per warp cost residency, and measured throughput is equal.
        // CTA durations differ between schedules, so partial waves cost only ~5 %: take the full residency.
        int best_ctas_spec = 0;
        float best_score = 0.f;
        if (p->algo == MCQ_ALGO_TABLE && (p->lanes_per_chain == 16 || p->lanes_per_chain == 32)) spec_lpc = p->lanes_per_chain;
        best_ctas_spec = max_ctas(spec_lpc);
        if (p->max_chains_per_sm > 0) best_ctas_spec = std::max(1, std::min(best_ctas_spec, p->max_chains_per_sm / (w * (32 / spec_lpc))));
        best_score = rate(spec_lpc, best_ctas_spec * w);
        if (best_ctas_spec == 0) return fail(MCQ_ENOMEM, "conflict-table slab does not fit in shared memory; lower warps_per_cta");
        cpc = w * (32 / spec_lpc);
        block = w * 32;
        grid = (nc + cpc - 1) / cpc;
        smem = cta_smem(spec_lpc);
        if (best_ctas_spec < max_ctas(spec_lpc)) {   // cap residency by padding the request (1 KB allocation granules)
            size_t pad = (smem_sm / best_ctas_spec - 1024) / 1024 * 1024;
            smem = std::max(smem, std::min(pad, smem_block));
        }
        if (getenv("MCQ_DEBUG"))
            fprintf(stderr, "[mcq] table kernel: lanes/chain %d, %d warps/CTA, %d CTAs/SM (max %d), grid %d, smem %zu B, score %.3f\n",
                    spec_lpc, w, best_ctas_spec, max_ctas(spec_lpc), grid, smem, best_score);
    } else if (!use_gmem && !use_wide) {   // cap residency (explicit limit, or balance the waves) by padding the shared-memory request
        int ctas = best_ctas;
        const int sms = ctx->prop.multiProcessorCount;
        if (p->max_chains_per_sm > 0) ctas = std::max(1, std::min(ctas, p->max_chains_per_sm / cpc));
        else {
            const long long resident = (long long)sms * ctas;
            const long long waves = (grid + resident - 1) / resident;
            const int per_sm = (int)((grid + sms * waves - 1) / (sms * waves));
            ctas = std::max(1, std::min(ctas, per_sm));
        }
        if (ctas < best_ctas) {
            size_t pad = smem_sm / ctas - 1024;
            pad = std::min(pad, smem_block) & ~(size_t)15;
            smem = std::max(smem, pad);
        }
    }

    // ---- inputs ----
    KArgs a;
    memset(&a, 0, sizeof a);
    a.full = full; a.N = p->n; a.Q = p->q; a.n_chains = nc; a.n_steps = ns; a.t_last = t_stop;
    a.patience = (!full && p->early_stop_patience >= 0) ? p->early_stop_patience : -1;
    a.lay = lay;
    a.sl = sl;
    a.w_best = w_best; a.w_ring = w_ring; a.w_xch = w_xch;
    if (use_spec) {
        const int nbr_key = full * 256 + p->n;
        if (!ctx->nbr.count(nbr_key)) {   // registered only once it is allocated AND built
            const int cells = p->n * p->n * p->n;
            DevBuf nb;
            if (nb.ensure((size_t)cells * sl.nbr_len * 2)) return fail(MCQ_ENOMEM, "device allocation failed (neighbour lists)");
            if (ctx->buf[B_MOVES].ensure((size_t)cells * sl.nbr_len * 2)) { nb.release(); return fail(MCQ_ENOMEM, "device allocation failed (neighbour lists)"); }
            build_neighbours_kernel<<<(cells + 127) / 128, 128, 0, s>>>(full, p->n, sl.nbr_len, full ? 2 : 1, static_cast<uint16_t *>(nb.p),
                                                                       static_cast<uint16_t *>(ctx->buf[B_MOVES].p));
            if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) { nb.release(); return fail(MCQ_ECUDA, cudaGetErrorString(e)); }
            ctx->nbr[nbr_key] = nb;
            ++launches;
        }
        a.nbr = static_cast<const uint16_t *>(ctx->nbr[nbr_key].p);
        if (full) {
            DevBuf &gb = ctx->geo[p->n];
            if (!gb.p) {
                const int cells = p->n * p->n * p->n;
                if (gb.ensure((size_t)sl.cta_bytes)) return fail(MCQ_ENOMEM, "device allocation failed (geometry tables)");
                CUDA_TRY(cudaMemsetAsync(gb.p, 0, (size_t)sl.cta_bytes, s));
                const int lut_words = sl.off_wide / 4;
                const int n = std::max(cells, lut_words);
                build_wide_lut_kernel<<<(n + 127) / 128, 128, 0, s>>>(p->n, reinterpret_cast<uint16_t *>(static_cast<char *>(gb.p) + sl.off_wide),
                                                                     static_cast<uint32_t *>(gb.p), lut_words);
                CUDA_TRY(cudaGetLastError());
                ++launches;
            }
            a.geo = static_cast<const uint32_t *>(gb.p);
        }
    }
    make_coefs(full, p->n, a.coef, a.csel);
    a.state_bytes = sbytes;
    void *d = nullptr;
    if (p->chain_seeds) { if (int rc = stage_in(ctx, B_SEEDS, p->chain_seeds, (size_t)nc * 8, mem, s, &d)) return rc; a.seeds = static_cast<unsigned long long *>(d); }
    if (p->chain_group) { if (int rc = stage_in(ctx, B_GROUP, p->chain_group, (size_t)nc * 4, mem, s, &d)) return rc; a.group = static_cast<int *>(d); }
    if (replay) {
        if (int rc = stage_in(ctx, B_BETA, p->beta_f64, (size_t)p->n_groups * ns * 8, mem, s, &d)) return rc; a.beta64 = static_cast<double *>(d);
        if (int rc = stage_in(ctx, B_RMOVES, p->replay_moves, (size_t)nc * ns * 4, mem, s, &d)) return rc; a.rmoves = static_cast<uint32_t *>(d);
        if (int rc = stage_in(ctx, B_RUNIF, p->replay_uniforms, (size_t)nc * ns * 8, mem, s, &d)) return rc; a.runif = static_cast<double *>(d);
    } else if (ns > 0) {
        // production: the float32 table of -beta log2 e the fast path reads is built on the device, from the schedule
        // parameters (experiments.py:13-77 evaluated in float64) or from the caller's float64 table of a closure;
        // the float64 accept rule (accept.cuh) goes back to the same source
        if (p->schedules) {
            static_assert(sizeof(SchedDev) == sizeof(mcq_schedule), "schedule structs must match");
            if (ctx->buf[B_SCHED].ensure((size_t)p->n_groups * sizeof(SchedDev))) return fail(MCQ_ENOMEM, "device allocation failed");
            CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_SCHED].p, p->schedules, (size_t)p->n_groups * sizeof(SchedDev), cudaMemcpyHostToDevice, s));
            a.sched = static_cast<const SchedDev *>(ctx->buf[B_SCHED].p);
        } else {
            if (int rc = stage_in(ctx, B_BETA64, p->beta_f64, (size_t)p->n_groups * ns * 8, mem, s, &d)) return rc;
            a.beta64 = static_cast<double *>(d);
        }
        const long long cells = (long long)p->n_groups * ns;
        // (padded: the lanes of a round that reach past the end of a launch read, and discard, up to 32 entries more)
        if (ctx->buf[B_BETA].ensure((size_t)(cells + BETA_PAD) * 4)) return fail(MCQ_ENOMEM, "device allocation failed (schedule table)");
        CUDA_TRY(cudaMemsetAsync(static_cast<float *>(ctx->buf[B_BETA].p) + cells, 0, BETA_PAD * 4, s));
        beta_table_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(a.sched, a.beta64, p->n_groups, ns, static_cast<float *>(ctx->buf[B_BETA].p));
        CUDA_TRY(cudaGetLastError());
        ++launches;
        a.beta_c = static_cast<float *>(ctx->buf[B_BETA].p);
    }
    a.band_abs = p->accept_all_f64 ? INFINITY : 4.0f;

    // ---- persistent record (always internal scratch; copied to the caller's arrays at the end) ----
    // layout of B_REC: 10 int arrays of nc + 1 word for replay_err
    if (ctx->buf[B_REC].ensure(((size_t)nc * 10 + 4) * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
    int *rec = static_cast<int *>(ctx->buf[B_REC].p);
    a.init_e = rec; a.cur_e = rec + nc; a.best_e = rec + 2 * (size_t)nc; a.best_step = rec + 3 * (size_t)nc;
    a.n_acc = rec + 4 * (size_t)nc; a.steps_done = rec + 5 * (size_t)nc; a.stale = rec + 6 * (size_t)nc;
    a.bin_mark = rec + 7 * (size_t)nc; a.near_cnt = reinterpret_cast<uint32_t *>(rec + 8 * (size_t)nc);
    a.flip_cnt = reinterpret_cast<uint32_t *>(rec + 9 * (size_t)nc);
    a.replay_err = reinterpret_cast<uint32_t *>(rec + 10 * (size_t)nc);
    CUDA_TRY(cudaMemsetAsync(rec, 0, ((size_t)nc * 10 + 4) * 4, s));
    if (ctx->buf[B_STATE].ensure((size_t)nc * sbytes) || ctx->buf[B_BEST].ensure((size_t)nc * sbytes)) return fail(MCQ_ENOMEM, "device allocation failed");
    a.state = static_cast<uint8_t *>(ctx->buf[B_STATE].p);
    a.best_state = static_cast<uint8_t *>(ctx->buf[B_BEST].p);

    // ---- initial states ----
    if (p->init_mode == MCQ_INIT_EXPLICIT) {
        CUDA_TRY(cudaMemcpyAsync(a.state, p->init_states, (size_t)nc * sbytes,
                                 mem == MCQ_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        if (int rc = validate_states(ctx, full, p->n, p->q, nc, a.state, sbytes, a.replay_err, s)) return rc;
        ++launches;
    } else {
        const int occ_words = full ? (p->n * p->n * p->n + 31) / 32 : 0;
        if (full && ctx->buf[B_OCC].ensure((size_t)nc * occ_words * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
        init_states_kernel<<<(nc + 127) / 128, 128, 0, s>>>(full, p->n, p->q, p->init_mode, nc, a.seeds, a.state, sbytes,
                                                             static_cast<uint32_t *>(ctx->buf[B_OCC].p), occ_words);
        CUDA_TRY(cudaGetLastError());
        ++launches;
    }
    const cudaMemcpyKind in_kind = mem == MCQ_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (resume) {   // the record and the best states continue from the previous segment
        CUDA_TRY(cudaMemcpyAsync(rec, p->resume_record, (size_t)nc * MCQ_RECORD_INTS * 4, in_kind, s));
        CUDA_TRY(cudaMemcpyAsync(a.best_state, p->resume_best_state, (size_t)nc * sbytes, in_kind, s));
    } else {
        CUDA_TRY(cudaMemcpyAsync(a.best_state, a.state, (size_t)nc * sbytes, cudaMemcpyDeviceToDevice, s));
    }
    if (use_gmem) {   // slabs in global memory: built once, reused by every launch of this call
        if (ctx->buf[B_GSLAB].ensure(((size_t)nc + 1) * lay.stride)) return fail(MCQ_ENOMEM, "device allocation failed (global-memory slabs)");
        a.gslab = static_cast<unsigned char *>(ctx->buf[B_GSLAB].p);
        a.gslab_dummy = nc;
        CUDA_TRY(cudaMemsetAsync(a.gslab + (size_t)nc * lay.stride, 0, lay.stride, s));
        a.chain_begin = 0;
        a.t_begin = t_start;
        CUDA_TRY(launch_gslab_build(a, nc, s));
        ++launches;
    }

    // ---- acceptance bins ----
    a.n_bins = p->n_bins;
    if (p->n_bins > 0) {
        if (p->bin_starts[0] != 0 || p->bin_starts[p->n_bins] != ns) return fail(MCQ_EINVAL, "bin_starts must run from 0 to n_steps");
        for (int b = 0; b < p->n_bins; ++b) if (p->bin_starts[b] > p->bin_starts[b + 1]) return fail(MCQ_EINVAL, "bin_starts must be non-decreasing");
        if (ctx->buf[B_BINS].ensure((size_t)(p->n_bins + 2) * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
        std::vector<int> bs(p->bin_starts, p->bin_starts + p->n_bins + 1);
        bs.push_back(0x7fffffff);
        CUDA_TRY(cudaMemcpyAsync(ctx->buf[B_BINS].p, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaStreamSynchronize(s));  // bs is a stack temporary
        a.bin_starts = static_cast<int *>(ctx->buf[B_BINS].p);
        if (mem == MCQ_MEM_DEVICE) a.acc_hist = p->accept_hist;
        else {
            if (ctx->buf[B_ACCH].ensure((size_t)nc * p->n_bins * 4)) return fail(MCQ_ENOMEM, "device allocation failed");
            a.acc_hist = static_cast<uint32_t *>(ctx->buf[B_ACCH].p);
        }
        if (!resume) CUDA_TRY(cudaMemsetAsync(a.acc_hist, 0, (size_t)nc * p->n_bins * 4, s));
        else if (mem == MCQ_MEM_HOST) CUDA_TRY(cudaMemcpyAsync(a.acc_hist, p->accept_hist, (size_t)nc * p->n_bins * 4, cudaMemcpyHostToDevice, s));
    }

    // ---- accept bitmap ----
    const long long abits_words = ((long long)ns + 31) / 32;
    if (p->accept_bits) {
        a.abits_pitch = abits_words;
        if (mem == MCQ_MEM_DEVICE) a.abits = p->accept_bits;
        else {
            if (ctx->buf[B_ABITS].ensure((size_t)nc * abits_words * 4 + 4)) return fail(MCQ_ENOMEM, "device allocation failed");
            a.abits = static_cast<uint32_t *>(ctx->buf[B_ABITS].p);
        }
        // (the table kernel stores non-zero words only; a resumed segment keeps the bits of the earlier ones)
        if (!resume) CUDA_TRY(cudaMemsetAsync(a.abits, 0, (size_t)nc * abits_words * 4, s));
        else if (mem == MCQ_MEM_HOST) CUDA_TRY(cudaMemcpyAsync(a.abits, p->accept_bits, (size_t)nc * abits_words * 4, cudaMemcpyHostToDevice, s));
    }

    // ---- statistics accumulators (difference form while the chains run; integrated after the launches) ----
    unsigned long long *d_sum_e = nullptr, *d_sum_e2 = nullptr;
    int *d_cnt = nullptr;
    const size_t stat_elems = (size_t)p->n_groups * ((size_t)ns + 1);
    if (want_stats) {
        if (mem == MCQ_MEM_DEVICE) {
            d_sum_e = reinterpret_cast<unsigned long long *>(p->stat_sum_e); d_sum_e2 = reinterpret_cast<unsigned long long *>(p->stat_sum_e2);
            d_cnt = p->stat_count;
        } else {
            if (ctx->buf[B_SUME].ensure(stat_elems * 8) || ctx->buf[B_SUME2].ensure(stat_elems * 8) ||
                (p->stat_count && ctx->buf[B_SCNT].ensure(stat_elems * 4))) return fail(MCQ_ENOMEM, "device allocation failed");
            d_sum_e = static_cast<unsigned long long *>(ctx->buf[B_SUME].p);
            d_sum_e2 = static_cast<unsigned long long *>(ctx->buf[B_SUME2].p);
            if (p->stat_count) d_cnt = static_cast<int *>(ctx->buf[B_SCNT].p);
        }
        if (!resume) {
            CUDA_TRY(cudaMemsetAsync(d_sum_e, 0, stat_elems * 8, s));
            CUDA_TRY(cudaMemsetAsync(d_sum_e2, 0, stat_elems * 8, s));
            if (d_cnt) CUDA_TRY(cudaMemsetAsync(d_cnt, 0, stat_elems * 4, s));
        } else if (mem == MCQ_MEM_HOST) {
            CUDA_TRY(cudaMemcpyAsync(d_sum_e, p->stat_sum_e, stat_elems * 8, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(d_sum_e2, p->stat_sum_e2, stat_elems * 8, cudaMemcpyHostToDevice, s));
            if (d_cnt) CUDA_TRY(cudaMemcpyAsync(d_cnt, p->stat_count, stat_elems * 4, cudaMemcpyHostToDevice, s));
        }
        a.dsum_e = d_sum_e; a.dsum_e2 = d_sum_e2; a.dcount = d_cnt; a.stat_pitch = (long long)ns + 1;
        a.stat_rows32 = (long long)p->n_groups * ((long long)ns + 1) < (1LL << 31);
    }

    // ---- history plan ----
    int hkind = MCQ_HIST_NONE;
    if (want_hist) hkind = p->hist_dtype;
    const size_t esz = hkind == MCQ_HIST_U16 ? 2 : 4;
    const bool direct = want_hist && mem == MCQ_MEM_DEVICE;  // kernel writes the caller's array itself
    int chunk = ns > 0 ? ns : 1;
    if (hkind != MCQ_HIST_NONE && !direct) {
        if (p->chunk_steps > 0) chunk = p->chunk_steps;
        else {
            const size_t budget = (size_t)1 << 29;  // per buffer (there are two)
            size_t c = budget / ((size_t)nc * esz);
            chunk = (int)std::min<size_t>(std::max<size_t>(c, 64), (size_t)std::max(ns, 1));
        }
    } else if (p->chunk_steps > 0) chunk = p->chunk_steps;
    chunk = std::max(HBLK, chunk / HBLK * HBLK);
    const long long chunk_pitch = ((long long)chunk + 1 + 7) / 8 * 8;   // rows stay 16-byte aligned
    // host histories: two chunk buffers, so that the D2H copy of chunk k (copy stream) runs under the kernel of chunk k+1
    const bool staged = hkind != MCQ_HIST_NONE && !direct;
    const bool two_bufs = staged && (long long)chunk < (long long)(t_stop - t_start);
    if (staged) {
        const size_t bytes = (size_t)nc * chunk_pitch * esz;
        if (ctx->buf[B_HIST0].ensure(bytes) || (two_bufs && ctx->buf[B_HIST1].ensure(bytes))) return fail(MCQ_ENOMEM, "device allocation failed (history chunk)");
    }

    // ---- the launches ----
    // All CTAs of a launch take about the same time, so a launch that needs a fractional number of waves
    // leaves SMs idle in its last one.  The chains are therefore split into two groups that advance on
    // their own streams: while one group's launch drains, the other group's CTAs fill the free slots.
    // Within a group everything is stream-ordered: kernel(chunk k) -> statistics / D2H of chunk k ->
    // kernel(chunk k+1), which is also what lets one history buffer serve every chunk.
    // (Measured: CTA durations vary enough between schedules that one stream loses only ~5 % to partial
    // waves and two streams do not recover it, so one stream is the default; MCQ_STREAMS=2..4 enables the split.)
    int n_sub = 1;
    if (const char *e = getenv("MCQ_STREAMS")) n_sub = std::max(1, std::min(4, std::min(atoi(e), grid)));
    cudaStream_t sub_stream[4];
    for (int b = 0; b < n_sub; ++b) sub_stream[b] = n_sub == 1 ? s : ctx->sub_stream[b];
    EventBag events;
    cudaEvent_t e_start, e_end, e_sub[4], e_kernel[2], e_copy[2];
    CUDA_TRY(events.make(&e_start, cudaEventDefault));
    CUDA_TRY(events.make(&e_end, cudaEventDefault));
    CUDA_TRY(cudaEventRecord(e_start, s));
    for (int b = 0; b < n_sub && n_sub > 1; ++b) {
        CUDA_TRY(events.make(&e_sub[b], cudaEventDisableTiming));
        CUDA_TRY(cudaStreamWaitEvent(sub_stream[b], e_start, 0));
    }
    const bool overlap_copy = two_bufs && n_sub == 1;
    for (int b = 0; b < 2 && overlap_copy; ++b) {
        CUDA_TRY(events.make(&e_kernel[b], cudaEventDisableTiming));
        CUDA_TRY(events.make(&e_copy[b], cudaEventDisableTiming));
    }
    int chunk_no = 0;
    for (int t0 = t_start, first_pass = 1; t0 < t_stop || first_pass; first_pass = 0, ++chunk_no) {
        const int t1 = std::min(t_stop, t0 + chunk);
        const int hb = overlap_copy ? (chunk_no & 1) : 0;   // history buffer of this chunk
        a.t_begin = t0; a.t_end = t1;
        a.hist_kind = hkind;
        if (hkind != MCQ_HIST_NONE) {
            if (direct) { a.hist = p->energy_history; a.hist_pitch = p->hist_pitch; a.h_origin = 0; }
            else { a.hist = ctx->buf[hb ? B_HIST1 : B_HIST0].p; a.hist_pitch = chunk_pitch; a.h_origin = t0 == 0 ? 0 : (long long)t0 + 1; }
        }
        if (overlap_copy && chunk_no >= 2) CUDA_TRY(cudaStreamWaitEvent(s, e_copy[hb], 0));   // the buffer's previous chunk has left
        a.bin_at_begin = 0;
        if (p->n_bins > 0) {
            // the bin that was still open when the previous launch ended (the kernels close a bin
            // when they reach its right edge, which may be the first step of this launch)
            int b = 0;
            if (t0 > 0) while (b + 1 < p->n_bins && p->bin_starts[b + 1] <= t0 - 1) ++b;
            a.bin_at_begin = b;
        }
        // first history column of this chunk and how many columns it produces
        const long long h0 = t0 == 0 ? 0 : (long long)t0 + 1;
        const int n_cols = (int)((long long)t1 + 1 - h0);
        for (int b = 0; b < n_sub; ++b) {
            const int cta_lo = (int)((long long)grid * b / n_sub), cta_hi = (int)((long long)grid * (b + 1) / n_sub);
            const int lo = cta_lo * cpc, hi = std::min(nc, cta_hi * cpc);
            if (hi <= lo) continue;
            cudaStream_t sb = sub_stream[b];
            a.chain_begin = lo; a.n_chains = hi;
            CUDA_TRY(use_spec ? (fast_serves(spec_lpc, a, replay) ? launch_fast(spec_lpc, a, cta_hi - cta_lo, block, smem, sb)
                                                                  : launch_spec(spec_lpc, a, replay, cta_hi - cta_lo, block, smem, sb))
                     : use_wide ? launch_wide(wide_threads, a, cta_hi - cta_lo, smem, sb)
                                : launch_anneal(G, a, replay, cta_hi - cta_lo, block, smem, sb));
            ++launches;
            if (want_hist && !direct && n_cols > 0) {
                cudaStream_t sc = sb;
                if (overlap_copy) {   // copy on the copy stream, ordered after this chunk's kernel
                    sc = ctx->copy_stream;
                    CUDA_TRY(cudaEventRecord(e_kernel[hb], sb));
                    CUDA_TRY(cudaStreamWaitEvent(sc, e_kernel[hb], 0));
                }
                CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(p->energy_history) + ((size_t)lo * p->hist_pitch + (size_t)h0) * esz,
                                           (size_t)p->hist_pitch * esz, static_cast<const char *>(a.hist) + (size_t)lo * chunk_pitch * esz,
                                           (size_t)chunk_pitch * esz, (size_t)n_cols * esz, hi - lo, cudaMemcpyDeviceToHost, sc));
                if (overlap_copy) CUDA_TRY(cudaEventRecord(e_copy[hb], sc));
            }
        }
        t0 = t1;
    }
    for (int b = 0; b < 2 && overlap_copy && b < chunk_no; ++b) CUDA_TRY(cudaStreamWaitEvent(s, e_copy[b], 0));
    a.chain_begin = 0; a.n_chains = nc;
    for (int b = 0; b < n_sub && n_sub > 1; ++b) {
        CUDA_TRY(cudaEventRecord(e_sub[b], sub_stream[b]));
        CUDA_TRY(cudaStreamWaitEvent(s, e_sub[b], 0));
    }
    if (want_stats) {   // differences -> sums over the columns this call produced: history indices [t_start + 1 (or 0), t_stop]
        const long long h0 = t_start == 0 ? 0 : (long long)t_start + 1, h1 = t_stop;
        if (h1 >= h0) {
            stat_integrate_kernel<unsigned long long><<<p->n_groups, 1024, 0, s>>>(d_sum_e, (long long)ns + 1, h0, h1);
            stat_integrate_kernel<unsigned long long><<<p->n_groups, 1024, 0, s>>>(d_sum_e2, (long long)ns + 1, h0, h1);
            if (d_cnt) stat_integrate_kernel<int><<<p->n_groups, 1024, 0, s>>>(d_cnt, (long long)ns + 1, h0, h1);
            CUDA_TRY(cudaGetLastError());
            launches += d_cnt ? 3 : 2;
        }
    }
    CUDA_TRY(cudaEventRecord(e_end, s));

    // ---- outputs ----
    if (int rc = copy_out(p->initial_energy, a.init_e, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->final_energy, a.cur_e, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->best_energy, a.best_e, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->steps_to_best, a.best_step, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->n_accepted, a.n_acc, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->steps_done, a.steps_done, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->n_near_threshold, a.near_cnt, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->n_fp32_flips, a.flip_cnt, (size_t)nc * 4, mem, s)) return rc;
    if (int rc = copy_out(p->record_out, rec, (size_t)nc * MCQ_RECORD_INTS * 4, mem, s)) return rc;
    if (int rc = copy_out(p->final_state, a.state, (size_t)nc * sbytes, mem, s)) return rc;
    if (int rc = copy_out(p->best_state, a.best_state, (size_t)nc * sbytes, mem, s)) return rc;
    if (mem == MCQ_MEM_HOST) {
        if (p->n_bins > 0) CUDA_TRY(cudaMemcpyAsync(p->accept_hist, a.acc_hist, (size_t)nc * p->n_bins * 4, cudaMemcpyDeviceToHost, s));
        if (p->accept_bits) CUDA_TRY(cudaMemcpyAsync(p->accept_bits, a.abits, (size_t)nc * abits_words * 4, cudaMemcpyDeviceToHost, s));
        if (want_stats) {
            CUDA_TRY(cudaMemcpyAsync(p->stat_sum_e, d_sum_e, stat_elems * 8, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(p->stat_sum_e2, d_sum_e2, stat_elems * 8, cudaMemcpyDeviceToHost, s));
            if (d_cnt) CUDA_TRY(cudaMemcpyAsync(p->stat_count, d_cnt, stat_elems * 4, cudaMemcpyDeviceToHost, s));
        }
    }
    uint32_t h_err = 0;
    if (replay) CUDA_TRY(cudaMemcpyAsync(&h_err, a.replay_err, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms_total = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms_total, e_start, e_end));   // device span of all launches of this call
    if (p->kernel_ms) *p->kernel_ms = ms_total;
    if (p->gpu_launches) *p->gpu_launches = launches;
    if (replay && h_err) return fail(MCQ_EREPLAY, "replayed stream contained an illegal proposal (occupied cell, same height or out of range)");
    return 0;
}

}  // extern "C"
