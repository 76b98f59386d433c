// Fused simulated-annealing kernel for the 3D N^2-queens chain (sm_100a).
//
// Replaces the inner loop of metropolis_mcmc (experiments.py:218-258) and
// metropolis_mcmc_board (experiments.py:308-355) for a whole batch of independent chains.
//
// Mapping.  A *lane group* of G lanes (G = 4, 8, 16 or 32) owns one chain; a warp carries 32/G
// chains and a CTA a few warps.  Everything a chain touches per step lives in its own slab of
// shared memory for the entire launch:
//
//   counters  one uint8 per attack line (13 families: 3 axis, 6 planar-diagonal, 4
//             space-diagonal; board mode drops the (i,j) column family).  Two distinct cells
//             share at most one line, so E = sum_lines C(count,2) and the conflict counts of
//             mcmc.py:185-226 / mcmc_board.py:147-193 become 13 (12) byte reads per cell.
//   state     board: heights[N*N] uint8;  full_3d: packed (i,j,k) per queen + N^3-bit occupancy
//   packets   the next PB steps' random words and beta values (produced PB steps at a time, one
//             Philox4x32-10 call per lane, so the generator is amortised over the group)
//   staging   32 steps of energy history, flushed to HBM as one contiguous run per chain
//
// Every line index is affine in (i,j,k): lane g of a group evaluates families g, g+G, ... with
// its own coefficient registers, for the old and the new cell, and the group reduces the signed
// sum with shuffles (REDUX when G == 32).  Proposal draw, delta-E, Metropolis test, counter
// update, best tracking and history append are one loop iteration; nothing leaves the SM
// except the history flush.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "accept.cuh"
#include "philox.cuh"

namespace mcq {

constexpr int NFAM = 13;      // attack-line families
constexpr int HBLK = 32;      // steps per history staging block / accept-bitmap word
constexpr int PKT_BYTES = 32; // one step's packet: 4 random words + beta (production) or move,u,beta (replay)

// byte offsets inside one chain's shared-memory slab
struct Layout {
    int n_cnt;      // counter bytes, multiple of 4
    int off_state;  // board: heights; full_3d: packed queen positions
    int off_occ;    // full_3d: occupancy bitset (N^3 bits)
    int off_pkt;    // PB packets
    int off_hst;    // HBLK int32 energies
    int stride;     // slab size, multiple of 16
    int pos32;      // full_3d: positions are uint32 (8-bit fields) instead of uint16 (5-bit fields)
    int pb;         // packets produced per batch = min(G, 8)
    int off_jrn;    // G == 1: journal of state elements changed since the last best-state snapshot
};

constexpr int JRN = 62;   // journal capacity (uint16 entries); slot JRN holds the count

// byte offsets inside one chain's slab for the conflict-table kernel (spec.cuh)
struct SLayout {
    int tbl;        // table bytes (N^3 + 1 scratch byte, rounded up to 4)
    int off_state;  // board: heights; full_3d: uint32 per queen = cell id | wide id << 16
    int off_occ;    // full_3d: occupancy bitset
    int off_rec;    // spare words
    int off_ring;   // ring of Philox words: 64 steps x 16 B
    int stride;     // slab size, multiple of 16
    int nbr_len;    // neighbour-row length: families*(N-1) rounded up to 32
    int rounds;     // nbr_len / 32
    int off_wide;   // CTA-shared geometry (full_3d): [lut bits | wide ids]; byte offset of the wide ids
    int cta_bytes;  // bytes of CTA-shared geometry in front of the slabs (0 in board mode)
    int wide_bias;  // (N-1)*(W^2+W+1), W = 2N-1
};

struct KArgs {
    int full;  // 0 board, 1 full_3d
    int N, Q;
    int n_chains;         // one past the last chain of this launch (absolute index)
    int chain_begin;      // first chain of this launch: CTA 0 starts here
    int n_steps;          // total steps of the schedule (row length of the beta tables)
    int t_begin, t_end;   // this launch covers steps [t_begin, t_end)
    int t_last;           // the call (segment) ends at this step: global-memory slabs write their state back then
    int patience;         // < 0: none
    Layout lay;
    SLayout sl;
    int4 coef[NFAM];      // idx = x*i + y*j + z*k + w  (w includes the family base) ...
    int4 csel[NFAM];      // ... + (x*i + y*j + w < 0 ? z : 0): the fold of the space-diagonal families (all zero for the others)
    // per-chain inputs
    const unsigned long long *seeds;
    const int *group;     // may be null
    const float *beta_c;  // [n_groups][n_steps]  float32(-beta*log2(e)), built on the device (beta_table_kernel)
    const SchedDev *sched; // [n_groups] schedule parameters (float64 rule of the accept test), or
    float band_abs;        // absolute part of the float32 error band (4; +inf = every uphill decision in float64)
    uint32_t *flip_cnt;    // [n_chains] production: band decisions that float32 alone would have got wrong
    // cross-replica statistics in difference form: column h of group g receives what the chains of g add to
    // sum E, sum E^2 and the number of live chains when they reach history index h (mcq_run integrates them)
    unsigned long long *dsum_e, *dsum_e2;
    int *dcount;           // may be null
    long long stat_pitch;  // elements per group row (n_steps + 1)
    int stat_rows32;       // n_groups * stat_pitch < 2^31: kernels may index the rows with 32 bits
    // replay
    const double *beta64;  // replay: exact betas; production: float64 table of a tabulated closure (else `sched`)
    const uint32_t *rmoves;
    const double *runif;
    uint32_t *near_cnt;
    uint32_t *replay_err;  // single counter
    // persistent per-chain record
    uint8_t *state;        // [n_chains][state_bytes] external format
    uint8_t *best_state;
    int state_bytes;
    int *init_e, *cur_e, *best_e, *best_step, *n_acc, *steps_done, *stale, *bin_mark;
    // history
    void *hist;            // chunk or full buffer
    int hist_kind;         // MCQ_HIST_*
    long long hist_pitch;
    long long h_origin;    // history index stored at column 0 of `hist`
    uint32_t *abits;
    long long abits_pitch;
    // acceptance bins
    const int *bin_starts; // device copy, [n_bins+1]
    int n_bins;
    int bin_at_begin;      // bin containing t_begin
    uint32_t *acc_hist;    // [n_chains][n_bins]
    int gslab_dummy;       // index of the dummy slab
    unsigned char *gslab;  // line-counter kernel with G == 1: slabs live in global memory ([n_chains + 1][lay.stride],
                           // the last one a zeroed dummy for the padding threads of the last CTA); null = shared memory
    int w_best, w_ring, w_xch;   // CTA-per-chain kernel (wide.cuh): byte offsets of the best-state copy, the word ring, the exchange words
    const uint16_t *nbr;   // conflict-table kernel: neighbour lists [N^3][sl.nbr_len]
    const uint32_t *geo;   // conflict-table kernel, full_3d: shared-line bits then wide ids (sl.cta_bytes)
};

// Counter index of the attack line of family (c, s) through cell (i, j, k).  Axis and planar-diagonal families are
// affine in (i, j, k).  A space-diagonal family is a hexagon of 3N^2-3N+1 lines inside the (2N-1)^2 square of
// (a, b) = (i -+ j, i -+ k) pairs; rows a and a - N of the hexagon together are 3N-2 long, so it folds into N rows of
// 3N-2 with idx = a(3N-3) + b + (N-1) + (a < 0 ? 3N^2-N-1 : 0): affine plus one select on the sign of a
// (make_coefs, mcq_api.cu).  That is 25 % fewer counter bytes than the square (N = 64: 105.6 KB instead of 121.5 KB
// per chain -- two chains per SM instead of one).
__device__ __forceinline__ int line_index(const int4 c, const int4 s, int i, int j, int k) {
    return c.x * i + c.y * j + c.z * k + c.w + ((s.x * i + s.y * j + s.w) < 0 ? s.z : 0);
}
// family known at compile time: only the four space-diagonal families (9..12) carry the select
template <int F>
__device__ __forceinline__ int line_index_f(const KArgs &a, int i, int j, int k) {
    const int4 c = a.coef[F];
    int idx = c.x * i + c.y * j + c.z * k + c.w;
    if constexpr (F >= 9) {
        const int4 s = a.csel[F];
        idx += (s.x * i + s.y * j + s.w) < 0 ? s.z : 0;
    }
    return idx;
}

// statistics in difference form (see KArgs::dsum_e): one lane per chain calls this when the chain's energy
// changes from e_old to e_new at history index h (dc = change of the live-chain count)
__device__ __forceinline__ void stat_delta(const KArgs &a, int grp, long long h, int e_old, int e_new, int dc) {
    const size_t idx = (size_t)grp * (size_t)a.stat_pitch + (size_t)h;
    const long long de = (long long)e_new - (long long)e_old;
    if (de) {
        atomicAdd(a.dsum_e + idx, (unsigned long long)de);
        atomicAdd(a.dsum_e2 + idx, (unsigned long long)(de * ((long long)e_new + (long long)e_old)));
    }
    if (dc && a.dcount) atomicAdd(a.dcount + idx, dc);
}

template <int G>
__device__ __forceinline__ int group_sum(int v) {
    if constexpr (G == 32) {
        return __reduce_add_sync(0xffffffffu, v);
    } else {
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        return v;
    }
}

__device__ __forceinline__ uint32_t pack_pos(int pos32, int i, int j, int k) {
    return pos32 ? (uint32_t)(i | (j << 8) | (k << 16)) : (uint32_t)(i | (j << 5) | (k << 10));
}
__device__ __forceinline__ void unpack_pos(int pos32, uint32_t p, int &i, int &j, int &k) {
    if (pos32) { i = p & 255; j = (p >> 8) & 255; k = (p >> 16) & 255; }
    else { i = p & 31; j = (p >> 5) & 31; k = (p >> 10) & 31; }
}
__device__ __forceinline__ uint32_t load_pos(const unsigned char *st, int pos32, int q) {
    return pos32 ? reinterpret_cast<const uint32_t *>(st)[q] : (uint32_t) reinterpret_cast<const uint16_t *>(st)[q];
}
__device__ __forceinline__ void store_pos(unsigned char *st, int pos32, int q, uint32_t p) {
    if (pos32) reinterpret_cast<uint32_t *>(st)[q] = p;
    else reinterpret_cast<uint16_t *>(st)[q] = (uint16_t)p;
}

// Build one chain's slab from its external state: counters, packed state, occupancy.  Returns the
// full energy (same on every lane of the group).  All lanes of the warp must call this.
template <int G>
__device__ int build_chain(const KArgs &a, unsigned char *S, const uint8_t *ext, bool live, int g) {
    uint32_t *W = reinterpret_cast<uint32_t *>(S);
    for (int w = g; w < a.lay.off_pkt / 4; w += G) W[w] = 0u;
    __syncwarp();
    if (live) {
        unsigned char *st = S + a.lay.off_state;
        uint32_t *occ = reinterpret_cast<uint32_t *>(S + a.lay.off_occ);
        for (int qi = g; qi < a.Q; qi += G) {
            int i, j, k;
            if (a.full) {
                i = ext[3 * qi]; j = ext[3 * qi + 1]; k = ext[3 * qi + 2];
                store_pos(st, a.lay.pos32, qi, pack_pos(a.lay.pos32, i, j, k));
                const int cid = (i * a.N + j) * a.N + k;
                atomicOr(&occ[cid >> 5], 1u << (cid & 31));
            } else {
                i = qi / a.N; j = qi - i * a.N; k = ext[qi];
                st[qi] = (unsigned char)k;
            }
            for (int f = a.full ? 0 : 1; f < NFAM; ++f) {
                const int idx = line_index(a.coef[f], a.csel[f], i, j, k);
                atomicAdd(&W[idx >> 2], 1u << ((idx & 3) * 8));
            }
        }
    }
    __syncwarp();
    int e = 0;
    for (int w = g; w < a.lay.n_cnt / 4; w += G) {
        const uint32_t v = W[w];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = (v >> (8 * b)) & 255;
            e += (c * (c - 1)) >> 1;
        }
    }
    return group_sum<G>(e);
}

// Per-lane view of the attack-line families this lane evaluates.
template <int G>
struct LaneLines {
    static constexpr int R = (NFAM + G - 1) / G;
    int4 c[R], s[R];
    const KArgs *ka;   // G == 1: a thread evaluates every family, straight from the kernel arguments (no register copies)
    __device__ __forceinline__ void init(const KArgs &a, int g) {
        ka = &a;
        if constexpr (G == 1) return;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int f = g + r * G;
            const bool ok = f < NFAM && (a.full || f != 0);
            // an unused slot maps old and new cell to the same counter: contributes 0, never updated
            c[r] = ok ? a.coef[f] : make_int4(0, 0, 0, 0);
            s[r] = ok ? a.csel[f] : make_int4(0, 0, 0, 0);
        }
    }
};

// One proposal's line work for this lane: the partial delta-E plus what the accept path needs.
template <int G>
struct LineEval {
    static constexpr int R = LaneLines<G>::R;
    int io[R], in[R];
    int co[R], cn[R];
    __device__ __forceinline__ int eval(const LaneLines<G> &L, const uint8_t *cnt, int i0, int j0, int k0, int i1,
                                        int j1, int k1) {
        int d = 0;
        if constexpr (G == 1) {
            // families are compile-time here: board mode maps family 0 (the (i,j) column) to one counter for both cells
            const KArgs &a = *L.ka;
            auto both = [&](auto fc) {
                constexpr int f = decltype(fc)::value;
                if (f == 0 && !a.full) { io[f] = 0; in[f] = 0; }
                else { io[f] = line_index_f<f>(a, i0, j0, k0); in[f] = line_index_f<f>(a, i1, j1, k1); }
            };
            both(std::integral_constant<int, 0>{}); both(std::integral_constant<int, 1>{}); both(std::integral_constant<int, 2>{});
            both(std::integral_constant<int, 3>{}); both(std::integral_constant<int, 4>{}); both(std::integral_constant<int, 5>{});
            both(std::integral_constant<int, 6>{}); both(std::integral_constant<int, 7>{}); both(std::integral_constant<int, 8>{});
            both(std::integral_constant<int, 9>{}); both(std::integral_constant<int, 10>{}); both(std::integral_constant<int, 11>{});
            both(std::integral_constant<int, 12>{});
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                io[r] = line_index(L.c[r], L.s[r], i0, j0, k0);
                in[r] = line_index(L.c[r], L.s[r], i1, j1, k1);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            co[r] = cnt[io[r]];
            cn[r] = cnt[in[r]];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // old_conf = sum(co-1); new_conf = sum(cn) - [cells share a line]; a shared line has io==in
            d += (io[r] != in[r]) ? (cn[r] - co[r] + 1) : 0;
        }
        return d;
    }
    __device__ __forceinline__ void apply(uint8_t *cnt) const {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (io[r] != in[r]) {
                cnt[io[r]] = (uint8_t)(co[r] - 1);
                cnt[in[r]] = (uint8_t)(cn[r] + 1);
            }
        }
    }
};

// uniform integer in [0, n) from the high part of w*n; returns the low part for the next digit
__device__ __forceinline__ int draw_digit(uint32_t &w, int n) {
    const uint32_t hi = __umulhi(w, (uint32_t)n);
    w = w * (uint32_t)n;
    return (int)hi;
}

template <int G, bool FULL, bool REPLAY>
__global__ void __launch_bounds__(G == 1 ? 128 : 256, G == 1 ? 3 : 1) anneal_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int cl = tid / G;
    const int cpc = blockDim.x / G;
    const int chain = a.chain_begin + blockIdx.x * cpc + cl;
    const bool live = chain < a.n_chains;
    const int N = a.N;
    const int PB = a.lay.pb;

    // G == 1 with a global slab: one THREAD per chain, counters in HBM/L2 (boards too large for shared
    // memory, e.g. N = 64: 121 KB per chain); the slab was built by gslab_build_kernel and persists
    // between launches.
    unsigned char *S = a.gslab ? a.gslab + (size_t)(live ? chain : a.gslab_dummy) * a.lay.stride
                               : smem + (size_t)cl * a.lay.stride;
    uint8_t *cnt = S;
    unsigned char *st = S + a.lay.off_state;
    uint32_t *occ = reinterpret_cast<uint32_t *>(S + a.lay.off_occ);
    unsigned char *pkt = S + a.lay.off_pkt;
    int *hst = reinterpret_cast<int *>(S + a.lay.off_hst);
    const int pos32 = a.lay.pos32;

    const uint8_t *ext = a.state + (size_t)(live ? chain : 0) * a.state_bytes;
    int E = a.gslab ? (live ? a.cur_e[chain] : 0) : build_chain<G>(a, S, ext, live, g);

    LaneLines<G> L;
    L.init(a, g);

    // ---- persistent record ----
    int best = E, best_step = 0, n_acc = 0, stale = 0, bin_mark = 0;
    int done = a.t_end;  // first step NOT executed (early stop lowers it)
    bool active = live;
    if (a.t_begin == 0) {
        if (live && g == 0) {
            if (a.init_e) a.init_e[chain] = E;
            if (a.hist_kind == 1) reinterpret_cast<uint16_t *>(a.hist)[(size_t)chain * a.hist_pitch] = (uint16_t)E;
            else if (a.hist_kind == 2) reinterpret_cast<int *>(a.hist)[(size_t)chain * a.hist_pitch] = E;
            if (a.dsum_e) stat_delta(a, a.group ? a.group[chain] : 0, 0, 0, E, 1);
        }
    } else if (live) {
        best = a.best_e[chain];
        best_step = a.best_step[chain];
        n_acc = a.n_acc[chain];
        stale = a.stale[chain];
        bin_mark = a.bin_mark[chain];
        const int sd = a.steps_done[chain];
        if (sd < a.t_begin) { active = false; done = sd; }
    }

    uint32_t k0 = 0, k1 = 0;
    int grp = 0;
    if (live) {
        const unsigned long long sd = a.seeds ? a.seeds[chain] : 0ull;
        k0 = (uint32_t)sd; k1 = (uint32_t)(sd >> 32);
        grp = a.group ? a.group[chain] : 0;
    }
    const float *beta_row = REPLAY ? nullptr : a.beta_c + (size_t)grp * a.n_steps;
    const double *beta64_row = REPLAY ? a.beta64 + (size_t)grp * a.n_steps : nullptr;
    const uint32_t *mv_row = REPLAY ? a.rmoves + (size_t)(live ? chain : 0) * a.n_steps : nullptr;
    const double *un_row = REPLAY ? a.runif + (size_t)(live ? chain : 0) * a.n_steps : nullptr;

    float c_pref = 0.f;  // beta for step (next batch start + g), fetched one batch ahead
    if (!REPLAY && live && g < PB && a.t_begin + g < a.t_end) c_pref = __ldg(beta_row + a.t_begin + g);

    int bin = a.bin_at_begin;
    int next_edge = a.n_bins > 0 ? a.bin_starts[bin + 1] : 0x7fffffff;
    uint32_t accbits = 0u;
    uint32_t near = 0u, flips = 0u;
    int blk_t0 = a.t_begin;  // first step staged in the current history block

    for (int t = a.t_begin; t < a.t_end; ++t) {
        // ---------------- packet production (once per PB steps) ----------------
        // (a single lane per chain keeps its step's words in registers: no staging)
        uint4 w1 = make_uint4(0u, 0u, 0u, 0u);
        uint32_t w1_4 = 0u;
        if constexpr (G == 1) {
            const int ts = min(t, a.t_end - 1);
            if constexpr (REPLAY) {
                const double u = un_row[ts], b = beta64_row[ts];
                w1 = make_uint4(mv_row[ts], (uint32_t)__double2loint(u), (uint32_t)__double2hiint(u), (uint32_t)__double2loint(b));
                w1_4 = (uint32_t)__double2hiint(b);
            } else {
                const Philox4 r = step_words<FULL>((uint32_t)ts, k0, k1, (uint32_t)(N * N));
                w1 = make_uint4(r.x, r.y, r.z, r.w);
                w1_4 = __float_as_uint(c_pref);                       // fetched one step ahead
                if (live && ts + 1 < a.t_end) c_pref = __ldg(beta_row + ts + 1);
            }
        } else if (((t - a.t_begin) & (PB - 1)) == 0) {
            __syncwarp();
            if (g < PB) {
                const int ts = t + g;
                uint32_t *p = reinterpret_cast<uint32_t *>(pkt + g * PKT_BYTES);
                if (ts < a.t_end && live) {
                    if constexpr (REPLAY) {
                        const double u = un_row[ts], b = beta64_row[ts];
                        p[0] = mv_row[ts];
                        p[1] = (uint32_t)__double2loint(u); p[2] = (uint32_t)__double2hiint(u);
                        p[3] = (uint32_t)__double2loint(b); p[4] = (uint32_t)__double2hiint(b);
                    } else {
                        const Philox4 r = step_words<FULL>((uint32_t)ts, k0, k1, (uint32_t)(N * N));
                        *reinterpret_cast<uint4 *>(p) = make_uint4(r.x, r.y, r.z, r.w);
                        p[4] = __float_as_uint(c_pref);
                        const int tn = ts + PB;
                        if (tn < a.t_end) c_pref = __ldg(beta_row + tn);
                    }
                }
            }
            __syncwarp();
        }
        // ---------------- acceptance-bin bookkeeping (uniform in t) ----------------
        while (t == next_edge) {
            if (active && g == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
            bin_mark = n_acc;
            ++bin;
            next_edge = a.bin_starts[bin + 1];
        }

        const unsigned char *pk = pkt + ((t - a.t_begin) & (PB - 1)) * PKT_BYTES;
        const uint4 w = G == 1 ? w1 : *reinterpret_cast<const uint4 *>(pk);
        const uint32_t w4 = G == 1 ? w1_4 : reinterpret_cast<const uint32_t *>(pk)[4];

        // ---------------- proposal ----------------
        int i0, j0, k0c, i1, j1, k1c, qsel = 0, cid1 = 0;
        bool bad = false;
        if constexpr (FULL) {
            if constexpr (REPLAY) {
                qsel = w.x & 0xfff; i1 = (w.x >> 12) & 63; j1 = (w.x >> 18) & 63; k1c = (w.x >> 24) & 63;
                cid1 = (i1 * N + j1) * N + k1c;
                bad = qsel >= a.Q || i1 >= N || j1 >= N || k1c >= N;
                if (bad) { qsel = 0; i1 = j1 = k1c = 0; cid1 = 0; }
                bad = bad || ((occ[cid1 >> 5] >> (cid1 & 31)) & 1u);
            } else {
                qsel = (int)__umulhi(w.x, (uint32_t)a.Q);
                uint32_t word = w.y;
                int tries = 0;
                while (true) {
                    i1 = draw_digit(word, N); j1 = draw_digit(word, N); k1c = draw_digit(word, N);
                    cid1 = (i1 * N + j1) * N + k1c;
                    if (!live || !((occ[cid1 >> 5] >> (cid1 & 31)) & 1u)) break;
                    // occupied (the queen's own cell counts, experiments.py:230): redraw
                    if (tries == 0) word = w.w;
                    else if (tries == 1) word = w.x * (uint32_t)a.Q;   // what the queen draw left of word x
                    else {
                        const int e = tries - 2;
                        const Philox4 r = chain_words((uint32_t)t, k0, k1, 1u + (uint32_t)(e >> 2));
                        const int s = e & 3;
                        word = s == 0 ? r.x : s == 1 ? r.y : s == 2 ? r.z : r.w;
                    }
                    ++tries;
                }
            }
            unpack_pos(pos32, load_pos(st, pos32, qsel), i0, j0, k0c);
        } else {
            if constexpr (REPLAY) {
                i0 = w.x & 255; j0 = (w.x >> 8) & 255; k1c = (w.x >> 16) & 255;
                bad = i0 >= N || j0 >= N || k1c >= N;
                if (bad) { i0 = j0 = k1c = 0; }
                k0c = st[i0 * N + j0];
                bad = bad || (k1c == k0c);
            } else {
                uint32_t word = w.x;
                i0 = draw_digit(word, N); j0 = draw_digit(word, N);
                k0c = st[i0 * N + j0];
                // uniform over the N-1 other heights (== the redraw loop of experiments.py:317-319)
                k1c = k0c + 1 + (int)__umulhi(w.y, (uint32_t)(N - 1));
                k1c -= (k1c >= N) ? N : 0;
            }
            i1 = i0; j1 = j0;
        }

        // ---------------- delta-E from the line counters ----------------
        LineEval<G> ev;
        const int dE = group_sum<G>(ev.eval(L, cnt, i0, j0, k0c, i1, j1, k1c));

        // ---------------- Metropolis test (experiments.py:238-239 / :326-327) ----------------
        bool accept;
        if constexpr (REPLAY) {
            const double u = __hiloint2double((int)w.z, (int)w.y);
            const double b = __hiloint2double((int)w4, (int)w.w);
            const double p = exp(-b * (double)dE);
            accept = u < fmin(1.0, p);
            if (active && !bad && fabs(u - p) < 1e-6) ++near;
            if (bad) { accept = false; if (active && g == 0) atomicAdd(a.replay_err, 1u); }
        } else {
            bool near_band;
            metropolis_fast(dE, __uint_as_float(w4), w.z, a.band_abs, accept, near_band);
            if (near_band && active) {   // inside the float32 error band: the float64 rule decides (accept.cuh)
                const bool exact = metropolis_exact_call(a.sched, a.beta64, a.n_steps, grp, k0, k1, t, dE, w.z);
                ++near;
                flips += (uint32_t)(exact != accept);
                accept = exact;
            }
        }
        accept = accept && active;

        // ---------------- apply ----------------
        if (accept) {
            ev.apply(cnt);
            if (g == 0) {
                if constexpr (FULL) {
                    const int cid0 = (i0 * N + j0) * N + k0c;
                    occ[cid0 >> 5] &= ~(1u << (cid0 & 31));
                    occ[cid1 >> 5] |= 1u << (cid1 & 31);
                    store_pos(st, pos32, qsel, pack_pos(pos32, i1, j1, k1c));
                } else {
                    st[i0 * N + j0] = (unsigned char)k1c;
                }
            }
            if (a.dsum_e && g == 0 && dE != 0) stat_delta(a, grp, (long long)t + 1, E, E + dE, 0);
            E += dE;
            ++n_acc;
            accbits |= 1u << (t & 31);
            if constexpr (G == 1) {
                // one thread per chain: copying the whole state at every new best would dominate, so the
                // elements touched since the last snapshot are journalled and only those are copied
                uint16_t *jrn = reinterpret_cast<uint16_t *>(S + a.lay.off_jrn);
                const int jn = jrn[JRN];
                if (jn < JRN) jrn[jn] = (uint16_t)(FULL ? qsel : i0 * N + j0);
                if (jn <= JRN) jrn[JRN] = (uint16_t)(jn + 1);   // JRN + 1 = overflow: next snapshot is a full copy
            }
        }
        __syncwarp();
        bool improved = accept && (E < best);
        bool stop_now = false;
        if (!FULL && a.patience >= 0 && active) {
            stale = improved ? 0 : stale + 1;
            stop_now = stale >= a.patience;   // break happens before the history append (:349-355)
        }
        if (improved) {
            best = E;
            if (!stop_now) best_step = t + 1;
            // snapshot: the state at the first visit of the minimum (strict <, :252 / :340)
            uint8_t *bs = a.best_state + (size_t)chain * a.state_bytes;
            bool full_copy = true;
            if constexpr (G == 1) {
                uint16_t *jrn = reinterpret_cast<uint16_t *>(S + a.lay.off_jrn);
                const int jn = jrn[JRN];
                if (jn <= JRN) {
                    full_copy = false;
                    for (int e = 0; e < jn; ++e) {
                        const int el = jrn[e];
                        if constexpr (FULL) {
                            int i, j, k;
                            unpack_pos(pos32, load_pos(st, pos32, el), i, j, k);
                            bs[3 * el] = (uint8_t)i; bs[3 * el + 1] = (uint8_t)j; bs[3 * el + 2] = (uint8_t)k;
                        } else {
                            bs[el] = st[el];
                        }
                    }
                }
                jrn[JRN] = 0;
            }
            if (full_copy) {
                if constexpr (FULL) {
                    for (int qi = g; qi < a.Q; qi += G) {
                        int i, j, k;
                        unpack_pos(pos32, load_pos(st, pos32, qi), i, j, k);
                        bs[3 * qi] = (uint8_t)i; bs[3 * qi + 1] = (uint8_t)j; bs[3 * qi + 2] = (uint8_t)k;
                    }
                } else {
                    for (int c = g; c < a.Q; c += G) bs[c] = st[c];
                }
            }
        }
        if (stop_now) {
            active = false;
            done = t;
            if (a.dsum_e && g == 0) stat_delta(a, grp, (long long)t + 1, E, 0, -1);   // no energy is appended from here on
            if (g == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
        }
        if (active && g == 0 && (a.hist_kind || G > 1)) hst[t & (HBLK - 1)] = E;

        // ---------------- history flush ----------------
        if ((t & (HBLK - 1)) == HBLK - 1 || t == a.t_end - 1) {
            __syncwarp();
            if (live) {
                const int lim = done < t + 1 ? done : t + 1;   // steps < lim have an appended energy
                if (a.hist_kind == 1) {
                    uint16_t *h = reinterpret_cast<uint16_t *>(a.hist) + (size_t)chain * a.hist_pitch;
                    for (int s = blk_t0 + g; s < lim; s += G) h[(long long)s + 1 - a.h_origin] = (uint16_t)hst[s & (HBLK - 1)];
                } else if (a.hist_kind == 2) {
                    int *h = reinterpret_cast<int *>(a.hist) + (size_t)chain * a.hist_pitch;
                    for (int s = blk_t0 + g; s < lim; s += G) h[(long long)s + 1 - a.h_origin] = hst[s & (HBLK - 1)];
                }
                if (g == 0 && a.abits) a.abits[(size_t)chain * a.abits_pitch + (blk_t0 >> 5)] = accbits;
            }
            accbits = 0u;
            blk_t0 = t + 1;
            __syncwarp();
        }
    }

    // ---------------- write the record back ----------------
    if (live) {
        if (a.t_end == a.n_steps && a.n_bins > 0 && g == 0 && a.acc_hist && done == a.t_end) {
            a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
        }
        uint8_t *out = a.state + (size_t)chain * a.state_bytes;
        if (!a.gslab || a.t_end == a.t_last) {   // a global slab keeps the state between the launches of a call
            if constexpr (FULL) {
                for (int qi = g; qi < a.Q; qi += G) {
                    int i, j, k;
                    unpack_pos(pos32, load_pos(st, pos32, qi), i, j, k);
                    out[3 * qi] = (uint8_t)i; out[3 * qi + 1] = (uint8_t)j; out[3 * qi + 2] = (uint8_t)k;
                }
            } else {
                for (int c = g; c < a.Q; c += G) out[c] = st[c];
            }
        }
        if (g == 0) {
            a.cur_e[chain] = E;
            a.best_e[chain] = best;
            a.best_step[chain] = best_step;
            a.n_acc[chain] = n_acc;
            a.stale[chain] = stale;
            a.bin_mark[chain] = bin_mark;
            a.steps_done[chain] = done;
            if (a.near_cnt) a.near_cnt[chain] += near;
            if (!REPLAY && a.flip_cnt) a.flip_cnt[chain] += flips;
        }
    }
}

// Builds the global-memory slabs (counters, packed state, occupancy) and the initial energies: one warp per chain.
static __global__ void __launch_bounds__(32) gslab_build_kernel(const __grid_constant__ KArgs a) {
    const int chain = a.chain_begin + blockIdx.x;
    if (chain >= a.n_chains) return;
    const int e = build_chain<32>(a, a.gslab + (size_t)chain * a.lay.stride, a.state + (size_t)chain * a.state_bytes, true, threadIdx.x);
    if (threadIdx.x == 0) {
        a.cur_e[chain] = e;
        // fresh run: best state == state, empty journal; resumed segment: the best state is an older one, so the
        // journal starts overflowed and the next new best copies the whole state
        reinterpret_cast<uint16_t *>(a.gslab + (size_t)chain * a.lay.stride + a.lay.off_jrn)[JRN] = a.t_begin > 0 ? JRN + 1 : 0;
    }
}

}  // namespace mcq
