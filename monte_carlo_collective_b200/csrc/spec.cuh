// Speculative warp-per-chain annealing kernel with a per-cell conflict table (sm_100a).
//
// Same chain, same random stream and therefore the same trajectory, bit for bit, as
// anneal_kernel (anneal.cuh) -- but organised so that every lane does distinct useful work:
//
//   * A warp owns ONE chain.  Lane l evaluates the proposal of step t+l against the chain's
//     current state, all 32 in parallel.  Rejected proposals do not change the state, so the
//     sequential chain of experiments.py:218-258 / :308-355 is reproduced exactly by committing
//     the FIRST accepted lane (ballot + ffs), advancing t past it, and re-evaluating the
//     discarded lanes in the next round with the same counter-based Philox words (the random
//     numbers of step s depend on (seed, s) only).  With ~3 % acceptance a round retires ~21
//     proposals.
//   * delta-E is two table loads: the slab holds T[cell] = number of queens on the 13 (12 in board mode)
//     attack lines through `cell`, a queen standing on the cell itself counted once -- for an empty
//     cell exactly what conflicts_for_position counts, for a queen's cell conflicts_for_queen + 1
//     (mcmc.py:185-226, mcmc_board.py:147-193):
//         conflicts(old) = T[old] - 1,   conflicts(new) = T[new] - [old and new share a line]
//     "old and new share a line" depends on the coordinate differences only and is one bit of a
//     (2N-1)^3-bit table.  T is maintained on accept only: -1 on the old cell and every cell of the
//     lines through it, +1 on the new one and its lines; the cell lists are rows of a neighbour
//     table shared by all chains (geometry only, L2-resident), 32 cells per warp load.  In full_3d
//     bit 15 of an entry says "a queen stands here" (the occupancy test of a candidate cell and its
//     conflict count are one load); it is flipped by the same read-modify-write that counts the
//     queen herself, because slot 0 of every row is the cell itself.
//   * cells are handled as linear ids c = (i*N+j)*N+k throughout; coordinates are only
//     reconstructed when a state leaves the kernel.
//   * per-chain shared memory is ~N^3 + 4*Q bytes (2.6 KB at N=12), so the register file, not
//     shared memory, bounds residency.
//
// Table entries are uint8: T <= 13*N (12*N in board mode), so the kernel serves N <= 19 in
// full_3d and N <= 21 in board mode; larger boards use anneal_kernel's line counters.
#pragma once
#include <type_traits>

#include "anneal.cuh"
#include "geometry.cuh"

namespace mcq {

// CN > 0: board size (and Q = N^2) known at compile time; CN == 0: taken from the kernel arguments
template <bool FULL, int CN>
struct SpecGeom {
    static __device__ __forceinline__ constexpr SLayout layout(const KArgs &) { return spec_layout(FULL, CN, CN * CN); }
    static __device__ __forceinline__ constexpr int n(const KArgs &) { return CN; }
    static __device__ __forceinline__ constexpr int q(const KArgs &) { return CN * CN; }
};
template <bool FULL>
struct SpecGeom<FULL, 0> {
    static __device__ __forceinline__ SLayout layout(const KArgs &a) { return a.sl; }
    static __device__ __forceinline__ int n(const KArgs &a) { return a.N; }
    static __device__ __forceinline__ int q(const KArgs &a) { return a.Q; }
};

// ---- shared memory is addressed by 32-bit byte offsets into the dynamic array -----------------
// (no generic pointers: one base register per slab instead of 64-bit pointer pairs)
// Explicit shared-state-space accesses on 32-bit addresses: one base register per slab, and no
// generic-pointer arithmetic in the loop.
extern __shared__ __align__(16) unsigned char smem[];

template <typename T>
struct SmRef {
    uint32_t addr;   // shared-window address in bytes
    __device__ __forceinline__ operator T() const {
        uint32_t v;
        if constexpr (sizeof(T) == 1) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
        else if constexpr (sizeof(T) == 2) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
        else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        return (T)v;
    }
    // stores take the low bits of a 32-bit register: no masking needed (st.shared.u8/u16 truncate)
    __device__ __forceinline__ void put(uint32_t v) const {
        if constexpr (sizeof(T) == 1) asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
        else if constexpr (sizeof(T) == 2) asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
        else asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
    }
    __device__ __forceinline__ uint32_t get() const {
        uint32_t v;
        if constexpr (sizeof(T) == 1) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
        else if constexpr (sizeof(T) == 2) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
        else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        return v;
    }
    __device__ __forceinline__ const SmRef &operator=(T x) const {
        const uint32_t v = (uint32_t)x;
        if constexpr (sizeof(T) == 1) asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
        else if constexpr (sizeof(T) == 2) asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
        else asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
        return *this;
    }
    __device__ __forceinline__ const SmRef &operator+=(T x) const { return *this = (T)((T) * this + x); }
    __device__ __forceinline__ const SmRef &operator&=(T x) const { return *this = (T)((T) * this & x); }
    __device__ __forceinline__ const SmRef &operator|=(T x) const { return *this = (T)((T) * this | x); }
};
__device__ __forceinline__ void sm_red_or(uint32_t addr, uint32_t bits) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(bits) : "memory");
}
#define SM8(off) (SmRef<unsigned char>{sbase + (uint32_t)(off)})
#define SM16(off) (SmRef<uint16_t>{sbase + (uint32_t)(off)})
#define SM32(off) (SmRef<uint32_t>{sbase + (uint32_t)(off)})
#define SM8X(addr) (SmRef<unsigned char>{(uint32_t)(addr)})    // absolute shared-window address
#define SM32X(addr) (SmRef<uint32_t>{(uint32_t)(addr)})

// T[c] += delta for every entry of one neighbour row (NR*32 ids, this lane's column starting at `row`);
// the row's first entry of lane 0 is the cell itself and takes delta0 instead (its own weight and, in
// full_3d, the occupied flag).  aT = shared address of the chain's table.  All lanes call.
template <int NR, typename TE>
__device__ __forceinline__ void table_row_add(uint32_t aT, const uint16_t *row, uint32_t delta, uint32_t delta0) {
    uint32_t c[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) c[r] = __ldg(row + r * 32);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const SmRef<TE> cell{aT + c[r] * (uint32_t)sizeof(TE)};
        cell.put(cell.get() + (r == 0 ? delta0 : delta));
    }
}

// NR == 0: row length known at run time only (replay kernels); otherwise compiled in.
template <int NR, typename TE>
__device__ __forceinline__ void table_lines_add(uint32_t aT, const uint16_t *row, int nr, uint32_t delta, uint32_t delta0) {
    if constexpr (NR > 0) {
        table_row_add<NR, TE>(aT, row, delta, delta0);
    } else {
        switch (nr) {
            case 1: table_row_add<1, TE>(aT, row, delta, delta0); break;
            case 2: table_row_add<2, TE>(aT, row, delta, delta0); break;
            case 3: table_row_add<3, TE>(aT, row, delta, delta0); break;
            case 4: table_row_add<4, TE>(aT, row, delta, delta0); break;
            case 5: table_row_add<5, TE>(aT, row, delta, delta0); break;
            case 6: table_row_add<6, TE>(aT, row, delta, delta0); break;
            case 7: table_row_add<7, TE>(aT, row, delta, delta0); break;
            default: table_row_add<8, TE>(aT, row, delta, delta0); break;
        }
    }
}

// base + idx * scale as ONE 64-bit multiply-add (the compiler otherwise widens, shifts and adds with carries)
template <typename T>
__device__ __forceinline__ T *ptr_mad(T *base, uint32_t idx, uint32_t scale) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(idx), "r"(scale), "l"((uint64_t)base));
    return reinterpret_cast<T *>(r);
}

// One lane group of LPC lanes (32 or 16) owns a chain; a warp carries 32/LPC chains.  Narrower
// groups waste fewer evaluated proposals per round (at 3 % acceptance a 16-lane group commits
// 12.9 of 16, a 32-lane group 20.7 of 32); committed moves are applied by the whole warp, one
// chain after the other, so the table update keeps 32 lanes busy either way.
template <bool FULL, bool REPLAY, bool EARLY, int NR, int LPC, int CN = 0, int HK = -1>
__global__ void __launch_bounds__(128, LPC == 32 ? MCQ_SPEC_MINB : 5) spec_kernel(const __grid_constant__ KArgs a) {
    constexpr unsigned FULLMASK = 0xffffffffu;
    constexpr int NF = FULL ? NFAM : NFAM - 1;
    // full_3d: 16-bit table entries whose top bit says "a queen stands here", so the occupancy test of a
    // candidate cell and its conflict count are one load; board mode needs no occupancy and keeps bytes
    using TE = std::conditional_t<FULL, uint16_t, unsigned char>;
    constexpr int OCC = FULL ? 0x8000 : 0, CNT = FULL ? 0x7fff : 0xff;
#define TBL(base, c) (SmRef<TE>{sbase + (uint32_t)(base) + (uint32_t)(c) * (uint32_t)sizeof(TE)})
    constexpr int CPW = 32 / LPC;   // chains per warp
    constexpr unsigned LMASK = LPC == 32 ? 0xffffffffu : ((1u << LPC) - 1u);
    int lane = threadIdx.x & 31;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    // keep both in registers: the compiler otherwise re-derives them (S2R + address-window arithmetic) every round
    asm volatile("" : "+r"(lane), "+r"(sbase));
    const int sub = lane & (LPC - 1), half = LPC == 32 ? 0 : lane / LPC;
    const int N = SpecGeom<FULL, CN>::n(a), Q = SpecGeom<FULL, CN>::q(a);
    const SLayout sl = SpecGeom<FULL, CN>::layout(a);
    const int state_bytes = FULL ? 3 * Q : Q;
    // HK: history kind compiled in: 0 none, 1 uint16, 3 none + cross-replica statistics (difference form); < 0: run time
    const int hist_kind = HK == 3 ? 0 : HK >= 0 ? HK : a.hist_kind;
    const bool stats_rt = HK < 0 && a.dsum_e != nullptr;   // generic kernels: statistics behind the rare branch

    // ---- CTA-shared geometry: shared-line bits at offset 0, cell -> wide id at sl.off_wide ----
    if (FULL) {
        for (int w = threadIdx.x; w < sl.cta_bytes / 4; w += blockDim.x) SM32(4 * w) = __ldg(a.geo + w);
        __syncthreads();
    }
    const int slab0 = (int)(threadIdx.x >> 5) * CPW;                      // first slab of this warp
    const int chain0 = a.chain_begin + (blockIdx.x * (blockDim.x >> 5)) * CPW + slab0;   // first chain of this warp
    if (chain0 >= a.n_chains) return;
    const int chain = chain0 + half;
    const bool live = chain < a.n_chains;
    const int chain_c = live ? chain : chain0;                            // for pointer set-up only

    const int sT = sl.cta_bytes + (slab0 + half) * sl.stride;   // table T[N^3 + 1]
    const int sP = sT + sl.off_state;   // board: heights u8; full_3d: u32 per queen = cell id | wide id << 16
    const int sW = sl.off_wide;
    const int L = sl.nbr_len, rounds = sl.rounds;
    const uint32_t wide_bias = (uint32_t)sl.wide_bias;   // (N-1)*(W^2+W+1): centres the wide-id difference in the LUT

    const uint16_t *nbr_lane = a.nbr + lane;   // this lane's column of the neighbour rows
    // what a queen adds to the entry of her own cell (slot 0 of the cell's row, i.e. lane 0's first entry):
    // herself once, and the occupied flag; every other entry of the row changes by one
    const uint32_t d0 = lane == 0 ? 1u + (uint32_t)OCC : 1u;

    // ---- build the slabs from the external states: one chain at a time, all 32 lanes ----
    int E = 0;
    for (int h = 0; h < CPW; ++h) {
        const int bT = sl.cta_bytes + (slab0 + h) * sl.stride, bP = bT + sl.off_state;
        for (int w = lane; w < sl.stride / 4; w += 32) SM32(bT + 4 * w) = 0u;
        __syncwarp();
        if (chain0 + h >= a.n_chains) continue;   // no chain: the slab stays zero (harmless to evaluate)
        const uint8_t *ext = a.state + (size_t)(chain0 + h) * state_bytes;
        for (int qi = lane; qi < Q; qi += 32) {
            if (FULL) {
                const int cid = ((int)ext[3 * qi] * N + (int)ext[3 * qi + 1]) * N + (int)ext[3 * qi + 2];
                SM32(bP + 4 * qi) = (uint32_t)cid | ((uint32_t)SM16(sW + 2 * cid) << 16);
            } else {
                SM8(bP + qi) = ext[qi];
            }
        }
        __syncwarp();
        for (int qi = 0; qi < Q; ++qi) {
            const int c = FULL ? (int)(SM32(bP + 4 * qi) & 0xffffu) : qi * N + (int)SM8(bP + qi);
            table_lines_add<NR, TE>(sbase + (uint32_t)bT, ptr_mad(nbr_lane, (uint32_t)c, 2u * (uint32_t)L), rounds, 1u, d0);
            __syncwarp();
        }
        int e = 0;
        for (int qi = lane; qi < Q; qi += 32) {
            const int c = FULL ? (int)(SM32(bP + 4 * qi) & 0xffffu) : qi * N + (int)SM8(bP + qi);
            e += ((int)(TE)TBL(bT, c) & CNT) - 1;
        }
        e = __reduce_add_sync(FULLMASK, e) >> 1;   // every attacking pair was counted from both ends
        if (half == h) E = e;
    }

    // ---- persistent record (every lane of a group holds its chain's scalars) ----
    // best_step, the open acceptance bin and its mark are only touched behind the rare branch: they live in the
    // slab's record words (sR + 0, 4, 8), not in registers
    const int sR = sT + sl.off_rec;
    int best = E, stale = 0, n_acc = 0;
    int done = a.t_end;
    {
        int best_step0 = 0, bin_mark0 = 0;
        if (live && a.t_begin != 0) { best_step0 = a.best_step[chain]; bin_mark0 = a.bin_mark[chain]; }
        if (sub == 0) { SM32(sR) = (uint32_t)best_step0; SM32(sR + 4) = (uint32_t)a.bin_at_begin; SM32(sR + 8) = (uint32_t)bin_mark0; }
        __syncwarp();
    }
    int t = live ? a.t_begin : a.t_end;
    if (live) {
        if (a.t_begin == 0) {
            if (sub == 0) {
                if (a.init_e) a.init_e[chain] = E;
                if (hist_kind == 1) reinterpret_cast<uint16_t *>(a.hist)[(size_t)chain * a.hist_pitch] = (uint16_t)E;
                else if (hist_kind == 2) reinterpret_cast<int *>(a.hist)[(size_t)chain * a.hist_pitch] = E;
                if (HK == 3 || stats_rt) stat_delta(a, a.group ? a.group[chain] : 0, 0, 0, E, 1);
            }
        } else {
            best = a.best_e[chain];
            stale = a.stale[chain];
            n_acc = a.n_acc[chain];
            const int sd = a.steps_done[chain];
            if (sd < a.t_begin) { done = sd; t = a.t_end; }   // stopped in an earlier launch
        }
    }
    const unsigned long long sd64 = a.seeds ? a.seeds[chain_c] : 0ull;
    const uint32_t key0 = (uint32_t)sd64, key1 = (uint32_t)(sd64 >> 32);
    const int grp = a.group ? a.group[chain_c] : 0;
    const float *beta_row = REPLAY ? nullptr : a.beta_c + (size_t)grp * a.n_steps;
    const double *beta64_row = REPLAY ? a.beta64 + (size_t)grp * a.n_steps : nullptr;
    const uint32_t *mv_row = REPLAY ? a.rmoves + (size_t)chain_c * a.n_steps : nullptr;
    const double *un_row = REPLAY ? a.runif + (size_t)chain_c * a.n_steps : nullptr;
    int next_edge = a.n_bins > 0 ? a.bin_starts[a.bin_at_begin + 1] : 0x7fffffff;
    // statistics in difference form (KArgs::dsum_e): this chain's group row of sum E; sum E^2 sits at a fixed distance
    // (HK == 3 is only dispatched when n_groups * (n_steps + 1) fits 31 bits)
    [[maybe_unused]] const uint32_t srow = HK == 3 ? (uint32_t)grp * (uint32_t)a.stat_pitch : 0u;
    const int sG = sT + sl.off_ring;   // ring of random words: 64 steps x 16 B
    [[maybe_unused]] int tfill = t;       // steps < tfill of this chain have their words in the ring
    [[maybe_unused]] uint32_t near = 0u;
    // history row, addressed by history index h = step + 1 (h_origin = index held by column 0)
    unsigned char *hrow = static_cast<unsigned char *>(a.hist) +
                          ((long long)chain_c * a.hist_pitch - a.h_origin) * (hist_kind == 1 ? 2 : 4);
    const bool want_abits = a.abits != nullptr;

    while (CPW == 1 ? t < a.t_end : __any_sync(FULLMASK, t < a.t_end)) {
        const bool active = CPW == 1 || t < a.t_end;
        const int rem = a.t_end - t;                        // >= 1 while active
        const bool valid = active && sub < rem;
        const int s = min(t + sub, a.t_end - 1);            // lanes past the end redo the last step, masked below

        // ---------------- this lane's proposal: step s against the current state ----------------
        uint32_t c0, c1, aux;   // old cell, new cell, and what the commit needs besides them
        int dE;
        bool accept;
        [[maybe_unused]] bool bad = false, near_flag = false;
        [[maybe_unused]] unsigned nearm = 0u, flipm = 0u;   // production: lanes of this group decided by the float64 rule / flipped by it
        if constexpr (REPLAY) {
            const uint32_t mv = mv_row[s];
            const double u64 = un_row[s], b64 = beta64_row[s];
            if (FULL) {
                uint32_t q = mv & 0xfff;
                const uint32_t i1 = (mv >> 12) & 63, j1 = (mv >> 18) & 63, k1 = (mv >> 24) & 63;
                bad = q >= (uint32_t)Q || i1 >= (uint32_t)N || j1 >= (uint32_t)N || k1 >= (uint32_t)N;
                c1 = bad ? 0u : (i1 * N + j1) * N + k1;
                if (bad) q = 0;
                const int v1 = (int)(TE)TBL(sT, c1);
                bad = bad || (v1 & OCC);
                const uint32_t p0 = SM32(sP + 4 * q), w1 = SM16(sW + 2 * c1);
                c0 = p0 & 0xffffu;
                const uint32_t e = w1 - (p0 >> 16) + wide_bias;
                dE = (v1 & CNT) - ((int)(TE)TBL(sT, c0) & CNT) + 1 - (int)((SM32(4 * (e >> 5)) >> (e & 31)) & 1u);
                aux = q | (w1 << 16);
            } else {
                uint32_t i0 = mv & 255, j0 = (mv >> 8) & 255, k1 = (mv >> 16) & 255;
                bad = i0 >= (uint32_t)N || j0 >= (uint32_t)N || k1 >= (uint32_t)N;
                if (bad) { i0 = j0 = k1 = 0; }
                const uint32_t ij = i0 * N + j0, k0 = SM8(sP + ij);
                bad = bad || (k1 == k0);
                c0 = ij * N + k0; c1 = ij * N + k1;
                dE = (int)(TE)TBL(sT, c1) - (int)(TE)TBL(sT, c0) + 1;
                aux = ij | (k1 << 16);
            }
            const double p = exp(-b64 * (double)dE);
            accept = !bad && u64 < fmin(1.0, p);
            near_flag = !bad && fabs(u64 - p) < 1e-6;
        } else {
            const float cb = __ldg(ptr_mad(beta_row, (uint32_t)s, 4u));
            // Random words of step s come from a 64-step ring in the slab: a round consumes only the
            // steps it commits, so one Philox4x32-10 call per lane refills LPC steps that are all used,
            // instead of recomputing the discarded lanes' words every round.
            if (active && tfill < t + LPC) {
                const Philox4 w = step_words<FULL>((uint32_t)(tfill + sub), key0, key1, (uint32_t)(N * N));
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (uint32_t)(sG + 16 * ((tfill + sub) & 63))),
                             "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                tfill += LPC;
            }
            __syncwarp();
            Philox4 r;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                         : "r"(sbase + (uint32_t)(sG + 16 * (s & 63))));
            if (FULL) {
                const uint32_t N3 = (uint32_t)(N * N * N);
                const uint32_t q = __umulhi(r.x, (uint32_t)Q);
                const uint32_t p0 = SM32(sP + 4 * q);
                // uniform over the empty cells: redraw while occupied (the queen's own cell counts,
                // experiments.py:226-231).  mulhi(word, N^3) == the (i,j,k) digits of anneal_kernel.
                // The first two candidates (words y, w) are resolved without a branch.
                c1 = __umulhi(r.y, N3);
                const uint32_t c1b = __umulhi(r.w, N3);
                int v1 = (int)(TE)TBL(sT, c1);
                const int v1b = (int)(TE)TBL(sT, c1b);
                if (v1 & OCC) { c1 = c1b; v1 = v1b; }   // two selects
                if (v1 & OCC) {
                    // third candidate: what is left of word x after the queen draw (its next mixed-radix
                    // digit), so that only (Q/N^3)^3 of the proposals pay for another Philox call
                    c1 = __umulhi(r.x * (uint32_t)Q, N3);
                    v1 = (int)(TE)TBL(sT, c1);
                    int e = 0;
                    while (v1 & OCC) {
                        const Philox4 r2 = chain_words((uint32_t)s, key0, key1, 1u + (uint32_t)(e >> 2));
                        const int sel = e & 3;
                        c1 = __umulhi(sel == 0 ? r2.x : sel == 1 ? r2.y : sel == 2 ? r2.z : r2.w, N3);
                        v1 = (int)(TE)TBL(sT, c1);
                        ++e;
                    }
                }
                const uint32_t w1 = SM16(sW + 2 * c1);
                c0 = p0 & 0xffffu;
                const uint32_t e = w1 - (p0 >> 16) + wide_bias;
                // the moving queen itself sits on a line through the new cell iff the cells share one
                dE = (v1 & CNT) - ((int)(TE)TBL(sT, c0) & CNT) + 1 - (int)((SM32(4 * (e >> 5)) >> (e & 31)) & 1u);
                aux = q | (w1 << 16);
            } else {
                const uint32_t ij = __umulhi(r.x, (uint32_t)(N * N));
                const uint32_t k0 = SM8(sP + ij);
                // uniform over the N-1 other heights (== the redraw loop of experiments.py:317-319)
                uint32_t k1 = k0 + 1u + __umulhi(r.y, (uint32_t)(N - 1));
                k1 -= (k1 >= (uint32_t)N) ? (uint32_t)N : 0u;
                c0 = ij * N + k0; c1 = ij * N + k1;
                dE = (int)(TE)TBL(sT, c1) - (int)(TE)TBL(sT, c0) + 1;
                aux = ij | (k1 << 16);
            }
            // Metropolis (experiments.py:238-239 / :326-327): float32 decision, retaken with the float64 rule when the
            // uniform word lies inside the float32 error band of the threshold (accept.cuh)
            bool near_band;
            metropolis_fast(dE, cb, r.z, a.band_abs, accept, near_band);
            if (__any_sync(FULLMASK, near_band)) {   // rare: about 1e-4 of the rounds
                bool flip = false;
                near_band = near_band && valid;
                if (near_band) {
                    const bool exact = metropolis_exact(a.sched, a.beta64, a.n_steps, a.group ? a.group[chain_c] : 0, key0, key1, s, dE, r.z);
                    flip = exact != accept;
                    accept = exact;
                }
                nearm = (__ballot_sync(FULLMASK, near_band) >> (half * LPC)) & LMASK;
                flipm = (__ballot_sync(FULLMASK, flip) >> (half * LPC)) & LMASK;
            }
        }
        accept = accept && valid;

        // ---------------- each group commits its first accepted proposal ----------------
        const unsigned acc_mask = (__ballot_sync(FULLMASK, accept) >> (half * LPC)) & LMASK;
        int first = acc_mask ? __ffs(acc_mask) - 1 : -1;
        int adv = !active ? 0 : first >= 0 ? first + 1 : min(LPC, rem);   // steps consumed by this round
        int adv_h = adv;                                                  // steps whose energy is appended to the history
        bool stop = false;
        const int src = half * LPC + max(first, 0);                       // the group's winning lane
        const int wdE = __shfl_sync(FULLMASK, dE, src);
        int E_new = first >= 0 ? E + wdE : E;
        bool improved = E_new < best;
        if constexpr (EARLY) {
            // experiments.py:343-353: the counter resets on a strict improvement only, and the
            // break happens before the history append of the stopping step
            if (active) {
                const int e_stop = max(a.patience - stale - 1, 0);   // rejected step at which patience runs out
                if (first < 0 || first > e_stop) {
                    if (e_stop < min(LPC, rem)) { stop = true; first = -1; adv = e_stop + 1; adv_h = e_stop; stale += e_stop + 1; E_new = E; improved = false; }
                    else stale += adv;
                } else {
                    stale = improved ? 0 : stale + first + 1;
                    if (stale >= a.patience) { stop = true; adv_h = first; }
                }
            }
        }
        if constexpr (REPLAY) {
            const unsigned committed = adv >= 32 ? FULLMASK : ((1u << adv) - 1u);
            const unsigned nearm = (__ballot_sync(FULLMASK, near_flag && valid) >> (half * LPC)) & LMASK & committed;
            const unsigned badm = (__ballot_sync(FULLMASK, bad && valid) >> (half * LPC)) & LMASK & committed;
            near += (uint32_t)__popc(nearm);
            if (badm && sub == 0) atomicAdd(a.replay_err, (unsigned)__popc(badm));
        }
        // history: steps t .. t+adv_h-1; all but an accepted last one keep the old energy
        // (two predicated stores rather than a branch)
        if (HK != 0) {
            const bool wr = sub < adv_h;
            const int v = (sub == first) ? E_new : E;
            uint16_t *h16 = reinterpret_cast<uint16_t *>(ptr_mad(hrow + 2, (uint32_t)s, 2u));
            int *h32 = reinterpret_cast<int *>(ptr_mad(hrow + 4, (uint32_t)s, 4u));
            if (wr && hist_kind == 1) *h16 = (uint16_t)v;
            if (HK < 0 && wr && hist_kind == 2) *h32 = v;
        }
        // ---------------- apply the committed moves: whole warp, one chain after the other ----------------
        const bool has = first >= 0;
        const uint32_t wc0 = __shfl_sync(FULLMASK, c0, src), wc1 = __shfl_sync(FULLMASK, c1, src);
        const uint32_t waux = __shfl_sync(FULLMASK, aux, src);
        // one bit (the group's lane 0) per committing chain; a single group needs no vote
        unsigned upd = CPW == 1 ? (unsigned)has : __ballot_sync(FULLMASK, has && sub == 0);
        while (upd) {
            const int hl = CPW == 1 ? 0 : __ffs(upd) - 1;
            upd &= upd - 1;
            uint32_t bc0 = wc0, bc1 = wc1, baux = waux;
            int bT = sT;
            if constexpr (CPW > 1) {
                bc0 = __shfl_sync(FULLMASK, wc0, hl); bc1 = __shfl_sync(FULLMASK, wc1, hl); baux = __shfl_sync(FULLMASK, waux, hl);
                bT = sl.cta_bytes + (slab0 + hl / LPC) * sl.stride;
            }
            const uint32_t aT = sbase + (uint32_t)bT;
            if constexpr (NR > 0) {
                // both neighbour rows are fetched before the first phase so that only one L2 round trip is exposed
                uint32_t ra[NR], rb[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) ra[r] = __ldg(ptr_mad(nbr_lane, bc0, 2u * (uint32_t)L) + r * 32);
#pragma unroll
                for (int r = 0; r < NR; ++r) rb[r] = __ldg(ptr_mad(nbr_lane, bc1, 2u * (uint32_t)L) + r * 32);
                // the entries of one row are distinct cells (pads share a scratch entry nobody reads), so all
                // loads of a phase are issued before its stores: one shared-memory round trip per phase, not NR
                uint32_t va[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) { ra[r] = aT + ra[r] * (uint32_t)sizeof(TE); va[r] = SmRef<TE>{ra[r]}.get(); }
#pragma unroll
                for (int r = 0; r < NR; ++r) SmRef<TE>{ra[r]}.put(va[r] - (r == 0 ? d0 : 1u));
                __syncwarp();
#pragma unroll
                for (int r = 0; r < NR; ++r) { rb[r] = aT + rb[r] * (uint32_t)sizeof(TE); va[r] = SmRef<TE>{rb[r]}.get(); }
#pragma unroll
                for (int r = 0; r < NR; ++r) SmRef<TE>{rb[r]}.put(va[r] + (r == 0 ? d0 : 1u));
            } else {
                table_lines_add<NR, TE>(aT, ptr_mad(nbr_lane, bc0, 2u * (uint32_t)L), rounds, 0u - 1u, 0u - d0);
                __syncwarp();
                table_lines_add<NR, TE>(aT, ptr_mad(nbr_lane, bc1, 2u * (uint32_t)L), rounds, 1u, d0);
            }
            // the moved queen's new position (lane 0 only; a predicated store, not a branch)
            {
                const uint32_t aP = aT + (uint32_t)sl.off_state;
                if (FULL) {
                    if (lane == 0) SM32X(aP + 4 * (baux & 0xffffu)).put(bc1 | (baux & 0xffff0000u));
                } else {
                    if (lane == 0) SM8X(aP + (baux & 0xffffu)).put(baux >> 16);
                }
            }
            __syncwarp();
        }
        // ---------------- bookkeeping; everything infrequent sits behind ONE branch ----------------
        const int n_before = n_acc;
        if constexpr (HK == 3) {
            // sum E / sum E^2 over the replicas of a group, difference form: an accepted move adds dE and
            // E_new^2 - E_old^2 to the column of the history index it produces (one lane, two reductions)
            if (has && sub == 0 && E_new != E) {
                const long long de = (long long)(E_new - E);
                unsigned long long *pe = a.dsum_e + (srow + (uint32_t)(t + first + 1));
                atomicAdd(pe, (unsigned long long)de);
                atomicAdd(pe + (a.dsum_e2 - a.dsum_e), (unsigned long long)(de * (long long)(E_new + E)));
            }
        }
        const int E_old = E;
        if (has) { E = E_new; ++n_acc; }
        const bool edge = active && t + adv - 1 >= next_edge;
        if (edge || improved || stop || (has && (want_abits || stats_rt)) || (!REPLAY && nearm != 0u)) {
            int bin = (int)SM32(sR + 4).get(), bin_mark = (int)SM32(sR + 8).get();
            if constexpr (!REPLAY) {
                if (nearm) {   // band decisions among the steps this round consumed
                    const unsigned committed = adv >= 32 ? FULLMASK : ((1u << adv) - 1u);
                    if (sub == 0 && a.near_cnt) atomicAdd(a.near_cnt + chain, (unsigned)__popc(nearm & committed));
                    if (sub == 0 && a.flip_cnt && (flipm & committed)) atomicAdd(a.flip_cnt + chain, (unsigned)__popc(flipm & committed));
                }
            }
            if (stats_rt && sub == 0) {
                if (has && E != E_old) stat_delta(a, a.group ? a.group[chain] : 0, (long long)t + first + 1, E_old, E, 0);
                if (stop) stat_delta(a, a.group ? a.group[chain] : 0, (long long)t + adv, E, 0, -1);
            }
            // acceptance bins: close every bin that ends at or before the last consumed step (an accept
            // of this round belongs to the bin its step lies in, i.e. the one that stays open)
            while (active && t + adv - 1 >= next_edge) {
                if (sub == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_before - bin_mark);
                bin_mark = n_before;
                ++bin;
                next_edge = a.bin_starts[bin + 1];
            }
            if (has) {
                const int ta = t + first;
                if (sub == 0 && want_abits) atomicOr(a.abits + (size_t)chain * a.abits_pitch + (ta >> 5), 1u << (ta & 31));
                if (improved) {
                    // snapshot: the state at the first visit of the minimum (strict <, :252 / :340)
                    best = E;
                    if (!stop && sub == 0) SM32(sR) = (uint32_t)(ta + 1);
                    uint8_t *bs = a.best_state + (size_t)chain * state_bytes;
                    if (FULL) {
                        for (int qi = sub; qi < Q; qi += LPC) {
                            const int c = (int)(SM32(sP + 4 * qi) & 0xffffu);
                            bs[3 * qi] = (uint8_t)(c / (N * N)); bs[3 * qi + 1] = (uint8_t)((c / N) % N); bs[3 * qi + 2] = (uint8_t)(c % N);
                        }
                    } else {
                        for (int c = sub; c < Q; c += LPC) bs[c] = SM8(sP + c);
                    }
                }
            }
            if (stop) {
                done = t + adv - 1;
                if (sub == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
            }
            if (edge && sub == 0) { SM32(sR + 4) = (uint32_t)bin; SM32(sR + 8) = (uint32_t)bin_mark; }
        }
        t = stop ? a.t_end : t + adv;   // a stopped chain is finished; the other groups of the warp go on
    }

    // ---------------- write the record back ----------------
    __syncwarp();
    if (!live) return;
    if (sub == 0) {
        const int bin = (int)SM32(sR + 4).get(), bin_mark = (int)SM32(sR + 8).get();
        if (a.t_end == a.n_steps && a.n_bins > 0 && a.acc_hist && done == a.t_end)
            a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
        a.cur_e[chain] = E;
        a.best_e[chain] = best;
        a.best_step[chain] = (int)SM32(sR).get();
        a.n_acc[chain] = n_acc;
        a.stale[chain] = stale;
        a.bin_mark[chain] = bin_mark;
        a.steps_done[chain] = done;
        if (REPLAY && a.near_cnt) a.near_cnt[chain] += near;
    }
    uint8_t *out = a.state + (size_t)chain * state_bytes;
    if (FULL) {
        for (int qi = sub; qi < Q; qi += LPC) {
            const int c = (int)(SM32(sP + 4 * qi) & 0xffffu);
            out[3 * qi] = (uint8_t)(c / (N * N)); out[3 * qi + 1] = (uint8_t)((c / N) % N); out[3 * qi + 2] = (uint8_t)(c % N);
        }
    } else {
        for (int c = sub; c < Q; c += LPC) out[c] = SM8(sP + c);
    }
}

#undef TBL
#undef SM8
#undef SM16
#undef SM32
#undef SM8X
#undef SM32X

}  // namespace mcq
