// Production form of the speculative conflict-table kernel (sm_100a): same slab, same random stream, same
// decisions and therefore the same trajectories as spec_kernel (spec.cuh), with the control flow arranged around
// what a round of a cold chain actually does -- nothing.
//
// At the acceptance rates where an anneal spends most of its steps (a few per cent and below) most rounds of LPC
// speculative proposals contain no accepted move.  spec_kernel runs every round through the commit, history,
// bookkeeping and bin-edge logic; here a round whose vote finds neither an accepted proposal nor a proposal inside
// the float32 error band advances the chain and loops -- everything else sits behind that one warp-uniform branch:
//
//   * acceptance-bin edges and the end of the launch are not tested per round: a chain runs in SPANS that end at
//     the next bin edge (or the end of the launch); lanes past the end of the span are masked by one compare, and the
//     bins are closed between spans;
//   * the winner's proposal is shuffled out, the first accepted lane located and the new energy formed only in
//     rounds that commit;
//   * the schedule table is padded, so the step a lane evaluates needs no clamp.
//
// Serves production runs without early stop (the board patience changes what a round may consume: spec_kernel) and
// without a recorded stream (replay: spec_kernel), for history kinds none / uint16 / statistics-only.
#pragma once
#include "launch.h"
#include "spec.cuh"

namespace mcq {

#define SM8(off) (SmRef<unsigned char>{sbase + (uint32_t)(off)})
#define SM16(off) (SmRef<uint16_t>{sbase + (uint32_t)(off)})
#define SM32(off) (SmRef<uint32_t>{sbase + (uint32_t)(off)})
#define SM8X(addr) (SmRef<unsigned char>{(uint32_t)(addr)})
#define SM32X(addr) (SmRef<uint32_t>{(uint32_t)(addr)})
#define TBL(base, c) (SmRef<TE>{sbase + (uint32_t)(base) + (uint32_t)(c) * (uint32_t)sizeof(TE)})

#ifndef MCQ_FAST16_MINB
#define MCQ_FAST16_MINB 5
#endif

template <bool FULL, int NR, int LPC, int CN, int HK>
__global__ void __launch_bounds__(LPC == 32 ? 32 * MCQ_FAST_WARPS : 128, LPC == 32 ? MCQ_FAST_MINB : MCQ_FAST16_MINB) fast_kernel(const __grid_constant__ KArgs a) {
    static_assert(HK == 0 || HK == 1 || HK == 3, "history kinds: none, uint16, statistics");
    static_assert(NR > 0, "the neighbour-row length is compiled in");
    constexpr unsigned FULLMASK = 0xffffffffu;
    using TE = std::conditional_t<FULL, uint16_t, unsigned char>;
    constexpr int OCC = FULL ? 0x8000 : 0, CNT = FULL ? 0x7fff : 0xff;
    constexpr int CPW = 32 / LPC;
    constexpr unsigned LMASK = LPC == 32 ? 0xffffffffu : ((1u << LPC) - 1u);
    constexpr int RING = 64;        // steps of random words kept per chain (spec_layout reserves 64 x 16 B)
    int lane = threadIdx.x & 31;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("" : "+r"(lane), "+r"(sbase));
    const int sub = lane & (LPC - 1), half = LPC == 32 ? 0 : lane / LPC;
    const int N = SpecGeom<FULL, CN>::n(a), Q = SpecGeom<FULL, CN>::q(a);
    const SLayout sl = SpecGeom<FULL, CN>::layout(a);
    const int state_bytes = FULL ? 3 * Q : Q;

    // ---- CTA-shared geometry: shared-line bits at offset 0, cell -> wide id at sl.off_wide ----
    if (FULL) {
        for (int w = threadIdx.x; w < sl.cta_bytes / 4; w += blockDim.x) SM32(4 * w) = __ldg(a.geo + w);
        __syncthreads();
    }
    const int slab0 = (int)(threadIdx.x >> 5) * CPW;
    const int chain0 = a.chain_begin + (blockIdx.x * (blockDim.x >> 5)) * CPW + slab0;
    if (chain0 >= a.n_chains) return;
    const int chain = chain0 + half;
    const bool live = chain < a.n_chains;
    const int chain_c = live ? chain : chain0;

    const int sT = sl.cta_bytes + (slab0 + half) * sl.stride;
    const int sP = sT + sl.off_state;
    const int sW = sl.off_wide;
    const int L = sl.nbr_len;
    const uint32_t wide_bias = (uint32_t)sl.wide_bias;
    const uint16_t *nbr_lane = a.nbr + lane;
    const uint32_t d0 = lane == 0 ? 1u + (uint32_t)OCC : 1u;

    // ---- build the slabs from the external states (as spec_kernel does) ----
    int E = 0;
    for (int h = 0; h < CPW; ++h) {
        const int bT = sl.cta_bytes + (slab0 + h) * sl.stride, bP = bT + sl.off_state;
        for (int w = lane; w < sl.stride / 4; w += 32) SM32(bT + 4 * w) = 0u;
        __syncwarp();
        if (chain0 + h >= a.n_chains) continue;
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
            table_row_add<NR, TE>(sbase + (uint32_t)bT, ptr_mad(nbr_lane, (uint32_t)c, 2u * (uint32_t)L), 1u, d0);
            __syncwarp();
        }
        int e = 0;
        for (int qi = lane; qi < Q; qi += 32) {
            const int c = FULL ? (int)(SM32(bP + 4 * qi) & 0xffffu) : qi * N + (int)SM8(bP + qi);
            e += ((int)(TE)TBL(bT, c) & CNT) - 1;
        }
        e = __reduce_add_sync(FULLMASK, e) >> 1;
        if (half == h) E = e;
    }

    // ---- persistent record: what a round needs in registers, the rest in the slab's record words ----
    // sR + 0: step of the best energy, + 4: open acceptance bin, + 8: accepted moves when it opened
    const int sR = sT + sl.off_rec;
    int best = E, n_acc = 0;
    int t = live ? a.t_begin : a.t_end;
    const int grp = a.group ? a.group[chain_c] : 0;
    {
        int best_step0 = 0, bin_mark0 = 0;
        if (live) {
            if (a.t_begin == 0) {
                if (sub == 0) {
                    if (a.init_e) a.init_e[chain] = E;
                    if (HK == 1) reinterpret_cast<uint16_t *>(a.hist)[(size_t)chain * a.hist_pitch] = (uint16_t)E;
                    if (HK == 3) stat_delta(a, grp, 0, 0, E, 1);
                }
            } else {
                best = a.best_e[chain];
                n_acc = a.n_acc[chain];
                best_step0 = a.best_step[chain];
                bin_mark0 = a.bin_mark[chain];
            }
        }
        if (sub == 0) { SM32(sR) = (uint32_t)best_step0; SM32(sR + 4) = (uint32_t)a.bin_at_begin; SM32(sR + 8) = (uint32_t)bin_mark0; }
        __syncwarp();
    }
    const unsigned long long sd64 = a.seeds ? a.seeds[chain_c] : 0ull;
    const uint32_t key0 = (uint32_t)sd64, key1 = (uint32_t)(sd64 >> 32);
    const float *beta_row = a.beta_c + (size_t)grp * a.n_steps;
    asm volatile("" : "+l"(beta_row));   // kept in registers: re-deriving the row address costs four instructions a round
    // a chain runs in spans that end where an acceptance bin ends (or the launch does): no per-round edge test
    int span_end = a.n_bins > 0 ? min(a.t_end, a.bin_starts[a.bin_at_begin + 1]) : a.t_end;
    if (!live) span_end = a.t_end;
    const int sG = sT + sl.off_ring;
    int tfill = t;
    // statistics row of this chain's group (sum E; the sum E^2 row sits at a fixed distance)
    [[maybe_unused]] unsigned long long *srow_e = HK == 3 ? a.dsum_e + (size_t)grp * (size_t)a.stat_pitch : nullptr;
    [[maybe_unused]] const long long s2_off = HK == 3 ? (long long)(a.dsum_e2 - a.dsum_e) : 0;
    [[maybe_unused]] unsigned char *hrow = static_cast<unsigned char *>(a.hist) + ((long long)chain_c * a.hist_pitch - a.h_origin) * 2;
    const bool want_abits = a.abits != nullptr;

    for (;;) {
        // ---------------- spans: close the bins that end here, open the next span, or leave ----------------
        {
            if (t >= span_end && t < a.t_end) {
                int bin = (int)SM32(sR + 4).get(), bin_mark = (int)SM32(sR + 8).get();
                int edge = a.bin_starts[bin + 1];
                while (t >= edge) {   // t == edge: the bin is complete (empty bins close at once)
                    if (sub == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
                    bin_mark = n_acc;
                    ++bin;
                    edge = a.bin_starts[bin + 1];
                }
                span_end = min(a.t_end, edge);
                if (sub == 0) { SM32(sR + 4) = (uint32_t)bin; SM32(sR + 8) = (uint32_t)bin_mark; }
            }
            if (CPW == 1 ? t >= a.t_end : __all_sync(FULLMASK, t >= a.t_end)) break;
        }
        // ---------------- the rounds of this span: the loop condition is the only per-round edge test ----------------
        do {
        const int s = t + sub;                    // lanes past the end of the span evaluate too; they are masked below
        const bool valid = s < span_end;          // (a finished chain of a two-chain warp has no valid lane)

        // ---------------- this lane's proposal: step s against the current state ----------------
        const float cb = __ldg(ptr_mad(beta_row, (uint32_t)s, 4u));
        if (tfill < t + LPC && (CPW == 1 || t < a.t_end)) {   // (one chain per warp: t < span_end <= t_end inside this loop)
            if (FULL) {
                const Philox4 w = chain_words((uint32_t)(tfill + sub), key0, key1, PHILOX_STREAM_STEP);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (uint32_t)(sG + 16 * ((tfill + sub) & (RING - 1)))),
                             "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
            } else {   // a board step needs two words: Philox2x32-10, eight bytes of the ring
#ifdef MCQ_DBG_BOARD_PHILOX4   // sensitivity probe only (not a product configuration): the cost of the four-word generator
                const Philox4 w4 = chain_words((uint32_t)(tfill + sub), key0, key1, PHILOX_STREAM_STEP);
                Philox2 w; w.x = w4.x ^ w4.y; w.z = w4.z ^ w4.w;
#else
                const Philox2 w = board_step_words((uint32_t)(tfill + sub), key0, key1);
#endif
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sbase + (uint32_t)(sG + 8 * ((tfill + sub) & (RING - 1)))),
                             "r"(w.x), "r"(w.z) : "memory");
            }
            tfill += LPC;
        }
        __syncwarp();
        Philox4 r;
        if (FULL) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                               : "r"(sbase + (uint32_t)(sG + 16 * (s & (RING - 1)))));
        else asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.z) : "r"(sbase + (uint32_t)(sG + 8 * (s & (RING - 1)))));
        uint32_t c0, c1, aux_lo, aux_hi;   // what the commit needs besides the two cells: packed only in rounds that commit
        int dE;
        if (FULL) {
            const uint32_t N3 = (uint32_t)(N * N * N);
            const uint32_t q = __umulhi(r.x, (uint32_t)Q);
            const uint32_t p0 = SM32(sP + 4 * q);
            c1 = __umulhi(r.y, N3);
            const uint32_t c1b = __umulhi(r.w, N3);
            int v1 = (int)(TE)TBL(sT, c1);
            const int v1b = (int)(TE)TBL(sT, c1b);
            if (v1 & OCC) { c1 = c1b; v1 = v1b; }
            if (v1 & OCC) {
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
#ifdef MCQ_DBG_NOLUT    // sensitivity probe only: no shared-line correction (wrong delta-E on ~17 % of the proposals)
            dE = (v1 & CNT) - ((int)(TE)TBL(sT, c0) & CNT) + 1 - (int)(e & 1u);
#else
            dE = (v1 & CNT) - ((int)(TE)TBL(sT, c0) & CNT) + 1 - (int)((SM32(4 * (e >> 5)) >> (e & 31)) & 1u);
#endif
            aux_lo = q; aux_hi = w1;
        } else {
            const unsigned long long col = (unsigned long long)r.x * (uint32_t)(N * N);   // column digit, then the height offset
            const uint32_t ij = (uint32_t)(col >> 32);
            const uint32_t k0 = SM8(sP + ij);
            uint32_t k1 = k0 + 1u + __umulhi((uint32_t)col, (uint32_t)(N - 1));
            k1 -= (k1 >= (uint32_t)N) ? (uint32_t)N : 0u;
            c0 = ij * N + k0; c1 = ij * N + k1;
            dE = (int)(TE)TBL(sT, c1) - (int)(TE)TBL(sT, c0) + 1;
            aux_lo = ij; aux_hi = k1;
        }
        bool accept, near_band;
        metropolis_fast(dE, cb, r.z, a.band_abs, accept, near_band);
#ifdef MCQ_DBG_NONEAR   // sensitivity probe only (not a product configuration): float32 decisions, no band
        near_band = false;
#endif

        // ---------------- the vote: most rounds of a cold chain end here ----------------
        unsigned acc_all = __ballot_sync(FULLMASK, accept && valid);
        const unsigned near_all = __ballot_sync(FULLMASK, near_band && valid);
        if ((acc_all | near_all) == 0u) {
            if (HK == 1 && valid) *reinterpret_cast<uint16_t *>(ptr_mad(hrow + 2, (uint32_t)s, 2u)) = (uint16_t)E;
            t = min(t + LPC, span_end);
            continue;
        }
        // ---------------- something happened ----------------
        unsigned nearm = 0u, flipm = 0u;
        if (near_all) {   // some lane sits inside the float32 error band: the float64 rule decides (accept.cuh)
            bool flip = false;
            near_band = near_band && valid;
            if (near_band) {
                const bool exact = metropolis_exact(a.sched, a.beta64, a.n_steps, grp, key0, key1, s, dE, r.z);
                flip = exact != accept;
                accept = exact;
            }
            acc_all = __ballot_sync(FULLMASK, accept && valid);
            nearm = (near_all >> (half * LPC)) & LMASK;
            flipm = (__ballot_sync(FULLMASK, flip) >> (half * LPC)) & LMASK;
        }
        const unsigned acc_mask = (acc_all >> (half * LPC)) & LMASK;
        const int f = __ffs(acc_mask);                       // 0: this chain commits nothing
        const bool has = f != 0;
        const int first = f - 1;
        const int adv = has ? f : max(min(LPC, span_end - t), 0);   // steps consumed by this round
        const int src = half * LPC + max(first, 0);
        const int wdE = __shfl_sync(FULLMASK, dE, src);
        const uint32_t wc0 = __shfl_sync(FULLMASK, c0, src), wc1 = __shfl_sync(FULLMASK, c1, src);
        const uint32_t waux = __shfl_sync(FULLMASK, aux_lo | (aux_hi << 16), src);
        const int E_new = has ? E + wdE : E;
        if (HK == 1 && sub < adv) *reinterpret_cast<uint16_t *>(ptr_mad(hrow + 2, (uint32_t)s, 2u)) = (uint16_t)((sub == first) ? E_new : E);
        // ---------------- apply the committed moves: whole warp, one chain after the other ----------------
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
            uint32_t ra[NR], rb[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) ra[k] = __ldg(ptr_mad(nbr_lane, bc0, 2u * (uint32_t)L) + k * 32);
#pragma unroll
            for (int k = 0; k < NR; ++k) rb[k] = __ldg(ptr_mad(nbr_lane, bc1, 2u * (uint32_t)L) + k * 32);
            uint32_t va[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) { ra[k] = aT + ra[k] * (uint32_t)sizeof(TE); va[k] = SmRef<TE>{ra[k]}.get(); }
#pragma unroll
            for (int k = 0; k < NR; ++k) SmRef<TE>{ra[k]}.put(va[k] - (k == 0 ? d0 : 1u));
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NR; ++k) { rb[k] = aT + rb[k] * (uint32_t)sizeof(TE); va[k] = SmRef<TE>{rb[k]}.get(); }
#pragma unroll
            for (int k = 0; k < NR; ++k) SmRef<TE>{rb[k]}.put(va[k] + (k == 0 ? d0 : 1u));
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
        // ---------------- bookkeeping of the committing chains ----------------
        if (has) {
            const int ta = t + first;
            if (HK == 3 && sub == 0 && E_new != E) {
                // sum E / sum E^2 over the replicas of a group, difference form (KArgs::dsum_e)
                const long long de = (long long)(E_new - E);
                unsigned long long *pe = srow_e + (ta + 1);
                atomicAdd(pe, (unsigned long long)de);
                atomicAdd(pe + s2_off, (unsigned long long)(de * (long long)(E_new + E)));
            }
            E = E_new;
            ++n_acc;
            if (sub == 0 && want_abits) atomicOr(a.abits + (size_t)chain * a.abits_pitch + (ta >> 5), 1u << (ta & 31));
            if (E < best) {
                // snapshot: the state at the first visit of the minimum (strict <, experiments.py:252 / :340)
                best = E;
                if (sub == 0) SM32(sR) = (uint32_t)(ta + 1);
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
        if (nearm) {   // band decisions among the steps this round consumed
            const unsigned committed = adv >= 32 ? FULLMASK : ((1u << adv) - 1u);
            if (sub == 0 && a.near_cnt && (nearm & committed)) atomicAdd(a.near_cnt + chain, (unsigned)__popc(nearm & committed));
            if (sub == 0 && a.flip_cnt && (flipm & committed)) atomicAdd(a.flip_cnt + chain, (unsigned)__popc(flipm & committed));
        }
        t += adv;
        } while (CPW == 1 ? t < span_end : __all_sync(FULLMASK, t < span_end));
    }

    // ---------------- write the record back ----------------
    __syncwarp();
    if (!live) return;
    if (sub == 0) {
        const int bin = (int)SM32(sR + 4).get(), bin_mark = (int)SM32(sR + 8).get();
        if (a.t_end == a.n_steps && a.n_bins > 0 && a.acc_hist && bin < a.n_bins)
            a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
        a.cur_e[chain] = E;
        a.best_e[chain] = best;
        a.best_step[chain] = (int)SM32(sR).get();
        a.n_acc[chain] = n_acc;
        a.stale[chain] = 0;
        a.bin_mark[chain] = bin_mark;
        a.steps_done[chain] = a.t_end;
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
