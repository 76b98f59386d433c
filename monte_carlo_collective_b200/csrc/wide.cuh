// Speculative CTA-per-chain annealing kernel on line counters (sm_100a): the large-board path.
//
// Boards beyond the conflict-table kernel (N > 18 full_3d, N > 21 board; C5 is N = 64 with 4096 queens)
// keep one uint8 counter per attack line (anneal.cuh) -- 105.6 KB at N = 64 with the folded space-diagonal
// families, i.e. two chains per SM.  A single warp stepping through such a chain is latency-bound, and one thread
// per chain with the counters in global memory (anneal_kernel<1> + gslab) is bound by 25 random HBM sectors per
// proposal.  Here the whole CTA (1, 2, 4 or 8 warps) works on one chain, the same way the lanes of a warp do in
// spec.cuh:
//
//   * thread l evaluates the proposal of step t+l against the current state (Philox words of step s depend on
//     (seed, s) only and are kept in a ring of 2*blockDim steps); delta-E is 2 x 13 (12) counter loads;
//   * chains with early stop commit the FIRST accepted proposal of the round: it is found with
//     one `redux.min` per warp and a minimum over the warps' candidates in shared memory; all steps before it were
//     rejected and do not change the state, so the sequential chain of experiments.py:218-258 / :308-355 is
//     reproduced exactly.  The move is applied by two warps side by side: lane f of the winner's warp updates the
//     two counters of family f, a lane of the next warp the state, the occupancy and the best-state journal;
//     two __syncthreads per round;
//   * chains without early stop commit EVERY accepted proposal of a round that the earlier commits of the
//     round cannot have touched (a geometric test on the published moves; the commit block below): a 64-step round
//     retires ~48 steps and 2.6 accepted moves at the acceptance rates of an N = 64 anneal instead of ~17 and one;
//   * the number of threads that evaluate (`width`, 32..blockDim in whole warps) follows the steps a round
//     consumes: a hot single-commit chain commits after a handful of steps, so evaluating 256 of them would only
//     queue up shared-memory traffic.  The trajectory does not depend on the width.
//
// Random stream, proposal rule and Metropolis test are those of anneal_kernel, bit for bit: the same seeds give
// the same trajectory on either kernel (tests/test_gpu_production.py).
#pragma once
#include "launch.h"
#include "spec.cuh"

namespace mcq {

// WIDE_THREADS, WIDE_JCAP, WIDE_XCH_BYTES: launch.h (the host sizes shared memory with them);
// the ring of random words keeps 2 * blockDim steps

template <bool FULL, bool EARLY, int NT>
__global__ void __launch_bounds__(NT, 1) wide_kernel(const __grid_constant__ KArgs a) {
    constexpr unsigned FULLMASK = 0xffffffffu;
    constexpr int NW = NT / 32, RING = 2 * NT;
    // a single-warp CTA needs no block barrier and no exchange through shared memory
    auto cta_sync = [] { if constexpr (NT == 32) __syncwarp(); else __syncthreads(); };
    constexpr int F0 = FULL ? 0 : 1;              // board mode has no (i,j) column family
    constexpr int NONE = 0x7fffffff;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chain = a.chain_begin + blockIdx.x;
    if (chain >= a.n_chains) return;
    const int N = a.N;
    const int pos32 = a.lay.pos32;

    uint8_t *cnt = smem;
    unsigned char *st = smem + a.lay.off_state;
    uint32_t *occ = reinterpret_cast<uint32_t *>(smem + a.lay.off_occ);
    uint8_t *best_out = a.best_state + (size_t)chain * a.state_bytes;   // state at the best energy: kept in global memory, external format
    uint4 *ring = reinterpret_cast<uint4 *>(smem + a.w_ring);
    int *xch = reinterpret_cast<int *>(smem + a.w_xch);               // [NW] first accepting thread per warp, [NW] its delta-E, scratch
    int *jcount = xch + 3 * NW;                                       // entries in the journal; > WIDE_JCAP: overflowed
    uint16_t *jrn = reinterpret_cast<uint16_t *>(xch + 3 * NW + 1);
    uint32_t *xmv = reinterpret_cast<uint32_t *>(xch) + 64;           // [NW][3]: old cell, new cell, queen of each warp's first acceptance
    // chains without early stop commit every accepted proposal of a round that the earlier commits of the round
    // cannot have touched (see the commit block): [NT] records of the accepting threads -- (move, delta-E) in board
    // mode, (old cell, new cell, delta-E, queen) in full_3d
    constexpr bool MULTI = !EARLY;
    [[maybe_unused]] uint32_t *xrec = reinterpret_cast<uint32_t *>(smem + a.w_xch + WIDE_XCH_BYTES);

    // ---- build the slab from the external state ----
    {
        uint32_t *W = reinterpret_cast<uint32_t *>(smem);
        for (int w = tid; w < a.lay.off_pkt / 4; w += NT) W[w] = 0u;
        __syncthreads();
        const uint8_t *ext = a.state + (size_t)chain * a.state_bytes;
        for (int qi = tid; qi < a.Q; qi += NT) {
            int i, j, k;
            if (FULL) {
                i = ext[3 * qi]; j = ext[3 * qi + 1]; k = ext[3 * qi + 2];
                store_pos(st, pos32, qi, pack_pos(pos32, i, j, k));
                const int cid = (i * N + j) * N + k;
                atomicOr(&occ[cid >> 5], 1u << (cid & 31));
            } else {
                i = qi / N; j = qi - i * N; k = ext[qi];
                st[qi] = (unsigned char)k;
            }
#pragma unroll
            for (int f = F0; f < NFAM; ++f) {
                const int idx = line_index(a.coef[f], a.csel[f], i, j, k);
                atomicAdd(&W[idx >> 2], 1u << ((idx & 3) * 8));
            }
        }
        __syncthreads();
    }
    int E;
    if (a.t_begin == 0) {
        const uint32_t *W = reinterpret_cast<const uint32_t *>(smem);
        int e = 0;
        for (int w = tid; w < a.lay.n_cnt / 4; w += NT) {
            const uint32_t v = W[w];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = (v >> (8 * b)) & 255;
                e += (c * (c - 1)) >> 1;
            }
        }
        e = __reduce_add_sync(FULLMASK, e);
        if (lane == 0) xch[2 * NW + warp] = e;
        __syncthreads();
        E = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) E += xch[2 * NW + w];
    } else {
        E = a.cur_e[chain];
    }

    // ---- persistent record (every thread holds the chain's scalars) ----
    int best = E, stale = 0, n_acc = 0, best_step = 0, bin_mark = 0;
    int done = a.t_end;
    int t = a.t_begin;
    // A new best copies only the state elements moved since the previous snapshot (hot chains set a new best
    // at almost every acceptance).  jfresh: the journal restarts at the next commit; the first snapshot of a
    // launch copies everything.
    bool jfresh = false;
    if (tid == 0) *jcount = WIDE_JCAP + 1;
    if (a.t_begin == 0) {
        if (tid == 0) {
            if (a.init_e) a.init_e[chain] = E;
            if (a.hist_kind == 1) reinterpret_cast<uint16_t *>(a.hist)[(size_t)chain * a.hist_pitch] = (uint16_t)E;
            else if (a.hist_kind == 2) reinterpret_cast<int *>(a.hist)[(size_t)chain * a.hist_pitch] = E;
            if (a.dsum_e) stat_delta(a, a.group ? a.group[chain] : 0, 0, 0, E, 1);
        }
    } else {
        best = a.best_e[chain];
        stale = a.stale[chain];
        n_acc = a.n_acc[chain];
        best_step = a.best_step[chain];
        bin_mark = a.bin_mark[chain];
        const int sd = a.steps_done[chain];
        if (sd < a.t_begin) { done = sd; t = a.t_end; }   // stopped in an earlier launch
    }
    const unsigned long long sd64 = a.seeds ? a.seeds[chain] : 0ull;
    const uint32_t key0 = (uint32_t)sd64, key1 = (uint32_t)(sd64 >> 32);
    const int grp = a.group ? a.group[chain] : 0;
    const float *beta_row = a.beta_c + (size_t)grp * a.n_steps;
    uint32_t near = 0u, flips = 0u;   // this thread's band decisions among the consumed steps
    int bin = a.bin_at_begin;
    int next_edge = a.n_bins > 0 ? a.bin_starts[bin + 1] : NONE;
    int tfill = t;
    int width = NT < 64 ? NT : 64;   // threads that evaluate a step this round (whole warps)
    unsigned char *hrow = static_cast<unsigned char *>(a.hist) +
                          ((long long)chain * a.hist_pitch - a.h_origin) * (a.hist_kind == 1 ? 2 : 4);
    uint32_t *abits_row = a.abits ? a.abits + (size_t)chain * a.abits_pitch : nullptr;

    while (t < a.t_end) {
        // ---------------- random words: refill the ring when this round would read past it ----------------
        if (tfill < t + NT) {
            const Philox4 w = step_words<FULL>((uint32_t)(tfill + tid), key0, key1, (uint32_t)(N * N));
            ring[(tfill + tid) & (RING - 1)] = make_uint4(w.x, w.y, w.z, w.w);
            tfill += NT;
            cta_sync();
        }
        if constexpr (MULTI) {
            // a round of the multi-commit form never crosses the edge of an acceptance bin: bins that ended are closed
            // here (every acceptance so far lies before t), and the round stops at the next edge
            while (t >= next_edge) {
                if (tid == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
                bin_mark = n_acc;
                ++bin;
                next_edge = a.bin_starts[bin + 1];
            }
        }
        const int rem = MULTI ? min(min(a.t_end, next_edge) - t, width) : min(a.t_end - t, width);
        const bool valid = tid < rem;
        const int s = min(t + tid, a.t_end - 1);          // threads past the end redo the last step, masked below
        const uint4 w = ring[s & (RING - 1)];
        const float cb = __ldg(beta_row + s);

        // ---------------- this thread's proposal: step s against the current state ----------------
        int i0 = 0, j0 = 0, k0c = 0, i1 = 0, j1 = 0, k1c = 0, qsel = 0, dE = 0;
        bool accept = false, was_near = false, was_flip = false;
        int io[NFAM], in[NFAM];   // the counters of the old and the new cell, per family
        [[maybe_unused]] int tries = 0, skip1 = -1;   // full_3d: occupied candidate cells skipped by this thread's draw, the first of them
        if (tid < width) {
        if constexpr (FULL) {
            qsel = (int)__umulhi(w.x, (uint32_t)a.Q);
            uint32_t word = w.y;
            tries = 0;
            while (true) {
                i1 = draw_digit(word, N); j1 = draw_digit(word, N); k1c = draw_digit(word, N);
                const int cid1 = (i1 * N + j1) * N + k1c;
                if (!((occ[cid1 >> 5] >> (cid1 & 31)) & 1u)) break;
                if (tries == 0) skip1 = cid1;   // the first occupied candidate (the commit block needs it)
                // occupied (the queen's own cell counts, experiments.py:230): redraw
                if (tries == 0) word = w.w;
                else if (tries == 1) word = w.x * (uint32_t)a.Q;   // what the queen draw left of word x
                else {
                    const int e = tries - 2;
                    const Philox4 r = chain_words((uint32_t)s, key0, key1, 1u + (uint32_t)(e >> 2));
                    const int sel = e & 3;
                    word = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
                }
                ++tries;
            }
            unpack_pos(pos32, load_pos(st, pos32, qsel), i0, j0, k0c);
        } else {
            uint32_t word = w.x;
            i0 = draw_digit(word, N); j0 = draw_digit(word, N);
            k0c = st[i0 * N + j0];
            // uniform over the N-1 other heights (== the redraw loop of experiments.py:317-319)
            k1c = k0c + 1 + (int)__umulhi(w.y, (uint32_t)(N - 1));
            k1c -= (k1c >= N) ? N : 0;
            i1 = i0; j1 = j0;
        }
        // delta-E from the line counters: old_conf = sum(co - 1), new_conf = sum(cn) - [shared line]
        {
            auto both = [&](auto fc) {
                constexpr int f = decltype(fc)::value;
                if constexpr (f >= F0) {
                    io[f] = line_index_f<f>(a, i0, j0, k0c);
                    // a board move changes k only: the new cell's line of a family is the old one's shifted along k
                    // (the fold of the space-diagonal families depends on i and j alone)
                    if constexpr (FULL) in[f] = line_index_f<f>(a, i1, j1, k1c);
                    else in[f] = io[f] + a.coef[f].z * (k1c - k0c);
                }
            };
            both(std::integral_constant<int, 0>{}); both(std::integral_constant<int, 1>{}); both(std::integral_constant<int, 2>{});
            both(std::integral_constant<int, 3>{}); both(std::integral_constant<int, 4>{}); both(std::integral_constant<int, 5>{});
            both(std::integral_constant<int, 6>{}); both(std::integral_constant<int, 7>{}); both(std::integral_constant<int, 8>{});
            both(std::integral_constant<int, 9>{}); both(std::integral_constant<int, 10>{}); both(std::integral_constant<int, 11>{});
            both(std::integral_constant<int, 12>{});
#pragma unroll
            for (int f = F0; f < NFAM; ++f) {
                const int co = cnt[io[f]], cn = cnt[in[f]];
                // a board move changes k only and every family but the dropped (i,j) one depends on k:
                // the old and the new cell never share a line
                if constexpr (FULL) dE += (io[f] != in[f]) ? (cn - co + 1) : 0;
                else dE += cn - co + 1;
            }
        }
        // Metropolis (experiments.py:238-239 / :326-327): u < exp(-beta dE), u = word / 2^32
        bool near_band;
        metropolis_fast(dE, cb, w.z, a.band_abs, accept, near_band);
        if (near_band && valid) {   // inside the float32 error band: the float64 rule decides (accept.cuh)
            const bool exact = metropolis_exact(a.sched, a.beta64, a.n_steps, grp, key0, key1, s, dE, w.z);
            was_near = true;
            was_flip = exact != accept;
            accept = exact;
        }
        accept = accept && valid;
        }

        if constexpr (MULTI) {
            // ---------------- the CTA commits every accepted proposal the earlier commits cannot have touched ----------------
            // A proposal reads 24 counters and its column's height.  A committed move changes the 24 counters of the
            // lines through its old and its new cell, and one height; so a LATER thread's delta-E -- and with it its
            // accept decision, a function of (delta-E, step) alone -- stands unless one of its two cells lies on a line
            // (of the 12 counted families) with one of the move's two cells, or in the move's column: about 1 % of the
            // later threads per commit at N = 64.  All accepting threads publish their move; every thread tests its
            // cells against the moves accepted before it; the round ends before the first thread that is touched (it is
            // evaluated again next round) and every accepted move before that thread is committed -- they touch
            // disjoint counters and columns, so they are applied side by side.  The chain is the sequential one of
            // experiments.py:308-355, step for step; a round retires several accepted moves instead of one.
            auto touches = [&](uint32_t mv) -> bool {   // does the move (i | j << 8 | old k << 16 | new k << 24) touch this thread's two cells?
                const int di = abs(i0 - (int)(mv & 255u)), dj = abs(j0 - (int)((mv >> 8) & 255u));
                const int c0 = (int)((mv >> 16) & 255u), c1 = (int)(mv >> 24);
                const int mag = max(di, dj);   // a line joins the two columns only if the nonzero ones of di, dj are equal
                const bool joined = (di == 0) | (dj == 0) | (di == dj);
                const int d00 = abs(k0c - c0), d01 = abs(k0c - c1), d10 = abs(k1c - c0), d11 = abs(k1c - c1);
                const bool on_line = (d00 == 0) | (d00 == mag) | (d01 == 0) | (d01 == mag) | (d10 == 0) | (d10 == mag) | (d11 == 0) | (d11 == mag);
                return joined & (on_line | (mag == 0));   // (mag == 0: the move is in this thread's column)
            };
            // full_3d: a committed move (queen from cell A to cell B) touches this thread if one of its two cells is
            // collinear (13 families; the same cell counts) with A or B -- that also covers "my queen moved" and "my new
            // cell got occupied" -- or if its draw skipped an occupied candidate that the move has vacated (exact for one
            // skipped candidate; a thread that skipped several counts as touched by any move)
            auto collinear = [](int ax, int ay, int az, int bx, int by, int bz) -> bool {
                const int dx = abs(ax - bx), dy = abs(ay - by), dz = abs(az - bz);
                const int m = max(dx, max(dy, dz));
                return ((dx == 0) | (dx == m)) & ((dy == 0) | (dy == m)) & ((dz == 0) | (dz == m));
            };
            auto touches_full = [&](uint32_t A, uint32_t B) -> bool {
                const int ax = A & 255u, ay = (A >> 8) & 255u, az = (A >> 16) & 255u;
                const int bx = B & 255u, by = (B >> 8) & 255u, bz = (B >> 16) & 255u;
                bool tch = collinear(i0, j0, k0c, ax, ay, az) | collinear(i0, j0, k0c, bx, by, bz) |
                           collinear(i1, j1, k1c, ax, ay, az) | collinear(i1, j1, k1c, bx, by, bz);
                tch |= (tries >= 2) | (tries == 1 && skip1 == (ax * N + ay) * N + az);
                return tch;
            };
            if constexpr (NT == 32 && !FULL) {
                // one warp per chain (boards up to N = 34, where lines are dense and a round rarely gets far past a
                // commit): the commits are taken one at a time, in step order, without any shared-memory exchange
                const uint32_t my_move = (uint32_t)(i0 | (j0 << 8) | (k0c << 16) | (k1c << 24));
                int L = rem, from = 0, e_run = 0, e_mine = 0;
                bool hit = false;
                for (;;) {
                    int mine = NONE;   // the earliest thread that is touched (low half 0) or accepts (delta-E, biased, in the low half)
                    if (tid >= from && tid < L) {
                        if (hit) mine = tid << 16;
                        else if (accept) mine = (tid << 16) | (dE + 0x8000);
                    }
                    const int cmin = __reduce_min_sync(FULLMASK, mine);
                    if (cmin == NONE) break;
                    const int c = cmin >> 16;
                    if ((cmin & 0xffff) == 0) { L = c; break; }
                    const int dEc = (cmin & 0xffff) - 0x8000;
                    const uint32_t mv = __shfl_sync(FULLMASK, my_move, c);
                    if (tid > c && !hit) hit = touches(mv);
                    if (lane >= 1 && lane < NFAM) {   // lane f owns family f
                        const int a0 = mv & 255u, b0 = (mv >> 8) & 255u, h0 = (mv >> 16) & 255u, h1 = mv >> 24;
                        const int4 cf = a.coef[lane], cs = a.csel[lane];
                        const int o = line_index(cf, cs, a0, b0, h0), n = line_index(cf, cs, a0, b0, h1);
                        const int vo = cnt[o], vn = cnt[n];
                        cnt[o] = (uint8_t)(vo - 1); cnt[n] = (uint8_t)(vn + 1);
                    }
                    const int E_old = E + e_run, E_new = E_old + dEc;
                    if (tid == c) {
                        const int col = i0 * N + j0;
                        const int jn = jfresh ? 0 : *jcount;
                        if (jn < WIDE_JCAP) jrn[jn] = (uint16_t)col;
                        *jcount = jn + 1;   // WIDE_JCAP + 1 and beyond: overflow, the next snapshot is a full copy
                        st[col] = (unsigned char)k1c;
                        if (a.dsum_e && dEc != 0) stat_delta(a, grp, (long long)t + c + 1, E_old, E_new, 0);
                        if (abits_row) atomicOr(abits_row + ((t + c) >> 5), 1u << ((t + c) & 31));
                    }
                    jfresh = false;
                    if (tid >= c) e_mine += dEc;
                    e_run += dEc;
                    ++n_acc;
                    __syncwarp();
                    if (E_new < best) {   // snapshot: the state at the first visit of the minimum (strict <, experiments.py:340)
                        best = E_new;
                        best_step = t + c + 1;
                        const int jn = *jcount;
                        if (jn <= WIDE_JCAP) {
                            for (int e = tid; e < jn; e += NT) best_out[jrn[e]] = st[jrn[e]];
                        } else {
                            for (int el = tid; el < a.Q; el += NT) best_out[el] = st[el];
                        }
                        jfresh = true;
                        __syncwarp();
                    }
                    from = c + 1;
                }
                const int adv = L;
                if (was_near && tid < adv) { ++near; flips += (uint32_t)was_flip; }
                if (tid < adv && a.hist_kind) {
                    const int v = E + e_mine;
                    if (a.hist_kind == 1) reinterpret_cast<uint16_t *>(hrow)[s + 1] = (uint16_t)v;
                    else reinterpret_cast<int *>(hrow)[s + 1] = v;
                }
                __syncwarp();   // counters and state are final before the next round reads them
                E += e_run;
                t += adv;
                if (adv * 2 > width) width = min(NT, width * 2);
                else if (adv * 8 < width) width = max(32, width >> 1);
                continue;
            }
            // [NT] records of the accepting threads
            uint2 *slot = reinterpret_cast<uint2 *>(xrec);   // board: (i | j << 8 | old k << 16 | new k << 24, delta-E)
            uint4 *slot4 = reinterpret_cast<uint4 *>(xrec);  // full_3d: (old cell, new cell, delta-E, queen), cells as i | j << 8 | k << 16
            if (accept) {
                if constexpr (FULL) slot4[tid] = make_uint4((uint32_t)(i0 | (j0 << 8) | (k0c << 16)), (uint32_t)(i1 | (j1 << 8) | (k1c << 16)), (uint32_t)dE, (uint32_t)qsel);
                else slot[tid] = make_uint2((uint32_t)(i0 | (j0 << 8) | (k0c << 16) | (k1c << 24)), (uint32_t)dE);
            }
            auto slot_dE = [&](int c) -> int { return FULL ? (int)slot4[c].z : (int)slot[c].y; };
            const unsigned wacc = __ballot_sync(FULLMASK, accept);
            unsigned am[NW];                                                      // accepting threads of the CTA, a word per warp
            if constexpr (NT > 32) {
                if (lane == 0) xch[warp] = (int)wacc;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < NW; ++w) am[w] = (unsigned)xch[w];
            } else {
                __syncwarp();
                am[0] = wacc;
            }
            // moves accepted before this thread: do they touch its cells?  (and what they add to the energy)
            bool hit = false;
            int e_before = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if (w > warp) break;
                unsigned m = w == warp ? (am[w] & ((1u << lane) - 1u)) : am[w];
                while (m) {
                    const int c = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    if constexpr (FULL) {
                        const uint4 mv = slot4[c];
                        hit |= touches_full(mv.x, mv.y);
                        e_before += (int)mv.z;
                    } else {
                        const uint2 mv = slot[c];
                        hit |= touches(mv.x);
                        e_before += (int)mv.y;
                    }
                }
            }
            const int jn0 = jfresh ? 0 : *jcount;   // journal entries so far (read before anyone appends: the barrier below orders it)
            // the round ends before the first touched thread
            int L = __reduce_min_sync(FULLMASK, (hit && tid < rem) ? tid : rem);
            if constexpr (NT > 32) {
                if (lane == 0) xch[NW + warp] = L;
                __syncthreads();
                L = __reduce_min_sync(FULLMASK, lane < NW ? xch[NW + lane] : rem);
            }
            const int adv = L;
            // the committed moves in step order: energies, best energy (every thread, uniform)
            int n_com = 0, e_run = 0, k_best = -1;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                unsigned m = am[w];
                if (w * 32 + 32 > L) m &= (w * 32 >= L) ? 0u : ((1u << (L - w * 32)) - 1u);
                am[w] = m;   // from here on: the committing threads
                while (m) {
                    const int c = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    e_run += slot_dE(c);
                    if (E + e_run < best) { best = E + e_run; best_step = t + c + 1; k_best = n_com; }
                    ++n_com;
                }
            }
            // counters: thread group g (one thread per counted family: 12 in board mode, 13 in full_3d) applies committed move g
            {
                constexpr int GS = NFAM - F0, GROUPS = NT / GS;
                const int grp12 = tid / GS, fam = tid - grp12 * GS + F0;
                for (int g0 = 0; g0 < n_com; g0 += GROUPS) {
                    const int g = g0 + grp12;
                    if (grp12 < GROUPS && g < n_com) {
                        // the g-th committing thread
                        int c = -1, left = g;
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            const int pc = __popc(am[w]);
                            if (c < 0 && left < pc) {
                                unsigned m = am[w];
                                for (int i = 0; i < left; ++i) m &= m - 1;   // (a handful of commits per round)
                                c = w * 32 + __ffs(m) - 1;
                            }
                            left -= pc;
                        }
                        const int4 cf = a.coef[fam], cs = a.csel[fam];
                        int o, n;
                        if constexpr (FULL) {
                            const uint4 mv = slot4[c];
                            o = line_index(cf, cs, (int)(mv.x & 255u), (int)((mv.x >> 8) & 255u), (int)((mv.x >> 16) & 255u));
                            n = line_index(cf, cs, (int)(mv.y & 255u), (int)((mv.y >> 8) & 255u), (int)((mv.y >> 16) & 255u));
                        } else {
                            const uint32_t mvx = slot[c].x;
                            const int a0 = mvx & 255u, b0 = (mvx >> 8) & 255u, h0 = (mvx >> 16) & 255u, h1 = mvx >> 24;
                            o = line_index(cf, cs, a0, b0, h0); n = line_index(cf, cs, a0, b0, h1);
                        }
                        if (o != n) {   // (board: always; full_3d: the old and the new cell may share a line)
                            const int vo = cnt[o], vn = cnt[n];
                            cnt[o] = (uint8_t)(vo - 1); cnt[n] = (uint8_t)(vn + 1);
                        }
                    }
                }
            }
            // state, journal, statistics, accept bitmap: every committing thread for itself
            const bool commits = accept && tid < L;
            int my_k = 0;   // index of this thread's commit in the round
#pragma unroll
            for (int w = 0; w < NW; ++w)
                if (w < warp) my_k += __popc(am[w]);
                else if (w == warp) my_k += __popc(am[w] & ((1u << lane) - 1u));
            if constexpr (NT == 32) __syncwarp();
            auto commit_state = [&](int k_lo, int k_hi, int jbase) {   // commits k_lo .. k_hi-1; journal slot = jbase + k
                if (commits && my_k >= k_lo && my_k < k_hi) {
                    const int jn = jbase + my_k;
                    if constexpr (FULL) {
                        if (jn < WIDE_JCAP) jrn[jn] = (uint16_t)qsel;
                        const int cid0 = (i0 * N + j0) * N + k0c, cid1 = (i1 * N + j1) * N + k1c;
                        atomicAnd(&occ[cid0 >> 5], ~(1u << (cid0 & 31)));   // (two commits of a round may share a word)
                        atomicOr(&occ[cid1 >> 5], 1u << (cid1 & 31));
                        store_pos(st, pos32, qsel, pack_pos(pos32, i1, j1, k1c));
                    } else {
                        const int col = i0 * N + j0;
                        if (jn < WIDE_JCAP) jrn[jn] = (uint16_t)col;
                        st[col] = (unsigned char)k1c;
                    }
                }
            };
            if (commits) {
                if (a.dsum_e && dE != 0) stat_delta(a, grp, (long long)t + tid + 1, E + e_before, E + e_before + dE, 0);
                if (abits_row) atomicOr(abits_row + ((t + tid) >> 5), 1u << ((t + tid) & 31));
            }
            if (k_best >= 0) {
                // snapshot: the state at the first visit of the minimum (strict <, experiments.py:340), i.e. right after
                // commit k_best -- the elements moved since the previous snapshot go to the global copy
                commit_state(0, k_best + 1, jn0);
                cta_sync();
                const int jn = jn0 + k_best + 1;   // (jn0 > WIDE_JCAP: the journal had overflowed)
                auto put = [&](int el) {
                    if constexpr (FULL) {
                        int i, j, k;
                        unpack_pos(pos32, load_pos(st, pos32, el), i, j, k);
                        best_out[3 * el] = (uint8_t)i; best_out[3 * el + 1] = (uint8_t)j; best_out[3 * el + 2] = (uint8_t)k;
                    } else {
                        best_out[el] = st[el];
                    }
                };
                if (jn <= WIDE_JCAP) {
                    for (int e = tid; e < jn; e += NT) put(jrn[e]);
                } else {
                    for (int el = tid; el < a.Q; el += NT) put(el);
                }
                cta_sync();
                commit_state(k_best + 1, n_com, -(k_best + 1));   // the journal restarts after the snapshot
                if (tid == 0) *jcount = n_com - (k_best + 1);
                jfresh = false;
            } else if (n_com) {
                commit_state(0, n_com, jn0);
                if (tid == 0) *jcount = jn0 + n_com;   // WIDE_JCAP + 1 and beyond: overflow, the next snapshot is a full copy
                jfresh = false;
            }
            if (was_near && tid < adv) { ++near; flips += (uint32_t)was_flip; }
            if (tid < adv && a.hist_kind) {
                const int v = E + e_before + (accept ? dE : 0);
                if (a.hist_kind == 1) reinterpret_cast<uint16_t *>(hrow)[s + 1] = (uint16_t)v;
                else reinterpret_cast<int *>(hrow)[s + 1] = v;
            }
            cta_sync();   // counters and state are final before the next round reads them
            E += e_run;
            n_acc += n_com;
            t += adv;
            if (adv * 2 > width) width = min(NT, width * 2);
            else if (adv * 8 < width) width = max(32, width >> 1);
            continue;
        }
        // ---------------- the CTA commits its first accepted proposal ----------------
        // thread index in the high half, delta-E (biased) in the low half: the minimum over the CTA is the first
        // accepting thread together with its delta-E
        const int mine = accept ? (tid << 16) | (dE + 0x8000) : NONE;
        const int wmin = __reduce_min_sync(FULLMASK, mine);
        int cmin = wmin;
        if constexpr (NT > 32) {
            // each warp's first accepting thread publishes its move, so that after the barrier any warp can
            // take part in applying the winner's (counters and state are updated by two warps side by side)
            if (accept && mine == wmin) {
                xmv[3 * warp] = (uint32_t)(i0 | (j0 << 8) | (k0c << 16));
                xmv[3 * warp + 1] = (uint32_t)(i1 | (j1 << 8) | (k1c << 16));
                xmv[3 * warp + 2] = (uint32_t)qsel;
            }
            if (lane == 0) xch[warp] = wmin;
            __syncthreads();
            cmin = __reduce_min_sync(FULLMASK, lane < NW ? xch[lane] : NONE);
        }
        int first = cmin == NONE ? -1 : cmin >> 16;
        int adv = first >= 0 ? first + 1 : rem;            // steps consumed by this round
        int adv_h = adv;                                   // steps whose energy is appended to the history
        bool stop = false;
        int E_new = first >= 0 ? E + (cmin & 0xffff) - 0x8000 : E;
        bool improved = E_new < best;
        if constexpr (EARLY) {
            // experiments.py:343-353: the counter resets on a strict improvement only, and the
            // break happens before the history append of the stopping step
            const int e_stop = max(a.patience - stale - 1, 0);   // rejected step at which patience runs out
            if (first < 0 || first > e_stop) {
                if (e_stop < rem) { stop = true; first = -1; adv = e_stop + 1; adv_h = e_stop; stale += e_stop + 1; E_new = E; improved = false; }
                else stale += adv;
            } else {
                stale = improved ? 0 : stale + first + 1;
                if (stale >= a.patience) { stop = true; adv_h = first; }
            }
        }
        const bool has = first >= 0;
        if (was_near && tid < adv) { ++near; flips += (uint32_t)was_flip; }
        // history: steps t .. t+adv_h-1; all but an accepted last one keep the old energy
        if (tid < adv_h && a.hist_kind) {
            const int v = (tid == first) ? E_new : E;
            if (a.hist_kind == 1) reinterpret_cast<uint16_t *>(hrow)[s + 1] = (uint16_t)v;
            else reinterpret_cast<int *>(hrow)[s + 1] = v;
        }
        // ---------------- the move is applied: lane f of one warp owns family f, another warp the state ----------------
        const int wf = first >> 5;                              // the winner's warp
        const int ws = NT > 32 ? (wf + 1) & (NW - 1) : wf;      // the warp that updates state, occupancy and journal
        if (has && (warp == wf || warp == ws)) {
            uint32_t po, pn;
            int wq;
            if constexpr (NT > 32) {
                po = xmv[3 * wf]; pn = xmv[3 * wf + 1]; wq = (int)xmv[3 * wf + 2];
            } else {
                const int src = first & 31;
                po = __shfl_sync(FULLMASK, (uint32_t)(i0 | (j0 << 8) | (k0c << 16)), src);
                pn = __shfl_sync(FULLMASK, (uint32_t)(i1 | (j1 << 8) | (k1c << 16)), src);
                wq = __shfl_sync(FULLMASK, qsel, src);
            }
            const int a0 = po & 255, b0 = (po >> 8) & 255, c0 = po >> 16, a1 = pn & 255, b1 = (pn >> 8) & 255, c1 = pn >> 16;
            if (warp == wf && lane >= F0 && lane < NFAM) {
                const int4 cf = a.coef[lane], cs = a.csel[lane];
                const int o = line_index(cf, cs, a0, b0, c0), n = line_index(cf, cs, a1, b1, c1);
                if (o != n) {
                    const int vo = cnt[o], vn = cnt[n];
                    cnt[o] = (uint8_t)(vo - 1); cnt[n] = (uint8_t)(vn + 1);
                }
            }
            if (warp == ws && lane == 31) {
                const int jn = jfresh ? 0 : *jcount;
                if (jn < WIDE_JCAP) jrn[jn] = (uint16_t)(FULL ? wq : a0 * N + b0);
                *jcount = jn + 1;   // WIDE_JCAP + 1 and beyond: overflow, the next snapshot is a full copy
                if constexpr (FULL) {
                    const int cid0 = (a0 * N + b0) * N + c0, cid1 = (a1 * N + b1) * N + c1;
                    occ[cid0 >> 5] &= ~(1u << (cid0 & 31));
                    occ[cid1 >> 5] |= 1u << (cid1 & 31);
                    store_pos(st, pos32, wq, pack_pos(pos32, a1, b1, c1));
                } else {
                    st[a0 * N + b0] = (unsigned char)c1;
                }
            }
        }
        cta_sync();   // counters and state are final before the next round (and before a snapshot) reads them

        // ---------------- bookkeeping ----------------
        const int n_before = n_acc;
        if (a.dsum_e && tid == 0) {   // statistics in difference form (KArgs::dsum_e)
            if (has && E_new != E) stat_delta(a, grp, (long long)t + first + 1, E, E_new, 0);
            if (stop) stat_delta(a, grp, (long long)t + adv, has ? E_new : E, 0, -1);   // nothing is appended from here on
        }
        if (has) { E = E_new; ++n_acc; jfresh = false; }
        if (t + adv - 1 >= next_edge || improved || stop || (has && abits_row != nullptr)) {
            // acceptance bins: close every bin that ends at or before the last consumed step
            while (t + adv - 1 >= next_edge) {
                if (tid == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_before - bin_mark);
                bin_mark = n_before;
                ++bin;
                next_edge = a.bin_starts[bin + 1];
            }
            if (has) {
                const int ta = t + first;
                if (tid == 0 && abits_row) atomicOr(abits_row + (ta >> 5), 1u << (ta & 31));
                if (improved) {
                    // snapshot: the state at the first visit of the minimum (strict <, :252 / :340).  The copy lives in
                    // global memory (shared memory is what limits the chains per SM); only the elements moved since
                    // the previous snapshot are written -- a few bytes, off the critical path
                    best = E;
                    if (!stop) best_step = ta + 1;
                    const int jn = *jcount;
                    auto put = [&](int el) {
                        if constexpr (FULL) {
                            int i, j, k;
                            unpack_pos(pos32, load_pos(st, pos32, el), i, j, k);
                            best_out[3 * el] = (uint8_t)i; best_out[3 * el + 1] = (uint8_t)j; best_out[3 * el + 2] = (uint8_t)k;
                        } else {
                            best_out[el] = st[el];
                        }
                    };
                    if (jn <= WIDE_JCAP) {
                        for (int e = tid; e < jn; e += NT) put(jrn[e]);
                    } else {
                        for (int el = tid; el < a.Q; el += NT) put(el);
                    }
                    jfresh = true;
                }
            }
            if (stop) {
                done = t + adv - 1;
                if (tid == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
            }
        }
        t = stop ? a.t_end : t + adv;
        // widen when the round was (nearly) used up, narrow when most of it was discarded
        if (adv * 2 > width) width = min(NT, width * 2);
        else if (adv * 8 < width) width = max(32, width >> 1);
    }

    // ---------------- write the record back ----------------
    __syncthreads();
    if (near && a.near_cnt) atomicAdd(a.near_cnt + chain, near);
    if (flips && a.flip_cnt) atomicAdd(a.flip_cnt + chain, flips);
    if (tid == 0) {
        if (a.t_end == a.n_steps && a.n_bins > 0 && a.acc_hist && done == a.t_end)
            a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
        a.cur_e[chain] = E;
        a.best_e[chain] = best;
        a.best_step[chain] = best_step;
        a.n_acc[chain] = n_acc;
        a.stale[chain] = stale;
        a.bin_mark[chain] = bin_mark;
        a.steps_done[chain] = done;
    }
    {
        uint8_t *out = a.state + (size_t)chain * a.state_bytes;
        if constexpr (FULL) {
            for (int qi = tid; qi < a.Q; qi += NT) {
                int i, j, k;
                unpack_pos(pos32, load_pos(st, pos32, qi), i, j, k);
                out[3 * qi] = (uint8_t)i; out[3 * qi + 1] = (uint8_t)j; out[3 * qi + 2] = (uint8_t)k;
            }
        } else {
            for (int c = tid; c < a.Q; c += NT) out[c] = st[c];
        }
    }
}

}  // namespace mcq
