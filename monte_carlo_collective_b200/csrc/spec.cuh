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
//   * delta-E is two byte loads: the slab holds T[cell] = sum over the 13 (12 in board mode)
//     attack lines through `cell` of the number of queens on that line, i.e. exactly what
//     conflicts_for_queen / conflicts_for_position (mcmc.py:185-226, mcmc_board.py:147-193) count,
//     plus 13 (12) for a queen standing on the cell itself:
//         conflicts(old) = T[old] - 13,   conflicts(new) = T[new] - [old and new share a line]
//     The table is maintained on accept only: -1 along the 13 lines through the old cell, +1 along
//     those through the new one (all 32 lanes cooperate, 13*N candidate cells per phase).
//   * per-chain shared memory is N^3 + state bytes (2.2 KB at N=12), so the register file, not
//     shared memory, bounds residency.
//
// Table entries are uint8: T <= 13*N (12*N in board mode), so the kernel serves N <= 19 in
// full_3d and N <= 21 in board mode; larger boards use anneal_kernel's line counters.
#pragma once
#include "anneal.cuh"

namespace mcq {

// direction of the attack line of family f (same family order as make_coefs / line_ids);
// the first non-zero component is always +1
__device__ __forceinline__ void family_dir(int f, int &dx, int &dy, int &dz) {
    // 2-bit fields (d+1), one base-4 digit per family, packed into immediates (no local array)
    // dx+1 per family F0..F12: 1,1,2,2,2,2,2,1,1,2,2,2,2
    // dy+1 per family F0..F12: 1,2,1,2,0,1,1,2,2,2,2,0,0
    // dz+1 per family F0..F12: 2,1,1,1,1,2,0,2,0,2,0,2,0
    constexpr unsigned long long PX = 1ull | 1ull << 2 | 2ull << 4 | 2ull << 6 | 2ull << 8 | 2ull << 10 | 2ull << 12 | 1ull << 14 |
                                      1ull << 16 | 2ull << 18 | 2ull << 20 | 2ull << 22 | 2ull << 24;
    constexpr unsigned long long PY = 1ull | 2ull << 2 | 1ull << 4 | 2ull << 6 | 0ull << 8 | 1ull << 10 | 1ull << 12 | 2ull << 14 |
                                      2ull << 16 | 2ull << 18 | 2ull << 20 | 0ull << 22 | 0ull << 24;
    constexpr unsigned long long PZ = 2ull | 1ull << 2 | 1ull << 4 | 1ull << 6 | 1ull << 8 | 2ull << 10 | 0ull << 12 | 2ull << 14 |
                                      0ull << 16 | 2ull << 18 | 0ull << 20 | 2ull << 22 | 0ull << 24;
    dx = (int)((PX >> (2 * f)) & 3) - 1;
    dy = (int)((PY >> (2 * f)) & 3) - 1;
    dz = (int)((PZ >> (2 * f)) & 3) - 1;
}

// Neighbour lists, shared by every chain of a launch (geometry only, L2-resident): row `cell` holds
// the ids of all cells on the attack lines through `cell` (itself excluded), padded to a multiple
// of 32 with the id N^3, a scratch byte at the end of every table, so updates need no predicate.
__global__ void build_neighbours_kernel(int full, int N, int L, uint16_t *nbr) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= N * N * N) return;
    const int x = cell / (N * N), y = (cell / N) % N, z = cell % N;
    uint16_t *row = nbr + (size_t)cell * L;
    int n = 0;
    for (int f = full ? 0 : 1; f < NFAM; ++f) {
        int dx, dy, dz;
        family_dir(f, dx, dy, dz);
        for (int tau = -(N - 1); tau <= N - 1; ++tau) {
            if (tau == 0) continue;
            const int u = x + tau * dx, v = y + tau * dy, w = z + tau * dz;
            if ((unsigned)u < (unsigned)N && (unsigned)v < (unsigned)N && (unsigned)w < (unsigned)N)
                row[n++] = (uint16_t)((u * N + v) * N + w);
        }
    }
    for (; n < L; ++n) row[n] = (uint16_t)(N * N * N);
}

constexpr int MAX_NBR_ROUNDS = 8;   // 13*(N-1) <= 256 for every N the uint8 table admits

// T[c] += delta for every cell c on the attack lines through one cell (its neighbour row).
// All lanes must call; `nr` = row length / 32.
__device__ __forceinline__ void table_lines_add(uint8_t *T, const uint16_t *row, int nr, int lane, int delta) {
    int c[MAX_NBR_ROUNDS];
#pragma unroll
    for (int r = 0; r < MAX_NBR_ROUNDS; ++r)
        if (r < nr) c[r] = __ldg(row + r * 32 + lane);
#pragma unroll
    for (int r = 0; r < MAX_NBR_ROUNDS; ++r)
        if (r < nr) T[c[r]] = (uint8_t)(T[c[r]] + delta);
}

template <bool FULL, bool REPLAY>
__global__ void __launch_bounds__(128, 8) spec_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr unsigned FULLMASK = 0xffffffffu;
    constexpr int NF = FULL ? NFAM : NFAM - 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int chain = blockIdx.x * (blockDim.x >> 5) + wid;
    const int N = a.N, rounds = a.sl.rounds;

    if (chain >= a.n_chains) return;
    const int L = a.sl.nbr_len;

    unsigned char *S = smem + (size_t)wid * a.sl.stride;
    uint8_t *T = S;
    unsigned char *st = S + a.sl.off_state;
    uint32_t *occ = reinterpret_cast<uint32_t *>(S + a.sl.off_occ);

    // ---- build the slab from the external state ----
    for (int w = lane; w < a.sl.stride / 4; w += 32) reinterpret_cast<uint32_t *>(S)[w] = 0u;
    __syncwarp();
    const uint8_t *ext = a.state + (size_t)chain * a.state_bytes;
    for (int qi = lane; qi < a.Q; qi += 32) {
        if (FULL) {
            const int i = ext[3 * qi], j = ext[3 * qi + 1], k = ext[3 * qi + 2];
            store_pos(st, 0, qi, pack_pos(0, i, j, k));
            const int cid = (i * N + j) * N + k;
            atomicOr(&occ[cid >> 5], 1u << (cid & 31));
        } else {
            st[qi] = ext[qi];
        }
    }
    __syncwarp();
    for (int qi = 0; qi < a.Q; ++qi) {
        int i, j, k;
        if (FULL) unpack_pos(0, load_pos(st, 0, qi), i, j, k);
        else { i = qi / N; j = qi - i * N; k = st[qi]; }
        const int c = (i * N + j) * N + k;
        table_lines_add(T, a.nbr + (size_t)c * L, rounds, lane, 1);
        if (lane == 0) T[c] = (uint8_t)(T[c] + NF);
        __syncwarp();
    }
    int E;
    {
        int e = 0;
        for (int qi = lane; qi < a.Q; qi += 32) {
            int i, j, k;
            if (FULL) unpack_pos(0, load_pos(st, 0, qi), i, j, k);
            else { i = qi / N; j = qi - i * N; k = st[qi]; }
            e += (int)T[(i * N + j) * N + k] - NF;
        }
        E = __reduce_add_sync(FULLMASK, e) >> 1;   // every attacking pair was counted from both ends
    }

    // ---- persistent record ----
    int best = E, best_step = 0, n_acc = 0, stale = 0, bin_mark = 0;
    int done = a.t_end;
    bool active = true;
    if (a.t_begin == 0) {
        if (lane == 0) {
            if (a.init_e) a.init_e[chain] = E;
            if (a.hist_kind == 1) reinterpret_cast<uint16_t *>(a.hist)[(size_t)chain * a.hist_pitch] = (uint16_t)E;
            else if (a.hist_kind == 2) reinterpret_cast<int *>(a.hist)[(size_t)chain * a.hist_pitch] = E;
        }
    } else {
        best = a.best_e[chain];
        best_step = a.best_step[chain];
        n_acc = a.n_acc[chain];
        stale = a.stale[chain];
        bin_mark = a.bin_mark[chain];
        const int sd = a.steps_done[chain];
        if (sd < a.t_begin) { active = false; done = sd; }
    }
    const unsigned long long sd64 = a.seeds ? a.seeds[chain] : 0ull;
    const uint32_t key0 = (uint32_t)sd64, key1 = (uint32_t)(sd64 >> 32);
    const int grp = a.group ? a.group[chain] : 0;
    const float *beta_row = REPLAY ? nullptr : a.beta_c + (size_t)grp * a.n_steps;
    const double *beta64_row = REPLAY ? a.beta64 + (size_t)grp * a.n_steps : nullptr;
    const uint32_t *mv_row = REPLAY ? a.rmoves + (size_t)chain * a.n_steps : nullptr;
    const double *un_row = REPLAY ? a.runif + (size_t)chain * a.n_steps : nullptr;

    int bin = a.bin_at_begin;
    int next_edge = a.n_bins > 0 ? a.bin_starts[bin + 1] : 0x7fffffff;
    int acc_blk = a.t_begin >> 5;
    uint32_t accbits = 0u, near = 0u;
    uint16_t *hist16 = a.hist_kind == 1 ? reinterpret_cast<uint16_t *>(a.hist) + (size_t)chain * a.hist_pitch : nullptr;
    int *hist32 = a.hist_kind == 2 ? reinterpret_cast<int *>(a.hist) + (size_t)chain * a.hist_pitch : nullptr;

    int t = a.t_begin;
    while (active && t < a.t_end) {
        const int s = t + lane;
        const bool valid = s < a.t_end;
        const int n_valid = min(32, a.t_end - t);

        // ---------------- this lane's proposal: step s against the current state ----------------
        int i0 = 0, j0 = 0, k0c = 0, i1 = 0, j1 = 0, k1c = 0, qsel = 0;
        bool bad = false, accept = false;
        uint32_t w_u = 0u;
        float cb = 0.f;
        double u64 = 0.0, b64 = 0.0;
        if (valid) {
            if constexpr (REPLAY) {
                const uint32_t mv = mv_row[s];
                u64 = un_row[s];
                b64 = beta64_row[s];
                if (FULL) {
                    qsel = mv & 0xfff; i1 = (mv >> 12) & 63; j1 = (mv >> 18) & 63; k1c = (mv >> 24) & 63;
                    bad = qsel >= a.Q || i1 >= N || j1 >= N || k1c >= N;
                    if (bad) { qsel = 0; i1 = j1 = k1c = 0; }
                    const int cid1 = (i1 * N + j1) * N + k1c;
                    bad = bad || ((occ[cid1 >> 5] >> (cid1 & 31)) & 1u);
                    unpack_pos(0, load_pos(st, 0, qsel), i0, j0, k0c);
                } else {
                    i0 = mv & 255; j0 = (mv >> 8) & 255; k1c = (mv >> 16) & 255;
                    bad = i0 >= N || j0 >= N || k1c >= N;
                    if (bad) { i0 = j0 = k1c = 0; }
                    k0c = st[i0 * N + j0];
                    bad = bad || (k1c == k0c);
                    i1 = i0; j1 = j0;
                }
            } else {
                cb = __ldg(beta_row + s);
                const Philox4 r = philox4x32_10((uint32_t)s, 0u, 0u, PHILOX_DOMAIN_STEP, key0, key1);
                w_u = r.z;
                if (FULL) {
                    qsel = (int)__umulhi(r.x, (uint32_t)a.Q);
                    uint32_t word = r.y;
                    int tries = 0;
                    while (true) {
                        i1 = draw_digit(word, N); j1 = draw_digit(word, N); k1c = draw_digit(word, N);
                        const int cid1 = (i1 * N + j1) * N + k1c;
                        if (!((occ[cid1 >> 5] >> (cid1 & 31)) & 1u)) break;
                        if (tries == 0) word = r.w;
                        else {
                            const int e = tries - 1;
                            const Philox4 r2 = philox4x32_10((uint32_t)s, 0u, 1u + (uint32_t)(e >> 2), PHILOX_DOMAIN_STEP, key0, key1);
                            const int sel = e & 3;
                            word = sel == 0 ? r2.x : sel == 1 ? r2.y : sel == 2 ? r2.z : r2.w;
                        }
                        ++tries;
                    }
                    unpack_pos(0, load_pos(st, 0, qsel), i0, j0, k0c);
                } else {
                    uint32_t word = r.x;
                    i0 = draw_digit(word, N); j0 = draw_digit(word, N);
                    k0c = st[i0 * N + j0];
                    k1c = k0c + 1 + (int)__umulhi(r.y, (uint32_t)(N - 1));
                    k1c -= (k1c >= N) ? N : 0;
                    i1 = i0; j1 = j0;
                }
            }
        }
        // ---------------- delta-E: two table reads ----------------
        int dE;
        {
            const int c0 = (i0 * N + j0) * N + k0c, c1 = (i1 * N + j1) * N + k1c;
            dE = (int)T[c1] - (int)T[c0] + NF;
            if (FULL) {
                const int da = abs(i1 - i0), db = abs(j1 - j0), dc = abs(k1c - k0c);
                const int m = max(da, max(db, dc));
                const bool shared = (da == 0 || da == m) && (db == 0 || db == m) && (dc == 0 || dc == m);
                dE -= shared ? 1 : 0;   // the moving queen itself sits on a line through the new cell
            }
        }
        // ---------------- Metropolis test ----------------
        bool near_flag = false;
        if constexpr (REPLAY) {
            const double p = exp(-b64 * (double)dE);
            accept = u64 < fmin(1.0, p);
            near_flag = valid && !bad && fabs(u64 - p) < 1e-6;
            if (bad) accept = false;
        } else {
            const float p = exp2f(cb * (float)dE);
            const uint32_t thr = __float2uint_rz(p * 4294967296.0f);
            accept = (dE <= 0) || (w_u < thr);
        }
        accept = accept && valid;

        // ---------------- commit the first accepted proposal ----------------
        const unsigned acc_mask = __ballot_sync(FULLMASK, accept);
        int first = acc_mask ? __ffs(acc_mask) - 1 : -1;
        int adv = first >= 0 ? first + 1 : n_valid;   // steps consumed by this round
        int adv_h = adv;                              // steps whose energy is appended to the history
        bool stop = false;
        int E_new = E;
        bool improved = false;
        int w_dE = 0;
        if (first >= 0) {
            w_dE = __shfl_sync(FULLMASK, dE, first);
            E_new = E + w_dE;
            improved = E_new < best;
        }
        if (!FULL && a.patience >= 0) {
            // experiments.py:343-353: the counter resets on a strict improvement only, and the
            // break happens before the history append of the stopping step
            const int e_stop = max(a.patience - stale - 1, 0);   // rejected step at which patience runs out
            if (first < 0 || first > e_stop) {
                if (e_stop < n_valid) { stop = true; first = -1; adv = e_stop + 1; adv_h = e_stop; stale += e_stop + 1; E_new = E; improved = false; }
                else stale += adv;
            } else {
                stale = improved ? 0 : stale + first + 1;
                if (stale >= a.patience) { stop = true; adv_h = first; }
            }
        }
        if constexpr (REPLAY) {
            const unsigned committed = adv >= 32 ? FULLMASK : ((1u << adv) - 1u);
            near += __popc(__ballot_sync(FULLMASK, near_flag) & committed);
            const unsigned badm = __ballot_sync(FULLMASK, bad && valid) & committed;
            if (badm && lane == 0) atomicAdd(a.replay_err, (unsigned)__popc(badm));
        }
        // history: steps t .. t+adv_h-1; all but an accepted last one keep the old energy
        if (lane < adv_h) {
            const int v = (lane == first) ? E_new : E;
            const long long h = (long long)s + 1 - a.h_origin;
            if (hist16) hist16[h] = (uint16_t)v;
            else if (hist32) hist32[h] = v;
        }
        // acceptance bins: close every bin that ends at or before the last consumed step
        const int t_last = t + adv - 1;
        while (t_last >= next_edge) {
            if (lane == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
            bin_mark = n_acc;
            ++bin;
            next_edge = a.bin_starts[bin + 1];
        }
        if (first >= 0) {
            const int wq = __shfl_sync(FULLMASK, qsel, first);
            const int wc0 = __shfl_sync(FULLMASK, i0 | (j0 << 8) | (k0c << 16), first);
            const int wc1 = __shfl_sync(FULLMASK, i1 | (j1 << 8) | (k1c << 16), first);
            const int ai = wc0 & 255, aj = (wc0 >> 8) & 255, ak = wc0 >> 16;
            const int bi = wc1 & 255, bj = (wc1 >> 8) & 255, bk = wc1 >> 16;
            const int ca = (ai * N + aj) * N + ak, cbn = (bi * N + bj) * N + bk;
            table_lines_add(T, a.nbr + (size_t)ca * L, rounds, lane, -1);
            if (lane == 0) T[ca] = (uint8_t)(T[ca] - NF);
            __syncwarp();
            table_lines_add(T, a.nbr + (size_t)cbn * L, rounds, lane, +1);
            if (lane == 0) {
                T[cbn] = (uint8_t)(T[cbn] + NF);
                if (FULL) {
                    occ[ca >> 5] &= ~(1u << (ca & 31));
                    occ[cbn >> 5] |= 1u << (cbn & 31);
                    store_pos(st, 0, wq, pack_pos(0, bi, bj, bk));
                } else {
                    st[ai * N + aj] = (unsigned char)bk;
                }
            }
            __syncwarp();
            E = E_new;
            ++n_acc;
            const int ta = t + first;
            if ((ta >> 5) != acc_blk) {
                if (accbits && lane == 0 && a.abits) a.abits[(size_t)chain * a.abits_pitch + acc_blk] = accbits;
                acc_blk = ta >> 5;
                accbits = 0u;
            }
            accbits |= 1u << (ta & 31);
            if (improved) {
                best = E;
                if (!stop) best_step = ta + 1;
                uint8_t *bs = a.best_state + (size_t)chain * a.state_bytes;
                if (FULL) {
                    for (int qi = lane; qi < a.Q; qi += 32) {
                        int i, j, k;
                        unpack_pos(0, load_pos(st, 0, qi), i, j, k);
                        bs[3 * qi] = (uint8_t)i; bs[3 * qi + 1] = (uint8_t)j; bs[3 * qi + 2] = (uint8_t)k;
                    }
                } else {
                    for (int c = lane; c < a.Q; c += 32) bs[c] = st[c];
                }
            }
        }
        if (stop) {
            active = false;
            done = t + adv - 1;
            if (lane == 0 && a.acc_hist) a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
        }
        t += adv;
    }

    // ---------------- write the record back ----------------
    if (accbits && lane == 0 && a.abits) a.abits[(size_t)chain * a.abits_pitch + acc_blk] = accbits;
    if (a.t_end == a.n_steps && a.n_bins > 0 && lane == 0 && a.acc_hist && done == a.t_end)
        a.acc_hist[(size_t)chain * a.n_bins + bin] = (uint32_t)(n_acc - bin_mark);
    uint8_t *out = a.state + (size_t)chain * a.state_bytes;
    if (FULL) {
        for (int qi = lane; qi < a.Q; qi += 32) {
            int i, j, k;
            unpack_pos(0, load_pos(st, 0, qi), i, j, k);
            out[3 * qi] = (uint8_t)i; out[3 * qi + 1] = (uint8_t)j; out[3 * qi + 2] = (uint8_t)k;
        }
    } else {
        for (int c = lane; c < a.Q; c += 32) out[c] = st[c];
    }
    if (lane == 0) {
        a.cur_e[chain] = E;
        a.best_e[chain] = best;
        a.best_step[chain] = best_step;
        a.n_acc[chain] = n_acc;
        a.stale[chain] = stale;
        a.bin_mark[chain] = bin_mark;
        a.steps_done[chain] = done;
        if (REPLAY && a.near_cnt) a.near_cnt[chain] += near;
    }
}

}  // namespace mcq
