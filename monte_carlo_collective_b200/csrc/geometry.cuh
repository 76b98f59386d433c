// Geometry of the conflict-table kernels (spec.cuh, fast.cuh): slab layout, neighbour rows, shared-line table.
// constexpr: the host computes it per call, and kernels compiled for a fixed board size fold it into immediates.
#pragma once
#include "anneal.cuh"

namespace mcq {

#ifndef MCQ_SPEC_MINB
#define MCQ_SPEC_MINB 7   // min CTAs (of 4 warps) per SM the register allocation of spec_kernel must allow
#endif
// CTA shape of fast_kernel (one chain per warp): warps per CTA and CTAs per SM the registers must allow
#ifndef MCQ_FAST_WARPS
#define MCQ_FAST_WARPS 4
#endif
#ifndef MCQ_FAST_MINB
#define MCQ_FAST_MINB 7
#endif
constexpr int MAX_NBR_ROUNDS = 8;

// direction of the attack line of family f (same family order as make_coefs / line_ids);
// the first non-zero component is always +1
__host__ __device__ constexpr void family_dir(int f, int &dx, int &dy, int &dz) {
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

// Slab geometry of the conflict-table kernel.  constexpr: the host computes it per call, and kernels compiled
// for a fixed board size (template parameter CN) fold it into immediates.
__host__ __device__ constexpr int spec_round_up(int x, int m) { return (x + m - 1) / m * m; }
// Number of entries of a cell's neighbour row: the cell itself + every other cell on its attack lines.  The
// line through (x,y,z) along (dx,dy,dz) runs over t in [lo, hi]; each axis with d != 0 bounds t.
__host__ __device__ constexpr int spec_row_entries(int full, int N, int x, int y, int z) {
    int n = 1;
    for (int f = full ? 0 : 1; f < NFAM; ++f) {
        int d[3] = {0, 0, 0};
        family_dir(f, d[0], d[1], d[2]);
        const int c[3] = {x, y, z};
        int lo = -(N - 1), hi = N - 1;
        for (int a = 0; a < 3; ++a) {
            if (d[a] > 0) { lo = lo > -c[a] ? lo : -c[a]; hi = hi < N - 1 - c[a] ? hi : N - 1 - c[a]; }
            if (d[a] < 0) { lo = lo > c[a] - (N - 1) ? lo : c[a] - (N - 1); hi = hi < c[a] ? hi : c[a]; }
        }
        n += hi - lo;
    }
    return n;
}
// longest row of the board: the count is a sum of concave functions of the position, symmetric under the
// reflections of the cube, so the maximum sits on the cells next to the centre
__host__ __device__ constexpr int spec_max_row(int full, int N) {
    int best = 0;
    const int m = (N - 1) / 2;
    for (int x = m; x <= N / 2; ++x)
        for (int y = m; y <= N / 2; ++y)
            for (int z = m; z <= N / 2; ++z) {
                const int n = spec_row_entries(full, N, x, y, z);
                best = n > best ? n : best;
            }
    return best;
}
__host__ __device__ constexpr SLayout spec_layout(int full, int N, int Q) {
    SLayout L{};
    L.tbl = spec_round_up((N * N * N + 1) * (full ? 2 : 1), 4);   // full_3d: uint16 entries (count | occupied << 15)
    L.off_state = L.tbl;
    const int state_b = full ? Q * 4 : N * N;
    L.off_occ = spec_round_up(L.off_state + state_b, 4);          // (the occupancy flag lives in the table entries)
    L.off_rec = L.off_occ;
    L.off_ring = spec_round_up(L.off_rec + 8 * 4, 16);
    L.stride = L.off_ring + 64 * 16;
    L.nbr_len = spec_round_up(spec_max_row(full, N), 32);   // the cell itself + its line neighbours, longest row
    L.rounds = L.nbr_len / 32;
    const int W = 2 * N - 1;
    const int lut_bytes = (W * W * W + 31) / 32 * 4;
    L.wide_bias = (N - 1) * (W * W + W + 1);
    L.off_wide = full ? lut_bytes : 0;
    L.cta_bytes = full ? spec_round_up(lut_bytes + N * N * N * 2, 16) : 0;
    return L;
}
// ---- geometry tables, built once per (mode, N) and shared by every chain ----------------------
// nbr[cell][L]: slot 0 = `cell` itself, then the ids of all other cells on the attack lines through it,
// padded to a multiple of 32 with the id N^3, a scratch entry at the end of every T, so updates need
// no predicate.  (Lane 0 therefore always handles the moved queen's own cell in its first entry.)
static __global__ void build_neighbours_kernel(int full, int N, int L, int esize, uint16_t *nbr, uint16_t *tmp) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= N * N * N) return;
    const int x = cell / (N * N), y = (cell / N) % N, z = cell % N;
    uint16_t *row = nbr + (size_t)cell * L;
    uint16_t *list = tmp + (size_t)cell * L;   // unordered neighbours
    int n = 0;
    for (int f = full ? 0 : 1; f < NFAM; ++f) {
        int dx, dy, dz;
        family_dir(f, dx, dy, dz);
        for (int tau = -(N - 1); tau <= N - 1; ++tau) {
            if (tau == 0) continue;
            const int u = x + tau * dx, v = y + tau * dy, w = z + tau * dz;
            if ((unsigned)u < (unsigned)N && (unsigned)v < (unsigned)N && (unsigned)w < (unsigned)N)
                list[n++] = (uint16_t)((u * N + v) * N + w);
        }
    }
    // Order the row so that the 32 entries one warp instruction touches fall into distinct shared-memory
    // banks where possible (bank = (id * esize / 4) % 32; ids in the same 4-byte word do not conflict).
    // Greedy: each group of 32 takes at most one word per bank; leftovers fill the remaining slots.
    const uint16_t PAD = (uint16_t)(N * N * N), TAKEN = 0xffffu;
    int placed = 0;
    for (int g = 0; g < L / 32; ++g) {
        unsigned banks = 0u;
        int word_of_bank[32];
        int k = 0;
        if (g == 0) {
            const int word = (cell * esize) >> 2;
            banks = 1u << (word & 31); word_of_bank[word & 31] = word;
            row[k++] = (uint16_t)cell;
        }
        for (int e = 0; e < n && k < 32; ++e) {
            if (list[e] == TAKEN) continue;
            const int word = (list[e] * esize) >> 2, bank = word & 31;
            if ((banks >> bank) & 1u) { if (word_of_bank[bank] != word) continue; }
            else { banks |= 1u << bank; word_of_bank[bank] = word; }
            row[g * 32 + k++] = list[e];
            list[e] = TAKEN;
            ++placed;
        }
        // not enough conflict-free entries for this group: top up with whatever is left only in the last
        // groups (keeps early groups conflict-free); here simply leave the rest of the group as padding
        for (; k < 32; ++k) row[g * 32 + k] = PAD;
    }
    // anything still unplaced (more than L/32 entries in one bank): overwrite padding slots from the end
    for (int e = 0, slot = L - 1; e < n && placed < n; ++e) {
        if (list[e] == TAKEN) continue;
        while (slot >= 0 && row[slot] != PAD) --slot;
        row[slot] = list[e];
        list[e] = TAKEN;
        ++placed;
    }
}

// wide[c] = i*W^2 + j*W + k (W = 2N-1): the difference of two wide ids identifies (di,dj,dk);
// lut bit (di+o)*W^2 + (dj+o)*W + (dk+o) (o = N-1) is set iff two cells that far apart share an
// attack line: the non-zero |d| components are all equal (mcmc.py:149-167).
static __global__ void build_wide_lut_kernel(int N, uint16_t *wide, uint32_t *lut, int lut_words) {
    const int W = 2 * N - 1, o = N - 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N * N * N) {
        const int x = idx / (N * N), y = (idx / N) % N, z = idx % N;
        wide[idx] = (uint16_t)((x * W + y) * W + z);
    }
    if (idx < lut_words) {
        uint32_t bits = 0u;
        for (int b = 0; b < 32; ++b) {
            const int e = idx * 32 + b;
            if (e >= W * W * W) break;
            const int da = abs(e / (W * W) - o), db = abs((e / W) % W - o), dc = abs(e % W - o);
            const int m = max(da, max(db, dc));
            const bool sh = m > 0 && (da == 0 || da == m) && (db == 0 || db == m) && (dc == 0 || dc == m);
            bits |= (uint32_t)sh << b;
        }
        lut[idx] = bits;
    }
}

}  // namespace mcq
