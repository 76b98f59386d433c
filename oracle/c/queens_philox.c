/*
 * CPU statement of the PRODUCTION chain: the reference's Metropolis loop driven by the engine's
 * counter-based random stream instead of NumPy's MT19937 (TEST INFRASTRUCTURE).
 *
 * The chain logic is the reference's (experiments.py:218-258 full_3d, :308-355 board): conflict
 * counts are O(Q) scans with the attack clauses of mcmc.py:149-167 / mcmc_board.py:104-120
 * (qo_conflicts / qo_energy in queens_oracle.c, pinned to the reference by the golden fixtures).
 * Only the source of randomness differs, and it is specified here, independently of the kernels:
 *
 *   chain_words(i, j)    Philox4x32-10(counter = (i, seed_lo, seed_hi, j), key = (0x243F6A88, 0x85A308D3)): the
 *                        chain's 64-bit seed sits in the counter, the cipher key is a constant
 *   mulhi(a, n)          floor(a * n / 2^32): uniform on [0, n) up to n / 2^32
 *
 *   board   (experiments.py:311-319)  a step needs 64 random bits and takes them from the two-word generator:
 *               (x, z) = Philox2x32-10(counter = (s, seed_lo), key = 0x243F6A88 ^ seed_hi)
 *           m = mulhi(x, N^2 (N - 1)) is one uniform index over the (column, other height) pairs, split as
 *           column ij = mulhi(x, N^2) (i = ij / N, j = ij % N), offset d = mulhi(lo32(x * N^2), N - 1)
 *           (m = ij (N - 1) + d exactly); new height = (old + 1 + d) mod N: uniform over the N-1 OTHER
 *           heights, the distribution of the reference's redraw loop.
 *   full_3d (experiments.py:221-231)  words of step s: (x, y, z, w) = chain_words(s, 0); extra words: stream
 *           j >= 1, chain_words(s, j).  queen q = mulhi(x, Q); the new cell is the first EMPTY one
 *           (the queen's own cell counts as occupied) of the candidates
 *               mulhi(y, N^3), mulhi(w, N^3), mulhi(lo32(x * Q), N^3),
 *               then mulhi(word e & 3 of stream 1 + (e >> 2), N^3) for e = 0, 1, 2, ...
 *           with cell id c = (i * N + j) * N + k: uniform over the empty cells, the distribution
 *           of the reference's rejection loop against occ_set.
 *   accept  (experiments.py:238-239, :326-327)  delta <= 0, or U < exp(-beta_s * delta) in float64
 *           with the 53-bit uniform U = (z * 2^21 + (v >> 11)) / 2^53, z the step's uniform word above (board: the
 *           second Philox2x32 word; full_3d: word z), v = word x of chain_words(s, 0x80000000), and beta_s
 *           the float64 schedule value.  The uniform is a function
 *           of the step, so "drawn every step" (appendix A.1 of SURVEY.md) holds trivially.
 *
 *   initial states (mcmc_board.py:26-59, mcmc.py:20-101): words come from stream 0x40000000,
 *           word 4c + m = component m of chain_words(c, 0x40000000).
 *           board random: heights in row-major order, one word each, mulhi(word, N); klarner
 *           fallback cells likewise.  full_3d random / klarner fill: one word per attempt,
 *           (i, j, k) = its three leading base-N digits; an occupied cell is skipped.
 *
 * The CUDA kernels evaluate the accept test in float32 and fall back to exactly this float64 test
 * inside an error band, so their trajectories must equal this file's bit for bit
 * (tests/test_gpu_philox_parity.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg
 * may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

long qo_energy(int with_column, int q, const int32_t *cells);
int qo_conflicts(int with_column, int q, const int32_t *cells, int skip, int ti, int tj, int tk);

/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), written from the paper's round function:
 * (c0, c1, c2, c3) -> (hi(M1*c2) ^ c1 ^ k0, lo(M1*c2), hi(M0*c0) ^ c3 ^ k1, lo(M0*c0)), key += Weyl */
void qp_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

/* Philox2x32-10 (same paper): (L, R) -> (hi(M*L) ^ key ^ R, lo(M*L)), key += Weyl */
void qp_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
    uint32_t l = ctr[0], r = ctr[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p = (uint64_t)0xD256D193u * l;
        l = (uint32_t)(p >> 32) ^ key ^ r;
        r = (uint32_t)p;
        key += 0x9E3779B9u;
    }
    out[0] = l; out[1] = r;
}

static inline uint32_t mulhi(uint32_t a, uint32_t n) { return (uint32_t)(((uint64_t)a * n) >> 32); }

/* the two words of a board step */
static void board_step_words(uint64_t seed, uint32_t step, uint32_t out[2]) {
    const uint32_t ctr[2] = {step, (uint32_t)seed};
    qp_philox2x32_10(ctr, 0x243F6A88u ^ (uint32_t)(seed >> 32), out);
}

/* the chain's seed sits in the counter; the cipher key is a constant (first 64 fractional bits of pi) */
static void chain_words(uint64_t seed, uint32_t index, uint32_t stream, uint32_t out[4]) {
    const uint32_t ctr[4] = {index, (uint32_t)seed, (uint32_t)(seed >> 32), stream};
    const uint32_t key[2] = {0x243F6A88u, 0x85A308D3u};
    qp_philox4x32_10(ctr, key, out);
}

static int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

static int occupied(int q, const int32_t *cells, int i, int j, int k) {
    for (int a = 0; a < q; ++a)
        if (cells[3 * a] == i && cells[3 * a + 1] == j && cells[3 * a + 2] == k) return 1;
    return 0;
}

/* Initial state of the production path.  mode 0 board: cells[i*n+j] = (i, j, height); mode 1
 * full_3d: q rows (i, j, k).  init_mode 0 random, 1 latin, 2 klarner.  Returns 0, or -1 for bad arguments. */
int qp_init_state(int mode, int n, int q, int init_mode, uint64_t seed, int32_t *cells) {
    uint32_t buf[4];
    uint32_t ctr = 0;
    int have = 0;
#define NEXT_WORD() (have == 0 ? (chain_words(seed, ctr++, 0x40000000u, buf), have = 3, buf[0]) : buf[4 - have--])
    int m = n;   /* Klarner core edge: n itself when gcd(n, 210) == 1, else the largest coprime M < n */
    if (init_mode == 2 && gcd_int(n, 210) != 1)
        for (m = n - 1; m > 0; --m) if (gcd_int(m, 210) == 1) break;
    if (init_mode < 0 || init_mode > 2) return -1;
    if (mode == 0) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                int k;
                if (init_mode == 1) k = (i + j) % n;                              /* mcmc_board.py:31 */
                else if (init_mode == 2 && i < m && j < m) k = (3 * i + 5 * j) % m; /* :34-36, :50-52 */
                else k = (int)mulhi(NEXT_WORD(), (uint32_t)n);                    /* :28, :54-57 */
                int32_t *c = cells + 3 * (i * n + j);
                c[0] = i; c[1] = j; c[2] = k;
            }
        return 0;
    }
    int placed = 0;
    if (init_mode != 0) {
        if (q != n * n) return -1;                                                /* mcmc.py:22-26 */
        const int edge = init_mode == 1 ? n : m;
        for (int i = 0; i < edge; ++i)
            for (int j = 0; j < edge; ++j) {
                cells[3 * placed] = i; cells[3 * placed + 1] = j;
                cells[3 * placed + 2] = init_mode == 1 ? (i + j) % n : (3 * i + 5 * j) % m;
                ++placed;
            }
    }
    while (placed < q) {   /* distinct uniformly random cells (mcmc.py:63-90, :97-101) */
        uint32_t w = NEXT_WORD();
        const int i = (int)mulhi(w, (uint32_t)n); w *= (uint32_t)n;
        const int j = (int)mulhi(w, (uint32_t)n); w *= (uint32_t)n;
        const int k = (int)mulhi(w, (uint32_t)n);
        if (occupied(placed, cells, i, j, k)) continue;
        cells[3 * placed] = i; cells[3 * placed + 1] = j; cells[3 * placed + 2] = k;
        ++placed;
    }
#undef NEXT_WORD
    return 0;
}

/*
 * One production chain.  cells is the initial state and is updated in place (final state).
 * history has n_steps+1 slots; moves (4 ints per step: board (i, j, new_k, 0), full_3d (q, i, j, k)) and
 * uniforms (the 53-bit U of every step) may be NULL; when given, feeding them to qo_replay must reproduce
 * this function's outputs.  Returns steps done (== n_steps unless the board patience stopped the chain).
 * n_near counts decisions with |U - exp(-beta*delta)| < 1e-6 (the replay path's definition).
 */
long qp_chain(int mode, int n, int q, int32_t *cells, long n_steps, uint64_t seed, const double *betas, long patience,
              int32_t *history, uint8_t *accepted, int32_t *best_cells, int32_t *best_energy, long *steps_to_best,
              long *n_near, int32_t *final_energy, int32_t *moves, double *uniforms) {
    const int with_column = mode == 1;
    const uint32_t n3 = (uint32_t)n * n * n;
    long cur = qo_energy(with_column, q, cells), best = cur, best_at = 0, stale = 0, near = 0, done = n_steps;
    memcpy(best_cells, cells, sizeof(int32_t) * 3 * (size_t)q);
    history[0] = (int32_t)cur;
    for (long s = 0; s < n_steps; ++s) {
        uint32_t r[4], z;
        int idx, ti, tj, tk;
        if (mode == 0) {
            uint32_t b[2];
            board_step_words(seed, (uint32_t)s, b);
            idx = (int)mulhi(b[0], (uint32_t)(n * n));
            ti = idx / n; tj = idx % n;
            tk = cells[3 * idx + 2] + 1 + (int)mulhi(b[0] * (uint32_t)(n * n), (uint32_t)(n - 1));
            if (tk >= n) tk -= n;
            z = b[1];
        } else {
            chain_words(seed, (uint32_t)s, 0u, r);
            z = r[2];
            idx = (int)mulhi(r[0], (uint32_t)q);
            uint32_t cand = mulhi(r[1], n3);
            for (int attempt = 0;; ++attempt) {
                ti = (int)(cand / ((uint32_t)n * n)); tj = (int)(cand / (uint32_t)n % (uint32_t)n); tk = (int)(cand % (uint32_t)n);
                if (!occupied(q, cells, ti, tj, tk)) break;
                if (attempt == 0) cand = mulhi(r[3], n3);
                else if (attempt == 1) cand = mulhi(r[0] * (uint32_t)q, n3);
                else {
                    uint32_t e[4];
                    const int extra = attempt - 2;
                    chain_words(seed, (uint32_t)s, 1u + (uint32_t)(extra >> 2), e);
                    cand = mulhi(e[extra & 3], n3);
                }
            }
        }
        const int before = qo_conflicts(with_column, q, cells, idx, cells[3 * idx], cells[3 * idx + 1], cells[3 * idx + 2]);
        const int after = qo_conflicts(with_column, q, cells, idx, ti, tj, tk);
        const int delta = after - before;
        uint32_t lo[4];
        chain_words(seed, (uint32_t)s, 0x80000000u, lo);
        const double u = (double)(((uint64_t)z << 21) | (uint64_t)(lo[0] >> 11)) * (1.0 / 9007199254740992.0);
        const double p = exp(-betas[s] * (double)delta);
        const int acc = delta <= 0 || u < p;
        if (fabs(u - p) < 1e-6) ++near;
        accepted[s] = (uint8_t)acc;
        if (moves) {
            int32_t *mv = moves + 4 * s;
            if (mode == 0) { mv[0] = ti; mv[1] = tj; mv[2] = tk; mv[3] = 0; }
            else { mv[0] = idx; mv[1] = ti; mv[2] = tj; mv[3] = tk; }
        }
        if (uniforms) uniforms[s] = u;
        int improved = 0;
        if (acc) {
            cells[3 * idx] = ti; cells[3 * idx + 1] = tj; cells[3 * idx + 2] = tk;
            cur += delta;
            if (cur < best) { best = cur; improved = 1; memcpy(best_cells, cells, sizeof(int32_t) * 3 * (size_t)q); }
        }
        if (mode == 0 && patience >= 0) {                              /* experiments.py:343-353 */
            stale = improved ? 0 : stale + 1;
            if (stale >= patience) { done = s; break; }
        }
        if (improved) best_at = s + 1;
        history[s + 1] = (int32_t)cur;
    }
    *best_energy = (int32_t)best;
    *steps_to_best = best_at;
    *n_near = near;
    *final_energy = (int32_t)cur;
    return done;
}
