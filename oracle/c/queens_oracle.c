/*
 * C restatement of the reference's chain for one recorded proposal stream (TEST INFRASTRUCTURE).
 *
 * Same algorithm as the reference, not the kernels' data structures: the conflict count of a cell
 * is an O(Q) scan over all other queens with the seven (six) attack clauses of
 * mcmc.py:149-167 / mcmc_board.py:104-120, and the chain loop is experiments.py:218-258 (full_3d)
 * / :308-355 (board) with the draws replaced by a given stream (move + float64 uniform per step).
 * Pinned to the reference by tests/test_oracle_golden.py (golden replay fixtures).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int iabs(int v) { return v < 0 ? -v : v; }

/* does a queen at a attack cell b?  (a != b assumed by callers) */
static inline int attacks(int with_column, const int32_t *a, int bi, int bj, int bk) {
    const int zi = a[0] == bi, zj = a[1] == bj, zk = a[2] == bk;
    const int di = iabs(a[0] - bi), dj = iabs(a[1] - bj), dk = iabs(a[2] - bk);
    if ((zi && zk) || (zj && zk)) return 1;            /* same (i,k), same (j,k) */
    if (zk && di == dj) return 1;                      /* planar diagonal, k-plane */
    if (zj && di == dk) return 1;                      /* planar diagonal, j-plane */
    if (zi && dj == dk) return 1;                      /* planar diagonal, i-plane */
    if (di == dj && dj == dk) return 1;                /* space diagonal */
    if (with_column && zi && zj) return 1;             /* same (i,j) column */
    return 0;
}

/* number of unordered attacking pairs (mcmc.py:134-169; board: mcmc_board.py:82-122) */
long qo_energy(int with_column, int q, const int32_t *cells) {
    long e = 0;
    for (int a = 0; a < q; ++a)
        for (int b = a + 1; b < q; ++b)
            e += attacks(with_column, cells + 3 * a, cells[3 * b], cells[3 * b + 1], cells[3 * b + 2]);
    return e;
}

/* queens other than `skip` attacking (ti,tj,tk): mcmc.py:185-226 / mcmc_board.py:147-193 */
int qo_conflicts(int with_column, int q, const int32_t *cells, int skip, int ti, int tj, int tk) {
    int c = 0;
    for (int a = 0; a < q; ++a)
        if (a != skip) c += attacks(with_column, cells + 3 * a, ti, tj, tk);
    return c;
}

/*
 * Replay one chain.  mode 0 = board (cells[i*n+j] = (i, j, height); move = (i, j, new_k, -)),
 * mode 1 = full_3d (move = (queen, i, j, k)).  cells is updated in place (final state).
 * history has n_steps+1 slots; returns the number of history entries written minus one
 * (== n_steps unless the board patience stopped the chain), or -1 on an illegal move.
 */
long qo_replay(int mode, int n, int q, int32_t *cells, long n_steps, const int32_t *moves, const double *uniforms,
               const double *betas, long patience, int32_t *history, uint8_t *accepted, int32_t *best_cells,
               int32_t *best_energy, long *steps_to_best, long *n_near, int32_t *final_energy) {
    const int with_column = mode == 1;
    long cur = qo_energy(with_column, q, cells), best = cur, best_at = 0, stale = 0, near = 0, done = n_steps;
    memcpy(best_cells, cells, sizeof(int32_t) * 3 * (size_t)q);
    history[0] = (int32_t)cur;
    for (long s = 0; s < n_steps; ++s) {
        const int32_t *mv = moves + 4 * s;
        int idx, ti, tj, tk;
        if (mode == 0) { idx = mv[0] * n + mv[1]; ti = mv[0]; tj = mv[1]; tk = mv[2]; if (tk == cells[3 * idx + 2]) return -1; }
        else {
            idx = mv[0]; ti = mv[1]; tj = mv[2]; tk = mv[3];
            for (int a = 0; a < q; ++a)
                if (cells[3 * a] == ti && cells[3 * a + 1] == tj && cells[3 * a + 2] == tk) return -1;
        }
        const int before = qo_conflicts(with_column, q, cells, idx, cells[3 * idx], cells[3 * idx + 1], cells[3 * idx + 2]);
        const int after = qo_conflicts(with_column, q, cells, idx, ti, tj, tk);
        const int delta = after - before;
        const double p = exp(-betas[s] * (double)delta);
        const int acc = uniforms[s] < (p < 1.0 ? p : 1.0);          /* experiments.py:238-239 / :326-327 */
        if (fabs(uniforms[s] - p) < 1e-6) ++near;
        accepted[s] = (uint8_t)acc;
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

/* ---- self-driving variant: same loop, proposals drawn here and RECORDED, so that a replay of the
 * recorded stream elsewhere (the CUDA kernels) must reproduce every output.  The generator is
 * splitmix64; proposal distributions follow experiments.py:221-231 / :311-319 (rejection loops). */
static inline uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline int draw_below(uint64_t *x, int n) {       /* unbiased: rejection on the top bits */
    const uint64_t lim = UINT64_MAX - UINT64_MAX % (uint64_t)n;
    uint64_t v;
    do v = splitmix(x); while (v >= lim);
    return (int)(v % (uint64_t)n);
}

long qo_generate(int mode, int n, int q, int32_t *cells, long n_steps, const double *betas, uint64_t seed, long patience,
                 int32_t *moves, double *uniforms, int32_t *history, uint8_t *accepted, int32_t *best_cells,
                 int32_t *best_energy, long *steps_to_best, long *n_near, int32_t *final_energy) {
    uint64_t x = seed;
    /* draw the stream step by step against an evolving copy of the state */
    int32_t *work = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)q);
    memcpy(work, cells, sizeof(int32_t) * 3 * (size_t)q);
    const int with_column = mode == 1;
    for (long s = 0; s < n_steps; ++s) {
        int32_t *mv = moves + 4 * s;
        int idx, ti, tj, tk;
        if (mode == 0) {
            ti = draw_below(&x, n); tj = draw_below(&x, n); idx = ti * n + tj;
            do tk = draw_below(&x, n); while (tk == work[3 * idx + 2]);
            mv[0] = ti; mv[1] = tj; mv[2] = tk; mv[3] = 0;
        } else {
            idx = draw_below(&x, q);
            for (;;) {
                ti = draw_below(&x, n); tj = draw_below(&x, n); tk = draw_below(&x, n);
                int occ = 0;
                for (int a = 0; a < q && !occ; ++a) occ = work[3 * a] == ti && work[3 * a + 1] == tj && work[3 * a + 2] == tk;
                if (!occ) break;
            }
            mv[0] = idx; mv[1] = ti; mv[2] = tj; mv[3] = tk;
        }
        uniforms[s] = (double)(splitmix(&x) >> 11) * (1.0 / 9007199254740992.0);
        const int before = qo_conflicts(with_column, q, work, idx, work[3 * idx], work[3 * idx + 1], work[3 * idx + 2]);
        const int after = qo_conflicts(with_column, q, work, idx, ti, tj, tk);
        const double p = exp(-betas[s] * (double)(after - before));
        if (uniforms[s] < (p < 1.0 ? p : 1.0)) { work[3 * idx] = ti; work[3 * idx + 1] = tj; work[3 * idx + 2] = tk; }
    }
    free(work);
    return qo_replay(mode, n, q, cells, n_steps, moves, uniforms, betas, patience, history, accepted, best_cells,
                     best_energy, steps_to_best, n_near, final_energy);
}
