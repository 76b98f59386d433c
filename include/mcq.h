/*
 * mcq.h -- C ABI of libmcq, the B200 (sm_100a) annealing engine for the 3D N^2-queens
 * MCMC hot path of galgantar/monte-carlo-collective.
 *
 * The reference has no FFI: its boundary for this path is a set of Python functions in
 * experiments.py.  Each entry point below replaces the *implementation* behind one of
 * them; the Python shim (monte_carlo_collective_b200/api.py) keeps the reference's
 * signatures and binds these symbols with ctypes (INTEGRATION.md shows the stub):
 *
 *   mcq_energy          <- State3DQueens._compute_energy        mcmc.py:134-169
 *                          State3DQueensBoard._compute_energy   mcmc_board.py:82-122
 *   mcq_delta_energy    <- conflicts_for_queen(new) - (old)     mcmc.py:185-226
 *                          conflicts_for_position(new) - (old)  mcmc_board.py:147-193
 *   mcq_run             <- metropolis_mcmc / metropolis_mcmc_board for a whole batch of
 *                          chains, i.e. the body of run_experiment
 *                          experiments.py:199-279, :282-376, :475-573
 *   mcq_philox4x32_10   -- the counter-based generator the chains draw from (host copy,
 *                          for known-answer tests; the reference uses NumPy's MT19937)
 *   mcq_philox4x32_10_device -- the same generator as compiled for the GPU (known-answer tests
 *                          of the device code itself)
 *   mcq_philox2x32_10, mcq_philox2x32_10_device -- the two-word generator of the same family that
 *                          board steps draw from (a board step needs 64 random bits)
 *   mcq_beta_table      <- constant_beta / linear / exponential / logarithmic / sinusoidal
 *                          annealing schedules, experiments.py:13-77, evaluated on the device
 *
 * Conventions: plain C types only; the caller owns every buffer (host or device, see
 * `mem`); nothing returned is owned by the library except the opaque context.  Every
 * function returns 0 on success and a negative MCQ_E* code otherwise, with a message
 * available from mcq_last_error() (thread-local).  Calls on one context must be
 * serialised by the caller.  All functions are synchronous: they return after the work
 * on `stream` has completed.
 */
#ifndef MCQ_H_
#define MCQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCQ_ABI_VERSION 4
#define MCQ_RECORD_INTS 8

/* state space (experiments.py:497-502: "board" or anything else => full_3d) */
#define MCQ_MODE_BOARD 0  /* one queen per (i,j) column; state = uint8 heights[N*N], row-major (i,j) */
#define MCQ_MODE_FULL3D 1 /* Q distinct cells; state = uint8 cells[Q][3] = (i,j,k) */

/* initial state (mcmc_board.py:26-59, mcmc.py:20-101) */
#define MCQ_INIT_RANDOM 0
#define MCQ_INIT_LATIN 1
#define MCQ_INIT_KLARNER 2
#define MCQ_INIT_EXPLICIT 3 /* take init_states as given */

/* where the caller's buffers live */
#define MCQ_MEM_HOST 0   /* host pointers; the library stages H2D/D2H itself */
#define MCQ_MEM_DEVICE 1 /* device pointers on `device`; no copies */

/* energy history element type */
#define MCQ_HIST_NONE 0
#define MCQ_HIST_U16 1 /* legal when 13*Q*(N-1)/2 < 65536 */
#define MCQ_HIST_I32 2

/* delta-E data structure / kernel */
#define MCQ_ALGO_AUTO 0   /* conflict table up to N = 18 (full_3d) / 21 (board), else one CTA per chain on line counters;
                             replays of large boards use LINES / GMEM */
#define MCQ_ALGO_LINES 1  /* per-line occupancy counters, `lanes_per_chain` lanes per chain (anneal.cuh) */
#define MCQ_ALGO_TABLE 2  /* per-cell conflict table, one warp per chain, speculative rounds (spec.cuh) */
#define MCQ_ALGO_GMEM 3   /* line counters in global memory, one thread per chain: boards too large for shared memory */
#define MCQ_ALGO_WIDE 4   /* line counters in shared memory, one CTA per chain, speculative rounds of 256 steps (wide.cuh) */

/* inverse-temperature schedules (experiments.py:13-77), evaluated on the device in float64 */
#define MCQ_SCHED_CONSTANT 0    /* beta_const */
#define MCQ_SCHED_LINEAR 1      /* b0 + (s/(n-1)) (b1-b0); b1 if n <= 1 */
#define MCQ_SCHED_EXPONENTIAL 2 /* b0 exp(ln(b1/b0) s/(n-1)) */
#define MCQ_SCHED_LOGARITHMIC 3 /* b0 + (b1-b0) ln(1+s)/ln(1+n) */
#define MCQ_SCHED_SINUSOIDAL 4  /* b0 + (b1-b0)(1-cos(pi s/n))/2 */

typedef struct mcq_schedule {
    int32_t type; /* MCQ_SCHED_* */
    int32_t reserved;
    double beta_const;
    double beta_start;
    double beta_end;
} mcq_schedule;

/* error codes */
#define MCQ_OK 0
#define MCQ_EINVAL -1   /* bad argument (the Python shim raises ValueError) */
#define MCQ_ECUDA -2    /* CUDA runtime error */
#define MCQ_ENOMEM -3   /* shared-memory or device-memory budget exceeded */
#define MCQ_EREPLAY -4  /* a replayed proposal was illegal (occupied cell / same height) */

typedef struct mcq_ctx mcq_ctx;

/*
 * One batch of independent annealing chains.  Chain c runs the Metropolis loop of
 * experiments.py:218-258 (full_3d) or :308-355 (board) for n_steps proposals with inverse
 * temperature beta_table[chain_group[c]][step].
 */
typedef struct mcq_run_params {
    uint32_t struct_size; /* sizeof(mcq_run_params): ABI check */
    int32_t mode;         /* MCQ_MODE_* */
    int32_t n;            /* N, 2..64 */
    int32_t q;            /* number of queens; must be N*N in board mode */
    int32_t n_steps;      /* proposals per chain */
    int32_t n_chains;
    int32_t n_groups;     /* number of beta schedules */
    int32_t init_mode;    /* MCQ_INIT_* */
    int32_t mem;          /* MCQ_MEM_* for every pointer below */
    int32_t early_stop_patience; /* board only (experiments.py:349-353); < 0 = none */

    /* ---- inputs ---- */
    const uint64_t *chain_seeds; /* [n_chains] Philox seed; depends on the chain only, never on placement */
    const int32_t *chain_group;  /* [n_chains] in [0,n_groups), or NULL = all 0 */
    const mcq_schedule *schedules; /* [n_groups] schedule parameters: beta_t is evaluated on the device (production path;
                                      HOST memory always), or NULL when beta_f64 carries a tabulated schedule */
    const uint8_t *init_states;  /* [n_chains][state_bytes] when init_mode == EXPLICIT, else NULL */

    /* beta_f64 alone: production run on a schedule tabulated by the caller (an arbitrary closure of step).
       ---- replay of a recorded proposal / uniform stream: beta_f64, replay_moves and replay_uniforms ---- */
    const double *beta_f64;         /* [n_groups][n_steps] exact float64 betas */
    const uint32_t *replay_moves;   /* [n_chains][n_steps]; board: i | j<<8 | k'<<16;
                                       full_3d: q | i<<12 | j<<18 | k<<24 */
    const double *replay_uniforms;  /* [n_chains][n_steps] */

    /* ---- per-step outputs (optional) ---- */
    int32_t hist_dtype;    /* MCQ_HIST_*; NONE => energy_history ignored */
    int64_t hist_pitch;    /* elements per chain row, >= n_steps+1 */
    void *energy_history;  /* [n_chains][hist_pitch]; index 0 = initial energy, s+1 = after step s */
    uint32_t *accept_bits; /* [n_chains][ceil(n_steps/32)] bit s = step s accepted, or NULL */
    /* cross-replica statistics per group (what plot_energy_histories consumes,
       experiments.py:591-595): sums over the chains of the group of E and E^2 */
    int64_t *stat_sum_e;   /* [n_groups][n_steps+1] or NULL */
    int64_t *stat_sum_e2;  /* [n_groups][n_steps+1] or NULL (both or none) */
    int32_t *stat_count;   /* [n_groups][n_steps+1] or NULL: chains of the group that have an energy at that index
                              (all of them unless the board patience stopped some: the denominator of the mean) */
    /* accepted moves per step bin (plot_acceptance_rates_binned, experiments.py:660-686) */
    int32_t n_bins;            /* 0 = none */
    const int32_t *bin_starts; /* [n_bins+1] first step of each bin; bin_starts[n_bins] = n_steps (HOST memory always) */
    uint32_t *accept_hist;     /* [n_chains][n_bins] */

    /* ---- per-chain outputs (each optional) ---- */
    int32_t *initial_energy; /* [n_chains] */
    int32_t *final_energy;   /* [n_chains] */
    int32_t *best_energy;    /* [n_chains] */
    int32_t *steps_to_best;  /* [n_chains] first history index of the minimum */
    int32_t *n_accepted;     /* [n_chains] */
    int32_t *steps_done;     /* [n_chains] history length - 1 (== n_steps unless early-stopped) */
    uint8_t *final_state;    /* [n_chains][state_bytes] */
    uint8_t *best_state;     /* [n_chains][state_bytes] state at the first visit of best_energy */
    uint32_t *n_near_threshold; /* [n_chains] replay: accept decisions with |u - exp(-beta dE)| < 1e-6; production:
                                   decisions inside the float32 error band, taken with the float64 rule instead */
    uint32_t *n_fp32_flips;     /* [n_chains] production: band decisions float32 alone would have got wrong */

    /* ---- measurements ---- */
    float *kernel_ms;       /* HOST pointer: sum of annealing-kernel durations (CUDA events) */
    int32_t *gpu_launches;  /* HOST pointer: kernels launched by this call */

    /* ---- tuning; 0 = automatic ---- */
    int32_t lanes_per_chain; /* 4, 8, 16 or 32 */
    int32_t warps_per_cta;
    int32_t chunk_steps;     /* steps per launch when the history is streamed */
    int32_t max_chains_per_sm;
    int32_t algo;            /* MCQ_ALGO_*; a non-zero lanes_per_chain with AUTO selects LINES */
    void *stream;            /* cudaStream_t, or NULL for the context's own stream */
    int32_t accept_all_f64;  /* non-zero: every uphill accept decision is taken with the float64 rule (the band is
                                infinite); same chains, slower -- a test of the float32 fast path */

    /* ---- checkpoint / resume (the reference has none: SURVEY 5.4; chains are resumable here because the random
     *      stream is counter-based -- step s of a chain depends on (seed, s) only) ----
     * A call executes steps [start_step, stop_step) of the n_steps-long schedule.  Both are multiples of 32 (or
     * n_steps); 0 / 0 = the whole run.  A segment with start_step > 0 takes the chains' states at start_step
     * as `init_states` (MCQ_INIT_EXPLICIT), their records and best states from the previous segment, and
     * continues into the SAME output arrays: history columns, accept bits, closed acceptance bins and
     * statistics of earlier segments are preserved (host arrays are uploaded first). */
    int32_t start_step;
    int32_t stop_step;
    const int32_t *resume_record;     /* [MCQ_RECORD_INTS][n_chains]: record_out of the previous segment */
    const uint8_t *resume_best_state; /* [n_chains][state bytes] */
    int32_t *record_out;              /* [MCQ_RECORD_INTS][n_chains] (optional): initial, current and best energy, step of the
                                         best, accepted moves, steps done, steps since the last improvement, accepted
                                         moves at the start of the open acceptance bin */
} mcq_run_params;

/* lifetime ------------------------------------------------------------------------- */
int mcq_abi_version(void);
int mcq_sizeof_run_params(void); /* sizeof(mcq_run_params) as compiled: lets a binding verify its mirror */
const char *mcq_last_error(void);
int mcq_device_count(int *count);
int mcq_create(int device, mcq_ctx **out);
int mcq_destroy(mcq_ctx *ctx);
int mcq_device_info(mcq_ctx *ctx, int *sm_count, int *smem_per_sm, int *smem_per_block_optin,
                    int *clock_khz, char *name, int name_len);

/* bytes of one chain's state in the external format above */
int mcq_state_bytes(int mode, int n, int q);
/* shared-memory bytes one resident chain needs for (mode, n, q, lanes_per_chain) */
int mcq_chain_smem_bytes(int mode, int n, int q, int lanes_per_chain);

/* full-board energies: out_energy[b] = number of attacking pairs of states[b] */
int mcq_energy(mcq_ctx *ctx, int mode, int n, int q, int n_states, const uint8_t *states,
               int32_t *out_energy, int mem, void *stream);

/*
 * Delta energies of candidate moves evaluated with the kernel's line-occupancy counters,
 * without applying them.  moves use the replay packing; out_delta[b][m] = conflicts(new) -
 * conflicts(old) exactly as experiments.py:235 / :323 compute it.  A move whose queen index or
 * coordinates are out of range yields INT32_MIN.  Malformed states (height or coordinate out of
 * range, two queens on one cell) make mcq_energy / mcq_delta_energy / mcq_run return MCQ_EINVAL,
 * as the reference's state constructors raise ValueError (mcmc.py:113-118, mcmc_board.py:62-65).
 */
int mcq_delta_energy(mcq_ctx *ctx, int mode, int n, int q, int n_states, const uint8_t *states,
                     int n_moves, const uint32_t *moves, int32_t *out_delta, int mem, void *stream);

/* the annealing batch */
int mcq_run(mcq_ctx *ctx, const mcq_run_params *params);

/* pinned host memory helpers for MCQ_MEM_HOST callers that want full-speed copies */
int mcq_host_alloc(void **ptr, uint64_t bytes);
int mcq_host_free(void *ptr);

/* Philox4x32-10 (Salmon et al., SC'11), host implementation identical to the device one */
void mcq_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
/* the device build of the same function: n calls, counters [n][4], keys [n][2], out [n][4] (HOST pointers) */
int mcq_philox4x32_10_device(mcq_ctx *ctx, int n, const uint32_t *counters, const uint32_t *keys, uint32_t *out);
/* Philox2x32-10, host and device builds: the words of board step s of a chain are
 * philox2x32_10(counter = (s, seed_lo), key = 0x243F6A88 ^ seed_hi).  Device form: counters [n][2], keys [n], out [n][2];
 * a key of 0x243F6A88 takes the compiled-constant path the kernels use for seeds below 2^32. */
void mcq_philox2x32_10(const uint32_t counter[2], uint32_t key, uint32_t out[2]);
int mcq_philox2x32_10_device(mcq_ctx *ctx, int n, const uint32_t *counters, const uint32_t *keys, uint32_t *out);

/* beta(step) of `n_groups` schedules evaluated on the device: out_beta[g][s] float64 (the value the float64 accept
 * rule uses), out_c[g][s] float32(-beta log2 e) (what the float32 fast path reads); either may be NULL.  HOST pointers. */
int mcq_beta_table(mcq_ctx *ctx, int n_groups, const mcq_schedule *schedules, int n_steps, double *out_beta, float *out_c);

#ifdef __cplusplus
}
#endif
#endif /* MCQ_H_ */
