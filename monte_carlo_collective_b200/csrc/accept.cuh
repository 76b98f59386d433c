// Metropolis accept rule of the production kernels and the inverse-temperature schedules on the device.
//
// Reference (experiments.py:238-239, :326-327): accept iff u < min(1, exp(-beta_t * dE)) in float64, u drawn
// every step.  The production rule, shared by every kernel and stated independently in
// oracle/c/queens_philox.c:
//
//     dE <= 0, or  U < exp(-beta_s * dE)  in float64,
//     U = (z * 2^21 + (v >> 11)) / 2^53,  z = word 2 of the step's words, v = word 0 of its stream 0x80000000 (philox.cuh)
//
// Fast path (every lane, every round): t = 2^32 * ex2.approx(c_s * dE) in float32 with c_s = float32(-beta_s log2 e),
// compared with z.  Whenever z lies within the band |z - t| <= t * 2^-10 + 4 the decision is taken again with
// the float64 rule above (metropolis_exact), so the chain equals the float64 chain bit for bit; band hits and
// the decisions float32 alone would have got wrong are counted per chain.  The band covers every float32 error
// on the way with a wide margin: c_s rounding 2^-24, the product 2^-24, ex2.approx 2^-22, |c_s dE| <= 33 where the
// threshold is not 0 => relative error of t <= 2^-17; integer conversion of z <= 2^-25 relative + 1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace mcq {

// schedule kinds (experiments.py:13-77); values are part of the C ABI (MCQ_SCHED_*)
enum : int { SCHED_CONSTANT = 0, SCHED_LINEAR = 1, SCHED_EXPONENTIAL = 2, SCHED_LOGARITHMIC = 3, SCHED_SINUSOIDAL = 4 };

struct SchedDev {
    int type;
    int pad;
    double beta_const, beta_start, beta_end;
};

constexpr double MCQ_LOG2E = 1.4426950408889634;

// beta(step) in float64 with the reference's formulas and operation order (experiments.py:13-77; the same
// expressions as schedules.beta_table on the host).  Explicit _rn intrinsics: no FMA contraction.
__device__ __forceinline__ double sched_beta64(const SchedDev &sp, int n_steps, int s) {
    const double b0 = sp.beta_start, b1 = sp.beta_end;
    if (sp.type == SCHED_CONSTANT) return sp.beta_const;
    if (n_steps <= 1) return b1;
    const double sd = (double)s, n = (double)n_steps;
    switch (sp.type) {
        case SCHED_LINEAR:        // b0 + (s / (n-1)) * (b1 - b0)
            return __dadd_rn(b0, __dmul_rn(__ddiv_rn(sd, n - 1.0), __dsub_rn(b1, b0)));
        case SCHED_EXPONENTIAL:   // b0 * exp(log(b1 / b0) * (s / (n-1)))
            return __dmul_rn(b0, exp(__dmul_rn(log(__ddiv_rn(b1, b0)), __ddiv_rn(sd, n - 1.0))));
        case SCHED_LOGARITHMIC:   // b0 + (b1 - b0) * (log(1 + s) / log(1 + n))
            return __dadd_rn(b0, __dmul_rn(__dsub_rn(b1, b0), __ddiv_rn(log(1.0 + sd), log(1.0 + n))));
        default:                  // b0 + (b1 - b0) * (1 - cos(pi * s / n)) / 2
            return __dadd_rn(b0, __ddiv_rn(__dmul_rn(__dsub_rn(b1, b0), __dsub_rn(1.0, cos(__ddiv_rn(__dmul_rn(3.141592653589793, sd), n)))), 2.0));
    }
}

// float32 table of c_s = -beta_s * log2(e) for every (group, step), from the schedule parameters or from a
// float64 table of beta (closures tabulated on the host): what the fast path of every kernel reads.
static __global__ void beta_table_kernel(const SchedDev *sched, const double *beta64, int n_groups, int n_steps, float *out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_groups * n_steps) return;
    const int g = (int)(idx / n_steps), s = (int)(idx - (long long)g * n_steps);
    const double b = beta64 ? beta64[idx] : sched_beta64(sched[g], n_steps, s);
    out[idx] = (float)__dmul_rn(-b, MCQ_LOG2E);
}

__device__ __forceinline__ float ex2_approx_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Fast path: `acc` is the float32 decision, `near` says it must be retaken in float64.
// band_abs: absolute part of the band, 4 in production; +inf sends every uphill decision to the float64 rule (tests).
__device__ __forceinline__ void metropolis_fast(int dE, float cb, uint32_t z, float band_abs, bool &acc, bool &near) {
    const float t = ex2_approx_ftz(cb * (float)dE) * 4294967296.0f;
    const float d = t - __uint2float_rn(z);
    const bool up = dE > 0;
    acc = !up || d > 0.0f;
    near = up && fabsf(d) <= fmaf(t, 0.0009765625f, band_abs);
}

// The float64 rule.  Rare: about 1e-5 of the evaluated proposals.
__device__ __forceinline__ bool metropolis_exact(const SchedDev *sched, const double *beta64, int n_steps, int grp,
                                              uint32_t key0, uint32_t key1, int s, int dE, uint32_t z) {
    const double beta = beta64 ? beta64[(size_t)grp * n_steps + s] : sched_beta64(sched[grp], n_steps, s);
    const Philox4 v = chain_words((uint32_t)s, key0, key1, PHILOX_STREAM_UNIFORM_LO);
    const unsigned long long m = ((unsigned long long)z << 21) | (unsigned long long)(v.x >> 11);
    const double u = (double)m * (1.0 / 9007199254740992.0);
    return u < exp(__dmul_rn(-beta, (double)dE));
}

// the same rule as a real call: for kernels whose register budget the inlined float64 code would strain
static __device__ __noinline__ bool metropolis_exact_call(const SchedDev *sched, const double *beta64, int n_steps, int grp,
                                                   uint32_t key0, uint32_t key1, int s, int dE, uint32_t z) {
    return metropolis_exact(sched, beta64, n_steps, grp, key0, key1, s, dE, z);
}

}  // namespace mcq
