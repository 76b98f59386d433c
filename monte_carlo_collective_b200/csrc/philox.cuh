// Philox4x32-10 and Philox2x32-10 counter-based generators (Salmon, Moraes, Dror, Shaw -- SC'11).
// One call turns (counter[4], key[2]) into four uniform 32-bit words (counter[2], key into two); there is no state, so a
// chain's stream depends only on the chain seed and the step index, never on which
// GPU / CTA / lane runs it.  The reference draws from NumPy's global MT19937 instead
// (experiments.py:221-239, :311-327); parity under independent RNG is statistical.
#pragma once
#include <stdint.h>

namespace mcq {

#if defined(__CUDA_ARCH__)
#define MCQ_HD __host__ __device__ __forceinline__
#else
#define MCQ_HD inline
#endif

MCQ_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

struct Philox4 {
    uint32_t x, y, z, w;
};

MCQ_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
#ifdef MCQ_DBG_PHILOX7   // sensitivity probe only (not a product configuration)
    for (int r = 0; r < 7; ++r) {
#else
    for (int r = 0; r < 10; ++r) {
#endif
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    Philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Philox2x32-10 (same paper): two words per call at about half the cost -- what a board step needs.
struct Philox2 {
    uint32_t x, z;   // proposal word, uniform word
};

MCQ_HD Philox2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k) {
    const uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi = mulhi32(M, c0), lo = M * c0;
        c0 = hi ^ k ^ c1;
        c1 = lo;
        k += W;
    }
    Philox2 o;
    o.x = c0; o.z = c1;
    return o;
}

// How a chain draws its words.  The chain's 64-bit seed sits in the COUNTER, the cipher key is a constant:
//     chain_words(i, seed, stream) = Philox4x32-10(counter = (i, seed_lo, seed_hi, stream), key = CHAIN_KEY)
// so the ten round keys are immediates in the compiled code (no per-chain key registers, no key arithmetic) and
// a chain's words still depend on (seed, i, stream) only.  Streams of a step i = s:
//     0               the four words of a full_3d step (proposal, uniform); a board step uses board_step_words below
//     1 + k           further candidate cells of a full_3d proposal (k = 0, 1, ...)
//     0x80000000      low bits of the step's 53-bit uniform (float64 accept rule)
// and of the initial state (i = block of four words): 0x40000000.
constexpr uint32_t CHAIN_KEY0 = 0x243F6A88u, CHAIN_KEY1 = 0x85A308D3u;   // first 64 fractional bits of pi
enum : uint32_t { PHILOX_STREAM_STEP = 0u, PHILOX_STREAM_INIT = 0x40000000u, PHILOX_STREAM_UNIFORM_LO = 0x80000000u };

MCQ_HD Philox4 chain_words(uint32_t i, uint32_t seed_lo, uint32_t seed_hi, uint32_t stream) {
    return philox4x32_10(i, seed_lo, seed_hi, stream, CHAIN_KEY0, CHAIN_KEY1);
}

// The two words of a BOARD step s (proposal, uniform): Philox2x32-10(counter = (s, seed_lo), key = CHAIN_KEY0 ^ seed_hi).
// A board proposal is one index over the N^2 (N - 1) (column, other height) pairs and the accept test one uniform:
// 64 bits.  Seeds below 2^32 (every seed the reference can express: np.random.seed takes 32 bits) leave the key
// a compile-time constant, so the two cases are compiled apart -- same function, ten immediates instead of ten adds.
MCQ_HD Philox2 board_step_words(uint32_t s, uint32_t seed_lo, uint32_t seed_hi) {
    if (seed_hi == 0u) return philox2x32_10(s, seed_lo, CHAIN_KEY0);
    return philox2x32_10(s, seed_lo, CHAIN_KEY0 ^ seed_hi);
}

// The words of step s in the layout every kernel decodes: x -> mulhi(x, N^2) the column (full_3d: mulhi(x, Q) the
// queen), y -> mulhi(y, N - 1) the height offset (board: y = lo32(x * N^2), the next mixed-radix digit of x; full_3d:
// first candidate cell), z the uniform word, w (full_3d only) the second candidate cell.
template <bool FULL>
MCQ_HD Philox4 step_words(uint32_t s, uint32_t seed_lo, uint32_t seed_hi, uint32_t n2) {
    if (FULL) return chain_words(s, seed_lo, seed_hi, PHILOX_STREAM_STEP);
    const Philox2 b = board_step_words(s, seed_lo, seed_hi);
    Philox4 o;
    o.x = b.x; o.y = b.x * n2; o.z = b.z; o.w = 0u;
    return o;
}

}  // namespace mcq
