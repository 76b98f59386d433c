// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw -- SC'11).
// One call turns (counter[4], key[2]) into four uniform 32-bit words; there is no state, so a
// chain's stream depends only on its key (the chain seed) and the step index, never on which
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
    for (int r = 0; r < 10; ++r) {
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

// counter domains (c3): keeps the streams of different uses of one chain key disjoint
enum : uint32_t { PHILOX_DOMAIN_STEP = 0u, PHILOX_DOMAIN_INIT = 1u };

}  // namespace mcq
