/* tiktak_rng.h -- the counter-based random numbers of the hot path, shared by host, device and oracle.
 *
 * The reference draws from the Fortran intrinsic RANDOM_NUMBER (getLotteryPoint, TiktakGlobalSearch.f90:758),
 * seeded with seed(1) = 123456 (genericParams.f90:65-69); that stream is compiler-specific and sequential.
 * Here every draw is a pure function of (seed, stream, counter): Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
 * "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter = (ctr_lo, ctr_hi, stream, 0), key = (seed, 0),
 * so restart k draws the same number on any GPU, in any order, and in the CPU oracle.
 */
#ifndef TIKTAK_RNG_H
#define TIKTAK_RNG_H
#include <stdint.h>

#include "tiktak_math.h"

#define TK_LOTTERY_SEED 123456u /* seed(1) of genericParams.f90:67 */
#define TK_LOTTERY_STREAM 7u

TK_HD void tk_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* uniform in [0, 1) with 53 random bits, like RANDOM_NUMBER's range (0 <= x < 1) */
TK_HD double tk_uniform01(uint32_t seed, uint32_t stream, uint64_t counter) {
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), stream, 0u};
    tk_philox4x32_10(c, seed, 0u);
    const uint64_t bits = (((uint64_t)c[0] << 32) | c[1]) >> 11; /* 53 bits */
    return (double)bits * 1.1102230246251565e-16;                /* 2^-53 */
}

#endif /* TIKTAK_RNG_H */
