/*
 * oracle/philox.h — TEST INFRASTRUCTURE. Host restatement of Philox4x32-10 (Salmon et al., SC'11,
 * "Parallel random numbers: as easy as 1, 2, 3") and of the engine's slot-addressed random-number
 * convention (include/aptb200.h, "Random-number slots").  The reference draws from R's sequential
 * global RNG (rnorm APT_functions.R:696,699; rcat :459; runif :463; rmnorm_chol :843); a sequential
 * stream cannot be shared with a massively parallel engine, so both the oracle and the CUDA engine
 * read the SAME counter-addressed slots instead.
 */
#ifndef APT_ORACLE_PHILOX_H
#define APT_ORACLE_PHILOX_H
#include <math.h>
#include <stdint.h>

static inline void apt_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 52-bit uniform strictly inside (0,1) from two 32-bit words (2^52 - 1/2 is representable, 2^53 - 1/2 is not) */
static inline double apt_u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 6) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 4503599627370496.0);
}

#define APT_SLOT_ACCEPT 0x8000u
#define APT_UNIT_SWAP 0xFFFFFFFFu

static inline void apt_slot_words(uint32_t seed, uint32_t ladder, uint64_t iter, uint32_t unit, uint32_t rung,
                                  uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)iter, (uint32_t)(iter >> 32), unit, (rung << 16) | block};
    uint32_t key[2] = {seed, ladder};
    apt_philox4x32_10(ctr, key, out);
}

/* two standard normals from slot block `block` of an update */
static inline void apt_slot_normal2(uint32_t seed, uint32_t ladder, uint64_t iter, uint32_t unit, uint32_t rung,
                                    uint32_t block, double* z0, double* z1) {
    uint32_t w[4];
    apt_slot_words(seed, ladder, iter, unit, rung, block, w);
    double u1 = apt_u53(w[0], w[1]), u2 = apt_u53(w[2], w[3]);
    double rad = sqrt(-2.0 * log(u1));
    double ang = 6.283185307179586476925286766559 * u2;
    *z0 = rad * cos(ang);
    *z1 = rad * sin(ang);
}

static inline double apt_slot_accept_u(uint32_t seed, uint32_t ladder, uint64_t iter, uint32_t unit, uint32_t rung) {
    uint32_t w[4];
    apt_slot_words(seed, ladder, iter, unit, rung, APT_SLOT_ACCEPT, w);
    return apt_u53(w[0], w[1]);
}

static inline void apt_slot_swap_u(uint32_t seed, uint32_t ladder, uint64_t iter, uint32_t proposal, double* ucat,
                                   double* uacc) {
    uint32_t w[4];
    apt_slot_words(seed, ladder, iter, APT_UNIT_SWAP, proposal, 0u, w);
    *ucat = apt_u53(w[0], w[1]);
    *uacc = apt_u53(w[2], w[3]);
}
#endif
