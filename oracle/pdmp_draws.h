/*
 * TEST INFRASTRUCTURE ONLY (oracle side) -- counter-based draw specification, CPU restatement.
 *
 * The product has its own device implementation of the same specification
 * (pdmpflux.jl_b200/csrc/philox.cuh); this file exists so the CPU oracle / CPU baseline consume
 * the *same* random numbers as the GPU when no draw tape is injected.
 *
 * Specification (not from the reference, which uses MersenneTwister(seed) -- AbstractPDMP.jl:100;
 * BASELINE.json north_star asks for Philox keyed by (seed, chain, event)):
 *   Philox4x32-10, key = (seed lo32, seed hi32),
 *   counter = (call index, event index, chain id lo32, stream | chain id hi bits << 8)
 *   stream 0 = E (randexp), 1 = U (rand), 2 = N (randn).  Slot counters restart at every event.
 *   E slot s : call s,    E = -log((k52 + 0.5) * 2^-52),  k52 = top 52 bits of (r1:r0)
 *   U slot s : call s,    U = k53 * 2^-53,                k53 = top 53 bits of (r1:r0)
 *   N slot j : call j>>1, Box-Muller on u1 = (k52(r1:r0)+0.5)*2^-52, u2 = k53(r3:r2)*2^-53:
 *              n = sqrt(-2 log u1) * (j&1 ? sin : cos)(2 pi u2)
 */
#ifndef PDMP_DRAWS_H
#define PDMP_DRAWS_H
#include <math.h>
#include <stdint.h>

static inline void pdmp_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                      uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void pdmp_draw_call(uint64_t seed, uint64_t chain, uint64_t event, uint32_t stream,
                                  uint32_t call, uint32_t out[4]) {
    pdmp_philox4x32_10(call, (uint32_t)event, (uint32_t)chain,
                       stream | ((uint32_t)(chain >> 32) << 8), (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

static inline double pdmp_u53(uint32_t lo, uint32_t hi) {
    uint64_t k = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)k * 0x1.0p-53;
}
static inline double pdmp_u52_open(uint32_t lo, uint32_t hi) {
    uint64_t k = (((uint64_t)hi << 32) | lo) >> 12;
    return ((double)k + 0.5) * 0x1.0p-52;
}

static inline double pdmp_draw_exp(uint64_t seed, uint64_t chain, uint64_t event, uint32_t slot) {
    uint32_t r[4];
    pdmp_draw_call(seed, chain, event, 0u, slot, r);
    return -log(pdmp_u52_open(r[0], r[1]));
}
static inline double pdmp_draw_uniform(uint64_t seed, uint64_t chain, uint64_t event, uint32_t slot) {
    uint32_t r[4];
    pdmp_draw_call(seed, chain, event, 1u, slot, r);
    return pdmp_u53(r[0], r[1]);
}
static inline double pdmp_draw_normal(uint64_t seed, uint64_t chain, uint64_t event, uint32_t slot) {
    uint32_t r[4];
    pdmp_draw_call(seed, chain, event, 2u, slot >> 1, r);
    double rad = sqrt(-2.0 * log(pdmp_u52_open(r[0], r[1])));
    double th = 6.283185307179586476925286766559 * pdmp_u53(r[2], r[3]);
    return rad * ((slot & 1u) ? sin(th) : cos(th));
}
#endif
