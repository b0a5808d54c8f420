/*
 * kv_detmath.h — deterministic fp32 arithmetic shared by the CUDA kernels and the CPU oracle.
 *
 * MCTS visit counts must be bit-exact between the GPU engine and the sequential oracle (north_star), so every
 * floating-point quantity that steers the search (PUCT score, softmax priors, Dirichlet noise) is computed from
 * IEEE-754 single-precision +, -, *, /, sqrt and integer operations only, in a fixed order, with no FMA
 * contraction (nvcc -fmad=false, gcc -ffp-contract=off): both sides then produce identical bits.
 * exp/log follow the classic Cephes single-precision polynomial scheme.
 * C and CUDA C++ compatible.  No state.
 */
#ifndef KV_DETMATH_H
#define KV_DETMATH_H
#include <stdint.h>

#if defined(__CUDACC__) && !defined(KV_HOST_EMU)
#define KVD_FN __host__ __device__ static __forceinline__
#define KVD_SQRTF(x) sqrtf(x)
#else
#include <math.h>
#define KVD_FN static inline
#define KVD_SQRTF(x) sqrtf(x)
#endif

KVD_FN uint32_t kvd_f2u(float f) {
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
}
KVD_FN float kvd_u2f(uint32_t u) {
    union { float f; uint32_t u; } c;
    c.u = u;
    return c.f;
}

/* exp(x), x clamped to [-87, 88] */
KVD_FN float kvd_expf(float x) {
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float t = x * 1.44269504088896341f;
    int k = (int)(t + (t < 0.0f ? -0.5f : 0.5f));
    float kf = (float)k;
    float r = x - kf * 0.693359375f;
    r = r - kf * -2.12194440e-4f;
    float p = 1.9875691500e-4f;
    p = p * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    float rr = r * r;
    p = p * rr + r;
    p = p + 1.0f;
    return p * kvd_u2f((uint32_t)(k + 127) << 23);
}

/* log(x) for normal x > 0 */
KVD_FN float kvd_logf(float x) {
    uint32_t u = kvd_f2u(x);
    int e = (int)((u >> 23) & 0xFF) - 126;
    float m = kvd_u2f((u & 0x007FFFFFu) | 0x3F000000u); /* [0.5, 1) */
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = m + m - 1.0f;
    } else {
        m = m - 1.0f;
    }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = p * m + -1.1514610310e-1f;
    p = p * m + 1.1676998740e-1f;
    p = p * m + -1.2420140846e-1f;
    p = p * m + 1.4249322787e-1f;
    p = p * m + -1.6668057665e-1f;
    p = p * m + 2.0000714765e-1f;
    p = p * m + -2.4999993993e-1f;
    p = p * m + 3.3333331174e-1f;
    float y = m * z * p;
    float ef = (float)e;
    y = y + ef * -2.12194440e-4f;
    y = y - 0.5f * z;
    float r = m + y;
    r = r + ef * 0.693359375f;
    return r;
}

KVD_FN uint64_t kvd_mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
/* counter-based generator: 24 uniform bits from (seed, a, b, c) */
KVD_FN uint32_t kvd_rand24(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    uint64_t h = kvd_mix64(seed ^ 0x9E3779B97F4A7C15ull);
    h = kvd_mix64(h + a * 0xD6E8FEB86659FD93ull);
    h = kvd_mix64(h + b * 0xA0761D6478BD642Full);
    h = kvd_mix64(h + c * 0xE7037ED1A0B428DBull);
    return (uint32_t)(h >> 40);
}
/* 48 uniform bits from one hash (two 24-bit uniforms per call) */
KVD_FN uint64_t kvd_rand48(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    uint64_t h = kvd_mix64(seed ^ 0x9E3779B97F4A7C15ull);
    h = kvd_mix64(h + a * 0xD6E8FEB86659FD93ull);
    h = kvd_mix64(h + b * 0xA0761D6478BD642Full);
    h = kvd_mix64(h + c * 0xE7037ED1A0B428DBull);
    return h >> 16;
}
/* uniform in (0,1): (bits + 0.5) / 2^24, exact in fp32 */
KVD_FN float kvd_u01(uint32_t bits24) { return ((float)bits24 + 0.5f) * 5.9604644775390625e-8f; }

/* Gamma(alpha, 1) for 0 < alpha < 1, Ahrens-Dieter GS rejection, at most 24 rounds; both uniforms of a round come from
 * one hash (counter c = round): the hash is most of the cost, and the reference-rule root noise draws 4096 variates per
 * position */
KVD_FN float kvd_gamma_small(float alpha, uint64_t seed, uint64_t a, uint64_t b) {
    const float e = 2.718281828459045f;
    const float bb = (e + alpha) / e;
    const float inv_alpha = 1.0f / alpha;
    float x = 0.5f;
    for (int round = 0; round < 24; round++) {
        const uint64_t r48 = kvd_rand48(seed, a, b, (uint64_t)round);
        float u1 = kvd_u01((uint32_t)(r48 >> 24));
        float u2 = kvd_u01((uint32_t)(r48 & 0xFFFFFFu));
        float p = bb * u1;
        if (p <= 1.0f) {
            x = kvd_expf(kvd_logf(p) * inv_alpha);
            if (u2 <= kvd_expf(-x)) break;
        } else {
            x = -kvd_logf((bb - p) * inv_alpha);
            if (u2 <= kvd_expf(kvd_logf(x) * (alpha - 1.0f))) break;
        }
    }
    if (!(x > 1e-30f)) x = 1e-30f;
    return x;
}

/* PUCT score of one edge (SPEC: DESIGN.md §MCTS).  q = W/N (0 when unvisited); sqrt_parent = sqrtf(parent visits) */
KVD_FN float kvd_puct(float W, uint32_t N, float P, float sqrt_parent, float c_puct) {
    float q = N ? W / (float)N : 0.0f;
    float u = (c_puct * P) * (sqrt_parent / (float)(1u + N));
    return q + u;
}

#endif /* KV_DETMATH_H */
