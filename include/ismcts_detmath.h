/*
 * ismcts_detmath.h — arithmetic that must give the same bits on the host and
 * on the device.
 *
 * The reference evaluates UCB1 (utility.h:96-99) and EXP3 (exp3.h:73-95) with
 * libm's log/exp/sqrt.  glibc and CUDA's libdevice differ in the last bits of
 * log and exp, so a search compared bit for bit between the CUDA kernels and
 * the CPU oracle may only use operations IEEE-754 defines exactly: +, -, *, /,
 * sqrt, floor, each rounded once (no fused multiply-add contraction).  This
 * header provides those operations under one name for both compilers and an
 * exp() built from them.  ln(n) for integer n comes from a host-built table
 * that both sides read.
 *
 * Host code including this header must be compiled with -ffp-contract=off.
 */
#ifndef ISMCTS_DETMATH_H
#define ISMCTS_DETMATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDA_ARCH__)
#define ISMCTS_HD __host__ __device__ __forceinline__
#define ISMCTS_DADD(a, b) __dadd_rn((a), (b))
#define ISMCTS_DSUB(a, b) __dsub_rn((a), (b))
#define ISMCTS_DMUL(a, b) __dmul_rn((a), (b))
#define ISMCTS_DDIV(a, b) __ddiv_rn((a), (b))
#define ISMCTS_DSQRT(a) __dsqrt_rn((a))
#else
#if defined(__CUDACC__)
#define ISMCTS_HD __host__ __device__ inline
#else
#define ISMCTS_HD static inline
#endif
#define ISMCTS_DADD(a, b) ((double)(a) + (double)(b))
#define ISMCTS_DSUB(a, b) ((double)(a) - (double)(b))
#define ISMCTS_DMUL(a, b) ((double)(a) * (double)(b))
#define ISMCTS_DDIV(a, b) ((double)(a) / (double)(b))
#define ISMCTS_DSQRT(a) sqrt((double)(a))
#endif

/* UCB1 value, utility.h:96-99 with X = score/visits (ucb1.h:36-39):
 *   score/visits + C * sqrt(ln(available) / visits)
 * evaluated from three factors that depend on ONE integer each,
 *   (score2 / 2) * [1 / visits]  +  C * ([sqrt(ln(available))] * [1 / sqrt(visits)]),
 * score2 = 2*score (rewards are multiples of 1/2).  The bracketed factors are the table entries below: the
 * kernels read them from tables the host builds with these very functions, the oracle calls the functions —
 * the same bits on both sides, and a child evaluation on the GPU is two table look-ups, three multiplications
 * and an addition instead of two double-precision divisions and a square root (which were 8 % of the
 * instructions of a search with deep trees, profiles/r02_hybrid.md).  The value differs from the reference's
 * expression by a few units in the last place (a deviation like that of any other libm); the order of the
 * children it induces is what matters, and is pinned against the reference statistically (tests/test_ref_stats.py,
 * tests/test_gpu_geometry.py). */
ISMCTS_HD double ismcts_ucb_value(uint32_t score2, double inv_visits, double inv_sqrt_visits, double sqrt_ln_avail,
                                  double exploration)
{
    /* 0 * x is exactly 0: many nodes have no win yet */
    const double x = score2 ? ISMCTS_DMUL(ISMCTS_DMUL((double)score2, 0.5), inv_visits) : 0.0;
    const double e = ISMCTS_DMUL(exploration, ISMCTS_DMUL(sqrt_ln_avail, inv_sqrt_visits));
    return ISMCTS_DADD(x, e);
}

/* The table entries (host code: log() is libm's; division and square root are IEEE's, rounded once).  Entry 0 of
 * the visit factors is never used: a child that can be selected has been visited. */
#if defined(__CUDACC__)
#define ISMCTS_HOST_ONLY __host__ static inline
#else
#define ISMCTS_HOST_ONLY static inline
#endif
ISMCTS_HOST_ONLY double ismcts_tab_inv_visits(uint32_t v) { return v ? 1.0 / (double)v : 0.0; }
ISMCTS_HOST_ONLY double ismcts_tab_inv_sqrt_visits(uint32_t v) { return v ? 1.0 / sqrt((double)v) : 0.0; }
ISMCTS_HOST_ONLY double ismcts_tab_sqrt_ln(uint32_t a) { return a ? sqrt(log((double)a)) : 0.0; }

/* exp(x) for x <= 0 (callers subtract the maximum first), relative error
 * about 2e-16, from exactly-rounded operations only. */
ISMCTS_HD double ismcts_det_exp(double x)
{
    if (!(x > -700.0)) return 0.0;
    if (x > 0.0) x = 0.0;
    const double kf = floor(ISMCTS_DADD(ISMCTS_DMUL(x, 1.4426950408889634074), 0.5));
    /* r = x - k*ln2 in two steps (Cody-Waite) */
    double r = ISMCTS_DSUB(x, ISMCTS_DMUL(kf, 6.93147180369123816490e-01));
    r = ISMCTS_DSUB(r, ISMCTS_DMUL(kf, 1.90821492927058770002e-10));
    /* Taylor series to r^13 by Horner; |r| <= 0.3466 so the tail is < 1e-17 */
    double p = 1.0 / 6227020800.0;
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 479001600.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 39916800.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 3628800.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 362880.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 40320.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 5040.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 720.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 120.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 24.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0 / 6.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 0.5);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0);
    p = ISMCTS_DADD(ISMCTS_DMUL(p, r), 1.0);
    /* scale by 2^k: k is in [-1010, 0], so 2^k is a normal double */
    const int64_t k = (int64_t)kf;
    const uint64_t bits = (uint64_t)(k + 1023) << 52;
    double scale;
#if defined(__CUDA_ARCH__)
    scale = __longlong_as_double((long long)bits);
#else
    memcpy(&scale, &bits, sizeof scale);
#endif
    return ISMCTS_DMUL(p, scale);
}

/* EXP3 exploration rate, exp3.h:91-94: min(1/K, sqrt(ln K / K / t)) with
 * std::min's semantics (returns the first argument when the second is NaN). */
ISMCTS_HD double ismcts_exp3_epsilon(uint32_t K, double ln_K, double t)
{
    const double a = ISMCTS_DDIV(1.0, (double)K);
    const double b = ISMCTS_DSQRT(ISMCTS_DDIV(ISMCTS_DDIV(ln_K, (double)K), t));
    return (b < a) ? b : a;
}

#endif /* ISMCTS_DETMATH_H */
