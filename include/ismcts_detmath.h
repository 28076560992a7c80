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
 * score2 = 2*score (rewards are multiples of 1/2), ln_avail = ln(available). */
ISMCTS_HD double ismcts_ucb_value(uint32_t score2, uint32_t visits, double ln_avail, double exploration)
{
    const double v = (double)visits;
    /* 0 / v is exactly 0: many nodes have no win yet, and a zero numerator takes the slow path of the GPU's division */
    const double x = score2 ? ISMCTS_DDIV(ISMCTS_DMUL((double)score2, 0.5), v) : 0.0;
    const double e = ISMCTS_DMUL(exploration, ISMCTS_DSQRT(ISMCTS_DDIV(ln_avail, v)));
    return ISMCTS_DADD(x, e);
}

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
