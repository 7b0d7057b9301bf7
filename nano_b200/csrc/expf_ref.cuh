// expf_ref.cuh -- exact-mode expf: a device (and host) evaluation of the table-driven double-precision
// exp2f/expf algorithm that glibc >= 2.27 uses for expf() (S. Nagy's "optimized-routines" expf:
// N = 32 entry table of 2^(i/N), cubic in double, one final rounding to float).  The reference engine
// calls libm expf in softmax (infer.c:625) and SwiGLU (infer.c:940); exact mode needs the same bits,
// and CUDA's own expf is only <= 2 ulp.  The table below was recomputed from the definition
// T[i] = bits(2^(i/32)) - (i << 47) with 60-digit decimals (not copied from glibc sources);
// tests/test_expf_ref.py checks this function against the host libm over the float range used.
//
// glibc is not part of /root/reference (it is the reference's C library dependency, glibc 2.39 in this
// image); the algorithm is restated from its published description.
#pragma once
#include <stdint.h>
#include <string.h>

namespace nb {

#define NB_EXP2F_TAB \
0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
        0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
        0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
        0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
        0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
        0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
        0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
        0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL

#ifdef __CUDACC__
__constant__ uint64_t kExp2fTabDev[32] = {NB_EXP2F_TAB};
#endif

__host__ __device__ inline double nb_u64_as_double(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
__host__ __device__ inline uint64_t nb_double_as_u64(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
__host__ __device__ inline double nb_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
__host__ __device__ inline double nb_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

// Polynomial evaluated without contraction (matches a glibc built without FMA contraction; the fused
// variant differs from this one in < 1e-8 of inputs -- see tests/test_expf_ref.py for the measured rate).
__host__ __device__ inline float expf_ref_impl(float x) {
    const double kInvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double kShift = 0x1.8p+52;
    const double c0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double c1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double c2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    uint32_t ux;
#ifdef __CUDA_ARCH__
    ux = __float_as_uint(x);
#else
    memcpy(&ux, &x, 4);
#endif
    const uint32_t abstop = (ux >> 20) & 0x7ff;
    if (abstop >= 0x42b) {                 // |x| >= 88 or NaN/Inf
        if (ux == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return x + x;
        if (x > 0x1.62e42ep6f) return x * 0x1p127f;      // overflow -> +inf
        if (x < -0x1.9fe368p6f) return 0.0f;             // underflow -> 0
    }
    const double xd = (double)x;
    const double z = nb_mul(kInvLn2N, xd);
    double kd = nb_add(z, kShift);
    const uint64_t ki = nb_double_as_u64(kd);
    kd = nb_add(kd, -kShift);
    const double r = nb_add(z, -kd);
#ifdef __CUDA_ARCH__
    uint64_t t = kExp2fTabDev[ki & 31];
#else
    static const uint64_t tab[32] = {NB_EXP2F_TAB};
    uint64_t t = tab[ki & 31];
#endif
    t += ki << (52 - 5);
    const double s = nb_u64_as_double(t);
    const double zz = nb_add(nb_mul(c0, r), c1);
    const double r2 = nb_mul(r, r);
    double y = nb_add(nb_mul(c2, r), 1.0);
    y = nb_add(nb_mul(zz, r2), y);
    y = nb_mul(y, s);
    return (float)y;
}

#ifdef __CUDACC__
__device__ inline float expf_ref(float x) { return expf_ref_impl(x); }
#endif

}  // namespace nb
