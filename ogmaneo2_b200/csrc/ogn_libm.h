// ogn_libm.h — bit-exact replays of glibc 2.39 x86-64 expf() and tanhf(), callable on host and device.
//
// Why this exists: the reference's learn rules call std::exp(float) (SparseCoder.cpp:78, Actor.cpp:65,167) and
// std::tanh(float) (Predictor.cpp:61), i.e. glibc's expf/tanhf. CUDA's expf/tanhf round differently, and a
// 1-ULP difference in a weight eventually flips a column arg-max. To keep free-running CSDR
// trajectories bit-identical to the reference we replay glibc's *algorithms* (published in glibc's
// sysdeps/ieee754/flt-32/{e_expf.c,s_tanhf.c,s_expm1f.c}; expf is the ARM optimized-routines kernel: 2^(k/N)
// table with N=32 plus a degree-3 polynomial evaluated in double) with IEEE fp32/fp64 arithmetic only.
// No code is shared with the reference; the restatement is pinned by tests/test_libm_replay.py, which compares
// it with the box's own libm on ~2^32 inputs (CPU) and on a dense sample on the device.
//
// Requirements on the including translation unit: no FMA contraction (nvcc -fmad=false, gcc -ffp-contract=off),
// IEEE division (nvcc -prec-div=true). The fused operations glibc's x86-64 FMA build performs are written as
// explicit fma() calls below.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define OGN_HD __host__ __device__ __forceinline__
#else
#define OGN_HD static inline
#endif

#ifndef OGN_EXPF_USE_FMA
// glibc's x86-64 multiarch expf selects the FMA build on any CPU with AVX2+FMA (every host this runs on).
#define OGN_EXPF_USE_FMA 1
#endif
#ifndef OGN_EXPF_FMA_K
#define OGN_EXPF_FMA_K 1
#endif

OGN_HD uint32_t ogn_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
OGN_HD float ogn_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
OGN_HD uint64_t ogn_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
OGN_HD double ogn_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// bits(2^(i/32)) - (i << 47), i = 0..31  (exp2f table of the optimized-routines expf, EXP2F_TABLE_BITS = 5)
#if defined(__CUDACC__)
__device__ __constant__ uint64_t ogn_exp2f_tab_dev[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
#endif
static const uint64_t ogn_exp2f_tab_host[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

OGN_HD double ogn_fma_or_not(double a, double b, double c) {
#if OGN_EXPF_USE_FMA
    return fma(a, b, c);
#else
    return a * b + c;
#endif
}

// glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c, TOINT_INTRINSICS == 0 on x86-64)
OGN_HD float ogn_expf(float x) {
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    double xd = (double)x;
    uint32_t abstop = (ogn_f2u(x) >> 20) & 0x7ff;
    if (abstop >= 0x42b) { // |x| >= 88 or nan (top12(88.0f) == 0x42b)
        if (ogn_f2u(x) == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return x + x;
        if (x > 0x1.62e42ep6f) return ogn_u2f(0x7f800000u);  // overflow -> +inf
        if (x < -0x1.9fe368p6f) return 0.0f;                 // underflow -> +0
    }
    double z = InvLn2N * xd;
#if OGN_EXPF_FMA_K
    double kd = fma(InvLn2N, xd, Shift);
#else
    double kd = z + Shift;
#endif
    uint64_t ki = ogn_d2u(kd);
    kd -= Shift;
    // GCC (-ffp-contract=fast) fuses the product into this subtraction in glibc's FMA build of expf; the exhaustive
    // comparison in tests/test_libm_replay.py only passes with this form.
    double r = fma(InvLn2N, xd, -kd);
#if defined(__CUDA_ARCH__)
    uint64_t t = ogn_exp2f_tab_dev[ki % 32];
#else
    uint64_t t = ogn_exp2f_tab_host[ki % 32];
#endif
    t += ki << (52 - 5);
    double s = ogn_u2d(t);
    z = ogn_fma_or_not(C0, r, C1);
    double r2 = r * r;
    double y = ogn_fma_or_not(C2, r, 1.0);
    y = ogn_fma_or_not(z, r2, y);
    y = y * s;
    return (float)y;
}

// glibc 2.39 expm1f (fdlibm float version, sysdeps/ieee754/flt-32/s_expm1f.c), pure fp32 arithmetic.
OGN_HD float ogn_expm1f(float x) {
    const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f;
    const float o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f,
                invln2 = 1.4426950216e+00f;
    const float Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f,
                Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
    int32_t k;
    uint32_t hx = ogn_f2u(x);
    uint32_t xsb = hx & 0x80000000u;
    hx &= 0x7fffffffu;

    if (hx >= 0x4195b844u) {          // |x| >= 27*ln2
        if (hx >= 0x42b17218u) {      // |x| >= 88.721...
            if (hx > 0x7f800000u) return x + x;
            if (hx == 0x7f800000u) return (xsb == 0) ? x : -1.0f;
            if (x > o_threshold) return huge * huge;
        }
        if (xsb != 0) return tiny - one;
    }
    if (hx > 0x3eb17218u) {           // |x| > 0.5 ln2
        if (hx < 0x3F851592u) {       // |x| < 1.5 ln2
            if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int32_t)(invln2 * x + ((xsb == 0) ? 0.5f : -0.5f));
            t = (float)k;
            hi = x - t * ln2_hi;
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x33000000u) {    // |x| < 2**-25
        t = huge + x;
        return x - (t - (huge + x));
    } else
        k = 0;

    hfx = 0.5f * x;
    hxs = x * hfx;
    r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0f - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0f - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (k == -1) return 0.5f * (x - e) - 0.5f;
    if (k == 1) {
        if (x < -0.25f) return -2.0f * (e - (x + 0.5f));
        return one + 2.0f * (x - e);
    }
    if (k <= -2 || k > 56) {
        y = one - (e - x);
        y = ogn_u2f(ogn_f2u(y) + ((uint32_t)k << 23));
        return y - one;
    }
    if (k < 23) {
        t = ogn_u2f(0x3f800000u - (0x1000000u >> k));
        y = t - (e - x);
        y = ogn_u2f(ogn_f2u(y) + ((uint32_t)k << 23));
    } else {
        t = ogn_u2f((uint32_t)(0x7f - k) << 23);
        y = x - (e + t);
        y += one;
        y = ogn_u2f(ogn_f2u(y) + ((uint32_t)k << 23));
    }
    return y;
}

// glibc 2.39 tanhf (fdlibm float version, sysdeps/ieee754/flt-32/s_tanhf.c)
OGN_HD float ogn_tanhf(float x) {
    const float one = 1.0f, two = 2.0f, tiny = 1.0e-30f;
    float t, z;
    uint32_t jx = ogn_f2u(x);
    uint32_t ix = jx & 0x7fffffffu;
    if (ix >= 0x7f800000u) {
        if ((int32_t)jx >= 0) return one / x + one;
        return one / x - one;
    }
    if (ix < 0x41b00000u) {           // |x| < 22
        if (ix == 0) return x;
        if (ix < 0x24000000u) return x * (one + x);
        if (ix >= 0x3f800000u) {      // |x| >= 1
            t = ogn_expm1f(two * fabsf(x));
            z = one - two / (t + two);
        } else {
            t = ogn_expm1f(-two * fabsf(x));
            z = -t / (t + two);
        }
    } else {
        z = one - tiny;
    }
    return ((int32_t)jx >= 0) ? z : -z;
}
