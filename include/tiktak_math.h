/* tiktak_math.h -- deterministic FP64 elementary functions shared by host and device.
 *
 * Why this exists: the local searches on the hot path (amoeba, minimize.f90:7-135; BOBYQB_H,
 * minimize.f90:405-1171) branch on floating-point comparisons of objective values, so a 1-ulp
 * difference between the host libm `cos` and CUDA's `cos` can flip a branch and make a GPU restart
 * follow a different trajectory from the CPU one.  Every synthetic objective plugin therefore calls
 * the functions below, which are built ONLY from IEEE-754 +,-,* and explicitly written fma(), all of
 * which round identically on x86-64 and on sm_100a.  The result is bit-identical on both sides
 * (tests/test_math.py pins that, and checks the functions against libm to <= 2 ulp).
 *
 * The reference itself calls the compiler intrinsic COS (objective.f90:103-106); the oracle can be
 * built against libm (faithful) or against this header (trajectory-exact) -- see oracle/README.
 *
 * Polynomial coefficients: the classic fdlibm k_sin.c / k_cos.c minimax sets (Sun Microsystems,
 * freely distributable); argument reduction: 3-term Cody-Waite with fma, exact first step.
 * Accuracy domain: |x| < 2^30 (error <= ~1.5 ulp); outside it NaN is returned.
 */
#ifndef TIKTAK_MATH_H
#define TIKTAK_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define TK_HD __host__ __device__ __forceinline__
#else
#define TK_HD static inline
#endif

/* Unfused primitives.  On the device nvcc contracts a*b+c into FMA unless told otherwise; the
 * reference is compiled without FMA (ifort default target, Makefile:4-18), so every formula that
 * restates reference arithmetic goes through these. Host builds use -ffp-contract=off. */
#if defined(__CUDA_ARCH__)
#define TK_MUL(a, b) __dmul_rn((a), (b))
#define TK_ADD(a, b) __dadd_rn((a), (b))
#define TK_SUB(a, b) __dsub_rn((a), (b))
#define TK_DIV(a, b) __ddiv_rn((a), (b))
#define TK_SQRT(a) __dsqrt_rn((a))
#define TK_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define TK_MUL(a, b) ((a) * (b))
#define TK_ADD(a, b) ((a) + (b))
#define TK_SUB(a, b) ((a) - (b))
#define TK_DIV(a, b) ((a) / (b))
#define TK_SQRT(a) sqrt((a))
#define TK_FMA(a, b, c) fma((a), (b), (c))
#endif

TK_HD uint64_t tk_d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
TK_HD double tk_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

#define TK_PIO2_1 1.5707963267948966        /* 0x3ff921fb54442d18 */
#define TK_PIO2_2 6.123233995736766e-17     /* 0x3c91a62633145c07 */
#define TK_PIO2_3 (-1.4973849048591698e-33) /* 0xb91f1976b7ed8fbc */
#define TK_2OPI 0.6366197723675814          /* 0x3fe45f306dc9c883 */
#define TK_RMAGIC 6755399441055744.0        /* 1.5 * 2^52 */

/* Shared core: reduce x = k*pi/2 + r, |r| <= pi/4(1+eps), n = k + shift (shift = 1 for cos, 0 for
 * sin: cos(x) = sin(x + pi/2)); then sin(x + shift*pi/2) = (n&1 ? cos(r) : sin(r)), negated when n&2.
 *     sin(r) = r + (z*r)*S(z)        cos(r) = (1 - z/2) + (z*z)*C(z)        z = r*r
 * Host and device run the same operations on the branch that is selected.  The device version
 * evaluates BOTH polynomials and selects the result: that costs 5 more DFMA but removes ~24 per-thread
 * coefficient selects, which is what bounded the Sobol kernel (issue slots, not the FP64 pipe), and it
 * reads the coefficients from one __constant__ table (two doubles per LDCU.128). */
#if defined(__CUDACC__)
static __constant__ double tk_sc_tab[16] = {
    /* S1..S6 */ -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    /* C1..C6 */ 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    /* reduction */ TK_2OPI, TK_PIO2_1, TK_PIO2_2, TK_PIO2_3};
#endif

TK_HD double tk_sincos_core(double x, int shift) {
#if defined(__CUDA_ARCH__)
    const double* const T = tk_sc_tab;
    const double t = __fma_rn(x, T[12], TK_RMAGIC);
    const int n = __double2loint(t) + shift;
    const double kd = __dsub_rn(t, TK_RMAGIC);
    double r = __fma_rn(-kd, T[13], x);
    r = __fma_rn(-kd, T[14], r);
    r = __fma_rn(-kd, T[15], r);
    const double z = __dmul_rn(r, r);
    double ps = __fma_rn(z, T[5], T[4]);
    double pc = __fma_rn(z, T[11], T[10]);
    ps = __fma_rn(z, ps, T[3]);
    pc = __fma_rn(z, pc, T[9]);
    ps = __fma_rn(z, ps, T[2]);
    pc = __fma_rn(z, pc, T[8]);
    ps = __fma_rn(z, ps, T[1]);
    pc = __fma_rn(z, pc, T[7]);
    ps = __fma_rn(z, ps, T[0]);
    pc = __fma_rn(z, pc, T[6]);
    const double vs = __fma_rn(__dmul_rn(z, r), ps, r);
    const double vc = __fma_rn(__dmul_rn(z, z), pc, __fma_rn(z, -0.5, 1.0));
    double v = (n & 1) ? vc : vs;
    /* negate when n&2: flip the sign bit with integer logic */
    v = __hiloint2double(__double2hiint(v) ^ ((n & 2) << 30), __double2loint(v));
    return (fabs(x) < 1073741824.0) ? v : __longlong_as_double(0x7ff8000000000000LL);
#else
    if (!(fabs(x) < 1073741824.0)) return tk_u2d(0x7ff8000000000000ULL);
    const double t = TK_FMA(x, TK_2OPI, TK_RMAGIC);
    const int n = (int)(uint32_t)(tk_d2u(t) & 0xffffffffu) + shift;
    const double kd = TK_SUB(t, TK_RMAGIC);
    double r = TK_FMA(-kd, TK_PIO2_1, x);
    r = TK_FMA(-kd, TK_PIO2_2, r);
    r = TK_FMA(-kd, TK_PIO2_3, r);
    const double z = TK_MUL(r, r);
    double v;
    if (n & 1) {
        double p = TK_FMA(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
        p = TK_FMA(z, p, -2.75573143513906633035e-07);
        p = TK_FMA(z, p, 2.48015872894767294178e-05);
        p = TK_FMA(z, p, -1.38888888888741095749e-03);
        p = TK_FMA(z, p, 4.16666666666666019037e-02);
        v = TK_FMA(TK_MUL(z, z), p, TK_FMA(z, -0.5, 1.0));
    } else {
        double p = TK_FMA(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
        p = TK_FMA(z, p, 2.75573137070700676789e-06);
        p = TK_FMA(z, p, -1.98412698298579493134e-04);
        p = TK_FMA(z, p, 8.33333333332248946124e-03);
        p = TK_FMA(z, p, -1.66666666666666324348e-01);
        v = TK_FMA(TK_MUL(z, r), p, r);
    }
    return (n & 2) ? -v : v;
#endif
}

TK_HD double tk_cos(double x) { return tk_sincos_core(x, 1); }
TK_HD double tk_sin(double x) { return tk_sincos_core(x, 0); }

/* Round trip of a value through the reference's '200f40.20' text records (sobolFnVal.dat written at
 * TiktakGlobalSearch.f90:451,467 and re-read by chooseSobol :860; searchResults.dat written :594 and
 * re-read by getSortedResults :815): 20 digits after the decimal point.  For |x| >= 2^-14 the nearest
 * 20-decimal string is closer than half an ulp, so the round trip is the identity; below that,
 * N = rint(x*1e20) is an integer < 2^53 and the re-read value is the correctly rounded N / 1e20
 * (1e20 = 2^20 * 5^20 is exactly representable).  Ties (x*1e20 exactly half-integral) round to even,
 * as glibc's printf does.  The oracle does the same with snprintf/strtod (independent path). */
TK_HD double tk_f4020_roundtrip(double x) {
    const double ax = fabs(x);
    if (!(ax < 6.103515625e-05)) return x; /* 2^-14; also passes NaN/Inf through */
    const double hi = TK_MUL(ax, 1e20);      /* < 6.2e15 < 2^53 */
    const double lo = TK_FMA(ax, 1e20, -hi); /* exact product = hi + lo, |lo| <= ulp(hi)/2 */
    double n = rint(hi);                     /* half-even on host (default rounding mode) and device */
    const double d = TK_SUB(hi, n);          /* exact */
    if (d == 0.5 || d == -0.5) {             /* hi is half-integral: the tail lo decides; lo == 0 is a true tie */
        const double fl = floor(hi);
        if (lo > 0.0) n = TK_ADD(fl, 1.0);
        else if (lo < 0.0) n = fl;
    }
    return copysign(TK_DIV(n, 1e20), x);
}

#endif /* TIKTAK_MATH_H */
