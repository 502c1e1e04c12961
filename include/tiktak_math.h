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

#ifndef __CUDACC_RTC__ /* NVRTC (the run-time compiled Sobol kernel, tiktak_b200/csrc/rtc.cu) has these built in */
#include <stdint.h>
#include <string.h>
#include <math.h>
#else
typedef unsigned int uint32_t; typedef int int32_t; typedef unsigned long long uint64_t; typedef long long int64_t;
#endif

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
#define TK_TRIG_DOMAIN 1073741824.0         /* tk_cos/tk_sin are defined for |x| < 2^30 */

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

#if defined(__CUDACC__)
/* coefficients by parity of the quadrant, highest degree first (see tk_cos_n) */
static __constant__ __align__(16) double tk_sel_tab[12] = {
    /* parity 0 (sin): S6..S1 */ 1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    /* parity 1 (cos): C6..C1 */ -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
/* The four constants that meet a second constant inside one FMA (x*2/pi + magic, z*S6 + S5, z*C6 + C5,
 * z*(-0.5) + 1), and the coefficient table of tk_cos_n.  A DFMA takes one constant-bank/uniform operand, so
 * each of these costs two register moves per call unless the value already sits in a register.  tk_trig_plain()
 * returns them as plain constants (ptxas re-materialises them at every use: right for latency-bound code such
 * as the Nelder-Mead kernels, where holding them in registers measured 7 % slower).  A throughput kernel with a
 * hot loop over tk_cos replaces them by values it has LOADED (per lane, from a shared-memory copy): a loaded
 * value is neither re-materialised nor moved to the uniform registers, which saves 8 issue slots per call --
 * same operations, same roundings (kernels_sobol.cu). */
struct tk_trig_regs { double twoopi, s6, c6, mhalf; const double2* sel; /* coefficient table of tk_cos_n: tk_sel_tab or a shared-memory copy */ };
__device__ __forceinline__ tk_trig_regs tk_trig_plain() {
    return tk_trig_regs{tk_sc_tab[12], tk_sc_tab[5], tk_sc_tab[11], -0.5, reinterpret_cast<const double2*>(tk_sel_tab)};
}
__device__ __forceinline__ double tk_sincos_core_dev(double x, int shift, double k_twoopi, double k_s6, double k_c6, double k_mhalf) {
    const double* const T = tk_sc_tab;
    const double t = __fma_rn(x, k_twoopi, TK_RMAGIC);
    const int n = __double2loint(t) + shift;
    const double kd = __dsub_rn(t, TK_RMAGIC);
    double r = __fma_rn(-kd, T[13], x);
    r = __fma_rn(-kd, T[14], r);
    r = __fma_rn(-kd, T[15], r);
    const double z = __dmul_rn(r, r);
    double ps = __fma_rn(z, k_s6, T[4]);
    double pc = __fma_rn(z, k_c6, T[10]);
    ps = __fma_rn(z, ps, T[3]);
    pc = __fma_rn(z, pc, T[9]);
    ps = __fma_rn(z, ps, T[2]);
    pc = __fma_rn(z, pc, T[8]);
    ps = __fma_rn(z, ps, T[1]);
    pc = __fma_rn(z, pc, T[7]);
    ps = __fma_rn(z, ps, T[0]);
    pc = __fma_rn(z, pc, T[6]);
    const double vs = __fma_rn(__dmul_rn(z, r), ps, r);
    const double vc = __fma_rn(__dmul_rn(z, z), pc, __fma_rn(z, k_mhalf, 1.0));
    const double v = (n & 1) ? vc : vs;
    /* sign (n&2) and the domain test |x| < 2^30 are folded into the high word with integer logic: the
     * FP64 pipe is the scarce one.  Outside the domain the result is a NaN (payload unspecified). */
    const int hx = __double2hiint(x) & 0x7fffffff;
    const int nanmask = (hx < 0x41d00000) ? 0 : 0x7ff80000;
    return __hiloint2double((__double2hiint(v) ^ ((n & 2) << 30)) | nanmask, __double2loint(v));
}
#endif

TK_HD double tk_sincos_core(double x, int shift) {
#if defined(__CUDA_ARCH__)
    return tk_sincos_core_dev(x, shift, tk_sc_tab[12], tk_sc_tab[5], tk_sc_tab[11], -0.5);
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
#if defined(__CUDACC__)
/* N independent arguments at once, stage by stage ("breadth first"), and with the branch of the host version:
 * the parity of n picks ONE coefficient set (S or C, read through k.sel with a per-lane index: three 16-byte
 * loads when the kernel has copied tk_sel_tab to shared memory) and one polynomial is evaluated -- the operations of the selected branch of tk_sincos_core,
 * hence the same values, with 14 FP64 instructions per argument instead of 20.  On sm_100a an FP64 instruction
 * holds the issue port for two cycles, so this is the cheaper form wherever the coefficient loads are not the
 * bottleneck; the N dependency chains sit side by side in the instruction stream, so a single warp keeps the
 * pipe fed (~10 registers per element).  CHECK = false leaves out the domain test (|x| < 2^30): for callers
 * that have verified the domain on the host. */
template <int N, bool CHECK>
__device__ __forceinline__ void tk_cos_n(const double (&x)[N], double (&y)[N], const tk_trig_regs& k) {
    const double* const T = tk_sc_tab;
    double t[N], r[N], z[N], p[N], h[N];
    double2 c65[N], c43[N], c21[N];
    int n[N], par[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = __fma_rn(x[i], k.twoopi, TK_RMAGIC);
#pragma unroll
    for (int i = 0; i < N; i++) {
        n[i] = __double2loint(t[i]) + 1;
        par[i] = n[i] & 1;
        const double2* tab = reinterpret_cast<const double2*>(reinterpret_cast<const char*>(k.sel) + 48 * par[i]);
        c65[i] = tab[0]; c43[i] = tab[1]; c21[i] = tab[2];
        t[i] = __dsub_rn(t[i], TK_RMAGIC);
    }
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = __fma_rn(-t[i], T[13], x[i]);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = __fma_rn(-t[i], T[14], r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = __fma_rn(-t[i], T[15], r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) z[i] = __dmul_rn(r[i], r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = __fma_rn(z[i], c65[i].x, c65[i].y);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = __fma_rn(z[i], p[i], c43[i].x);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = __fma_rn(z[i], p[i], c43[i].y);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = __fma_rn(z[i], p[i], c21[i].x);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = __fma_rn(z[i], p[i], c21[i].y);
#pragma unroll
    for (int i = 0; i < N; i++) h[i] = __fma_rn(z[i], k.mhalf, 1.0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const bool odd = par[i] != 0;
        const double m = odd ? z[i] : r[i]; /* cos: (z*z)*C + (1 - z/2);  sin: (z*r)*S + r */
        const double a = odd ? h[i] : r[i];
        const double v = __fma_rn(__dmul_rn(z[i], m), p[i], a);
        int hi = __double2hiint(v) ^ ((n[i] & 2) << 30);
        if (CHECK) {
            const int hx = __double2hiint(x[i]) & 0x7fffffff;
            hi |= (hx < 0x41d00000) ? 0 : 0x7ff80000;
        }
        y[i] = __hiloint2double(hi, __double2loint(v));
    }
}
/* same values, constants from registers (device code that calls them in a loop) */
__device__ __forceinline__ double tk_cos(double x, const tk_trig_regs& k) { return tk_sincos_core_dev(x, 1, k.twoopi, k.s6, k.c6, k.mhalf); }
__device__ __forceinline__ double tk_sin(double x, const tk_trig_regs& k) { return tk_sincos_core_dev(x, 0, k.twoopi, k.s6, k.c6, k.mhalf); }
#endif

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
