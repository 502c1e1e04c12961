/* objectives.cuh -- built-in objective plugins (contract: include/tiktak_obj.cuh).
 *
 * ObjTemplate restates the user section of the shipped dfovec (objective.f90:96-112) literally,
 * `call sleep(3)` (objective.f90:113) removed.  The other three are the synthetic benchmark
 * functions named in BASELINE.json, written in the same style: plain unfused arithmetic in index
 * order, v_err = f(x) + 1 "for relative criterion" (objective.f90:111).
 */
#pragma once
#ifndef __CUDACC_RTC__ /* (the run-time compiled Sobol kernel gets these files pasted in front of this one, rtc.cu) */
#include <math.h>

#include "../../include/tiktak_cuda.h"
#include "../../include/tiktak_obj.cuh"
#endif

struct ObjTemplate { /* val1 = sum x^2/200 ; val2 = cos(x1) * prod cos(x_i/sqrt(i)) ; v = val1 - val2 + 1 + 1 */
    static constexpr int kId = TIKTAK_OBJ_TEMPLATE;
    static constexpr bool kVector = false;
    static constexpr bool kTermLocal = true;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = sqrt((double)(i + 1) + 0.0); /* sqrt(ii+0.0D0), :106 */
    }
    static bool domain_ok(int nx, const double* lo, const double* hi, const double*) { /* |x_i / sqrt(i)| <= |x_i| */
        for (int i = 0; i < nx; i++) if (!(fmax(fabs(lo[i]), fabs(hi[i])) < TK_TRIG_DOMAIN * 0.5)) return false;
        return true;
    }
    static constexpr __host__ __device__ __forceinline__ int nterms(int nx) { return nx; }
    template <class X>
    static __device__ __forceinline__ void term(const tk_obj_ctx& c, const X& x, int i, double& a, double& b) {
        const double xi = x[i];
        a = TK_DIV(TK_MUL(xi, xi), 200.0);
        b = (i == 0) ? tk_cos(xi, c.trig) : tk_cos(TK_DIV(xi, c.aux[i]), c.trig); /* x/sqrt(1) == x exactly anyway */
    }
    template <int G, bool TRUSTED, class X>
    static __device__ __forceinline__ void terms(const tk_obj_ctx& c, const X& x, int i0, double (&a)[G], double (&b)[G]) {
        double arg[G];
#pragma unroll
        for (int g = 0; g < G; g++) arg[g] = (i0 + g == 0) ? x[0] : TK_DIV(x[i0 + g], c.aux[i0 + g]);
        tk_cos_n<G, !TRUSTED>(arg, b, c.trig);
#pragma unroll
        for (int g = 0; g < G; g++) a[g] = TK_DIV(TK_MUL(x[i0 + g], x[i0 + g]), 200.0);
    }
    static __device__ __forceinline__ void fold_init(const tk_obj_ctx&, double& s, double& p) { s = 0.0; p = 1.0; }
    static __device__ __forceinline__ void fold(double& s, double& p, double a, double b) {
        s = TK_ADD(s, a); /* 0 + a0 and 1 * b0 are exact, so this equals the reference's seeded loops */
        p = TK_MUL(p, b);
    }
    static __device__ __forceinline__ double finish(const tk_obj_ctx&, double s, double p) {
        return TK_ADD(TK_ADD(TK_SUB(s, p), 1.0), 1.0);
    }
};

struct ObjGriewank { /* f = 1 + sum x^2/4000 - prod cos(x_i/sqrt(i)) */
    static constexpr int kId = TIKTAK_OBJ_GRIEWANK;
    static constexpr bool kVector = false;
    static constexpr bool kTermLocal = true;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = 1.0 / sqrt((double)(i + 1));
    }
    static bool domain_ok(int nx, const double* lo, const double* hi, const double*) { /* aux <= 1 */
        for (int i = 0; i < nx; i++) if (!(fmax(fabs(lo[i]), fabs(hi[i])) < TK_TRIG_DOMAIN * 0.5)) return false;
        return true;
    }
    static constexpr __host__ __device__ __forceinline__ int nterms(int nx) { return nx; }
    template <class X>
    static __device__ __forceinline__ void term(const tk_obj_ctx& c, const X& x, int i, double& a, double& b) {
        const double xi = x[i];
        a = TK_MUL(TK_MUL(xi, xi), 2.5e-4);
        b = tk_cos(TK_MUL(xi, c.aux[i]), c.trig);
    }
    template <int G, bool TRUSTED, class X>
    static __device__ __forceinline__ void terms(const tk_obj_ctx& c, const X& x, int i0, double (&a)[G], double (&b)[G]) {
        double arg[G];
#pragma unroll
        for (int g = 0; g < G; g++) arg[g] = TK_MUL(x[i0 + g], c.aux[i0 + g]);
        tk_cos_n<G, !TRUSTED>(arg, b, c.trig);
#pragma unroll
        for (int g = 0; g < G; g++) a[g] = TK_MUL(TK_MUL(x[i0 + g], x[i0 + g]), 2.5e-4);
    }
    static __device__ __forceinline__ void fold_init(const tk_obj_ctx&, double& s, double& p) { s = 0.0; p = 1.0; }
    static __device__ __forceinline__ void fold(double& s, double& p, double a, double b) {
        s = TK_ADD(s, a);
        p = TK_MUL(p, b);
    }
    static __device__ __forceinline__ double finish(const tk_obj_ctx&, double s, double p) {
        return TK_ADD(TK_ADD(TK_SUB(s, p), 1.0), 1.0);
    }
};

struct ObjRastrigin { /* f = 10 n + sum (x^2 - 10 cos(2 pi x)) */
    static constexpr int kId = TIKTAK_OBJ_RASTRIGIN;
    static constexpr bool kVector = false;
    static constexpr bool kTermLocal = true;
    static constexpr bool kFoldUsesB = false;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = 0.0;
    }
    static bool domain_ok(int nx, const double* lo, const double* hi, const double*) { /* |2 pi x| < 8 |x| */
        for (int i = 0; i < nx; i++) if (!(fmax(fabs(lo[i]), fabs(hi[i])) < TK_TRIG_DOMAIN * 0.125)) return false;
        return true;
    }
    static constexpr __host__ __device__ __forceinline__ int nterms(int nx) { return nx; }
    template <class X>
    static __device__ __forceinline__ void term(const tk_obj_ctx& c, const X& x, int i, double& a, double& b) {
        const double xi = x[i];
        a = TK_SUB(TK_MUL(xi, xi), TK_MUL(10.0, tk_cos(TK_MUL(6.283185307179586, xi), c.trig)));
        b = 0.0;
    }
    template <int G, bool TRUSTED, class X>
    static __device__ __forceinline__ void terms(const tk_obj_ctx& c, const X& x, int i0, double (&a)[G], double (&b)[G]) {
        double arg[G], cs[G];
#pragma unroll
        for (int g = 0; g < G; g++) arg[g] = TK_MUL(6.283185307179586, x[i0 + g]);
        tk_cos_n<G, !TRUSTED>(arg, cs, c.trig);
#pragma unroll
        for (int g = 0; g < G; g++) { a[g] = TK_SUB(TK_MUL(x[i0 + g], x[i0 + g]), TK_MUL(10.0, cs[g])); b[g] = 0.0; }
    }
    static __device__ __forceinline__ void fold_init(const tk_obj_ctx& c, double& s, double& p) {
        s = TK_MUL(10.0, (double)c.nx);
        p = 0.0;
    }
    static __device__ __forceinline__ void fold(double& s, double&, double a, double) { s = TK_ADD(s, a); }
    static __device__ __forceinline__ double finish(const tk_obj_ctx&, double s, double) { return TK_ADD(s, 1.0); }
};

struct ObjRosenbrock { /* f = sum_{i<n} 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2 */
    static constexpr int kId = TIKTAK_OBJ_ROSENBROCK;
    static constexpr bool kVector = false;
    static constexpr bool kTermLocal = false;
    static constexpr bool kFoldUsesB = false;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = 0.0;
    }
    static constexpr __host__ __device__ __forceinline__ int nterms(int nx) { return nx - 1; }
    template <class X>
    static __device__ __forceinline__ void term(const tk_obj_ctx&, const X& x, int i, double& a, double& b) {
        const double x0 = x[i];
        const double t = TK_SUB(x[i + 1], TK_MUL(x0, x0));
        const double u = TK_SUB(1.0, x0);
        a = TK_ADD(TK_MUL(100.0, TK_MUL(t, t)), TK_MUL(u, u));
        b = 0.0;
    }
    static __device__ __forceinline__ void fold_init(const tk_obj_ctx&, double& s, double& p) { s = 0.0; p = 0.0; }
    static __device__ __forceinline__ void fold(double& s, double&, double a, double) { s = TK_ADD(s, a); }
    static __device__ __forceinline__ double finish(const tk_obj_ctx&, double s, double) { return TK_ADD(s, 1.0); }
};
