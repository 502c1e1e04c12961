/* objectives.cuh -- built-in objective plugins (contract: include/tiktak_obj.cuh).
 *
 * ObjTemplate restates the user section of the shipped dfovec (objective.f90:96-112) literally,
 * `call sleep(3)` (objective.f90:113) removed.  The other three are the synthetic benchmark
 * functions named in BASELINE.json, written in the same style: plain unfused arithmetic in index
 * order, v_err = f(x) + 1 "for relative criterion" (objective.f90:111).
 */
#pragma once
#include <math.h>

#include "../../include/tiktak_cuda.h"
#include "../../include/tiktak_obj.cuh"

struct ObjTemplate {
    static constexpr int kId = TIKTAK_OBJ_TEMPLATE;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = sqrt((double)(i + 1) + 0.0); /* sqrt(ii+0.0D0) */
    }
    template <class X>
    static __device__ __forceinline__ double value(const tk_obj_ctx& c, const X& x) {
        double val1 = TK_DIV(TK_MUL(x[0], x[0]), 200.0);
        double val2 = tk_cos(x[0]);
#pragma unroll
        for (int i = 1; i < c.nx; i++) {
            val1 = TK_ADD(val1, TK_DIV(TK_MUL(x[i], x[i]), 200.0));
            val2 = TK_MUL(val2, tk_cos(TK_DIV(x[i], c.aux[i])));
        }
        const double v = TK_ADD(TK_SUB(val1, val2), 1.0);
        return TK_ADD(v, 1.0);
    }
};

struct ObjGriewank {
    static constexpr int kId = TIKTAK_OBJ_GRIEWANK;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = 1.0 / sqrt((double)(i + 1));
    }
    template <class X>
    static __device__ __forceinline__ double value(const tk_obj_ctx& c, const X& x) {
        double s = 0.0, p = 1.0;
#pragma unroll
        for (int i = 0; i < c.nx; i++) {
            s = TK_ADD(s, TK_MUL(TK_MUL(x[i], x[i]), 2.5e-4));
            p = TK_MUL(p, tk_cos(TK_MUL(x[i], c.aux[i])));
        }
        const double v = TK_ADD(TK_SUB(s, p), 1.0);
        return TK_ADD(v, 1.0);
    }
};

struct ObjRastrigin {
    static constexpr int kId = TIKTAK_OBJ_RASTRIGIN;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = 0.0;
    }
    template <class X>
    static __device__ __forceinline__ double value(const tk_obj_ctx& c, const X& x) {
        double s = TK_MUL(10.0, (double)c.nx);
#pragma unroll
        for (int i = 0; i < c.nx; i++) {
            const double t = TK_SUB(TK_MUL(x[i], x[i]), TK_MUL(10.0, tk_cos(TK_MUL(6.283185307179586, x[i]))));
            s = TK_ADD(s, t);
        }
        return TK_ADD(s, 1.0);
    }
};

struct ObjRosenbrock {
    static constexpr int kId = TIKTAK_OBJ_ROSENBROCK;
    static void host_init(int nx, double* aux) {
        for (int i = 0; i < nx; i++) aux[i] = 0.0;
    }
    template <class X>
    static __device__ __forceinline__ double value(const tk_obj_ctx& c, const X& x) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i + 1 < c.nx; i++) {
            const double t = TK_SUB(x[i + 1], TK_MUL(x[i], x[i]));
            const double u = TK_SUB(1.0, x[i]);
            s = TK_ADD(s, TK_ADD(TK_MUL(100.0, TK_MUL(t, t)), TK_MUL(u, u)));
        }
        return TK_ADD(s, 1.0);
    }
};
