/* tiktak_obj.cuh -- the device-side objective plugin contract of libtiktak_cuda.
 *
 * Reference contract being replaced: MODULE OBJECTIVE (objective.f90:16-144), which exports
 *   objFun(theta), dfovec(np,nm,x,Fnout), getPenalty(theta), obj_initialize, GetData, diagnostic
 * and reads p_nx, p_nmom, p_bound, p_fvalmax from genericParams (objective.f90:17).  As in the
 * reference, the objective is COMPILED IN: a user edits one plugin struct and rebuilds the library.
 *
 * A plugin is a struct with
 *     static constexpr int kId;                          // TIKTAK_OBJ_* id it registers under
 *     static void host_init(int nx, double* aux);        // obj_initialize/GetData: per-coordinate
 *                                                        // constants aux[nx], computed once on the host
 *     static __host__ __device__ int nterms(int nx);     // how many per-coordinate terms (<= nx)
 *     template <class X>
 *     static __device__ void   term(const tk_obj_ctx&, const X& x, int i, double& a, double& b);
 *     static __device__ void   fold_init(const tk_obj_ctx&, double& s, double& p);
 *     static __device__ void   fold(double& s, double& p, double a, double b);
 *     static __device__ double finish(const tk_obj_ctx&, double s, double p);
 *     static constexpr bool kTermLocal;                  // true when term(.,i,.) reads x[i] only
 *     template <int G, bool TRUSTED, class X>            // optional: G consecutive terms at once
 *     static __device__ void   terms(const tk_obj_ctx&, const X& x, int i0, double (&a)[G], double (&b)[G]);
 *     static bool domain_ok(int nx, const double* lo, const double* hi, const double* aux);   // optional, host
 * Together they are the user part of dfovec (objective.f90:96-117), i.e. the scalar v_err that the
 * shipped template replicates over the nm moments:
 *     fold_init(s,p); for i = 0 .. nterms-1 { term(x,i,a,b); fold(s,p,a,b); }  v_err = finish(s,p)
 * `term` holds the expensive independent work of coordinate i (a cos, a square ...); `fold` is the
 * cheap accumulation, ALWAYS applied in index order.  The split lets every stage pick its own
 * parallelism without changing a single rounding: the Sobol kernel runs the loop in one thread with
 * x in registers; the Nelder-Mead kernel spreads the `term`s of its trial points over the lanes of a
 * warp and folds them in order on one lane -- so all stages, and the CPU oracle, see bit-identical
 * objective values.  X is any type with `double operator[](int) const`.  Everything around this
 * (getPenalty, the penalty branch of dfovec, DOT_PRODUCT, objFun) is framework code below,
 * restating objective.f90:51-74,92-95,125-142 operation for operation.  A plugin that cannot be
 * decomposed sets nterms = 1 and does everything in term(.,0,.).
 *
 * Arithmetic rules for plugin authors: use TK_MUL/TK_ADD/... or plain operators (the library is
 * compiled with -fmad=false, so nothing is contracted behind your back), call tk_cos/tk_sin from
 * tiktak_math.h instead of cos/sin if you need CPU/GPU trajectories to agree bit for bit (the
 * two-argument form tk_cos(x, c.trig) is the same function with its constants kept in registers), and return
 * a finite value >= fvalmax where the model is undefined (README.md:34) -- never NaN.
 *
 * Objectives whose dfovec fills a genuine VECTOR of moments (TIKTAK_OBJ_SMM) use the block-wide plugin form instead
 * (tiktak_b200/csrc/smm.cuh: kVector = true, `dfovec_block(user, theta, F, scratch)` computed by all threads of a block,
 * `smem_doubles`): DFNLS consumes the residuals per moment, Nelder-Mead and the Sobol stage take objFun = DOT_PRODUCT(F, F)
 * in index order + getPenalty from the same evaluation (kernels_dfnls.cu: dfnls_kernel, nm_vec_kernel; kernels_smm.cu).
 */
#ifndef TIKTAK_OBJ_CUH
#define TIKTAK_OBJ_CUH

#include "tiktak_math.h"

#define TK_NX_PARAM_MAX 64 /* largest nx whose tables travel as __grid_constant__ kernel parameters */

/* what objective code may read: the module variables of genericParams + the penalty constants of
 * obj_initialize (objective.f90:35-36).  Pointers are valid in the address space of the caller. */
struct tk_obj_ctx {
    int nx, nm;
    const double* bl;  /* p_bound(:,1) */
    const double* bu;  /* p_bound(:,2) */
    const double* aux; /* plugin constants from host_init */
    double fvalmax, penalty0, penaltyc, max_penalty;
    double sqrt_nm;    /* DSQRT(DBLE(nm)), objective.f90:94 */
    const void* user;
#ifdef __CUDACC__
    tk_trig_regs trig; /* register-resident constants of tk_cos/tk_sin: call tk_cos(x, c.trig) */
#endif
};

#define TK_MYEPS 2.220446049250313e-16 /* EPSILON(1.0_DP), objective.f90:24 */

#ifdef __CUDACC__

/* getPenalty (objective.f90:51-74).  When every coordinate is inside its bounds each term is an
 * exact +0, so the sum is skipped; otherwise the full sum is formed in index order. */
template <class X>
__device__ __forceinline__ double tk_get_penalty(const tk_obj_ctx& c, const X& x) {
    bool out = false;
#pragma unroll
    for (int i = 0; i < c.nx; i++) out = out || (x[i] < c.bl[i]) || (x[i] > c.bu[i]);
    if (!out) return 0.0;
    double penalty = 0.0;
    for (int i = 0; i < c.nx; i++) {
        const double xi = x[i];
        const double a = TK_DIV(fmax(0.0, TK_SUB(c.bl[i], xi)), fmax(0.1, fabs(c.bl[i])));
        const double b = TK_DIV(fmax(0.0, TK_SUB(xi, c.bu[i])), fmax(0.1, fabs(c.bu[i])));
        const double temp = TK_ADD(TK_MUL(c.penaltyc, TK_MUL(a, a)), TK_MUL(c.penaltyc, TK_MUL(b, b)));
        penalty = TK_ADD(penalty, temp);
    }
    if (penalty > TK_MYEPS) return fmin(c.max_penalty, TK_ADD(c.penalty0, penalty));
    return 0.0;
}

/* bounds test only: true when some coordinate is outside [bl, bu] */
template <class X>
__device__ __forceinline__ bool tk_out_of_bounds(const tk_obj_ctx& c, const X& x) {
    bool out = false;
#pragma unroll
    for (int i = 0; i < c.nx; i++) out = out || (x[i] < c.bl[i]) || (x[i] > c.bu[i]);
    return out;
}

/* the user part of dfovec, evaluated sequentially by one thread */
template <class Obj, class X>
__device__ __forceinline__ double tk_value_seq(const tk_obj_ctx& c, const X& x) {
    double s, p;
    Obj::fold_init(c, s, p);
    const int nt = Obj::nterms(c.nx);
#pragma unroll
    for (int i = 0; i < nt; i++) {
        double a, b;
        Obj::term(c, x, i, a, b);
        Obj::fold(s, p, a, b);
    }
    return Obj::finish(c, s, p);
}

/* The same value with the terms taken G at a time through the plugin's optional batched form
 *     template <int G, bool TRUSTED, class X> static __device__ void terms(ctx, x, i0, double (&a)[G], double (&b)[G]);
 * (terms i0 .. i0+G-1, each computed by the operations of `term`); the fold stays in index order, so
 * nothing is rounded differently.  NT = nterms, known at compile time; G must divide NT.  TRUSTED: the caller
 * has checked on the host, through the plugin's optional
 *     static bool domain_ok(int nx, const double* lo, const double* hi, const double* aux);
 * that every tk_cos/tk_sin argument stays inside the functions' domain for all x in the box [lo, hi], so the
 * per-call domain test may be left out (tk_cos_n<N, false>). */
template <class Obj, int NT, int G, bool TRUSTED, class X>
__device__ __forceinline__ double tk_value_grouped(const tk_obj_ctx& c, const X& x) {
    static_assert(NT % G == 0, "group size must divide the number of terms");
    double s, p;
    Obj::fold_init(c, s, p);
#pragma unroll
    for (int i0 = 0; i0 < NT; i0 += G) {
        double a[G], b[G];
        Obj::template terms<G, TRUSTED>(c, x, i0, a, b);
#pragma unroll
        for (int g = 0; g < G; g++) Obj::fold(s, p, a[g], b[g]);
    }
    return Obj::finish(c, s, p);
}

/* objFun (objective.f90:125-142) given dfovec's common moment value f and getPenalty:
 * DOT_PRODUCT over the nm identical moments in index order, then + getPenalty. */
__device__ __forceinline__ double tk_objfun_finish(const tk_obj_ctx& c, double f, double pen) {
    const double ff = TK_MUL(f, f);
    double dot = ff; /* 0 + ff is exact, so the first term needs no addition */
    for (int m = 1; m < c.nm; m++) dot = TK_ADD(dot, ff);
    return TK_ADD(dot, pen);
}

/* objFun for plugins whose moments are nm copies of one value (the shipped template) */
template <class Obj, class X>
__device__ __forceinline__ double tk_objfun(const tk_obj_ctx& c, const X& x) {
    const double pen = tk_get_penalty(c, x);
    double f;
    if (pen > TK_MYEPS) f = TK_DIV(pen, c.sqrt_nm); /* objective.f90:93-94 */
    else f = tk_value_seq<Obj>(c, x);
    return tk_objfun_finish(c, f, pen);
}

#endif /* __CUDACC__ */
#endif /* TIKTAK_OBJ_CUH */
