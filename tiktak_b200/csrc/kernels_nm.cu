/* kernels_nm.cu -- the restart stage on the device: theta-blend, batched warm-start Nelder-Mead, result table.
 *
 * Replaces, per restart k: getBestLine/getModifiedParam (TiktakGlobalSearch.f90:683-720, 781-823: the
 * theta-blend toward the best local minimum so far -- or toward a rank-lottery row, getLotteryPoint :722-766 --
 * which the reference obtains by re-reading and re-sorting searchResults.dat), runAmoeba (:602-633: regular
 * simplex from simplex_coordinates2 scaled by range/K^(1/nx), clamped to the bounds, nx+1 evaluations),
 * amoeba/amotry (minimize.f90:7-135) and the bookkeeping of completeSearch (:582-597).
 *
 * Three Nelder-Mead kernels, same arithmetic and decisions, chosen by nx:
 *   nm_quad_kernel       nx <= 31   one restart per block of four warps; every warp carries the simplex state
 *                                   (sorted across lanes) and evaluates one of the four trial points
 *   nm_quad_wide_kernel  nx <= 127  one master warp keeps the state, all four warps evaluate
 *   nm_kernel            fallback   one restart per warp, state in the warp's shared memory
 * The sequential part of amoeba -- reflect, then expand OR contract -- is collapsed into one evaluation phase per
 * iteration by evaluating the four possible trial points (reflection R, expansion from R, contraction after an
 * accepted R, contraction after a rejected R) at once; which of them count, and the `iter` bookkeeping, follow
 * minimize.f90:97-114 exactly, so the trajectory is the sequential one.  All arithmetic is unfused, in the
 * reference's order; psum is updated incrementally (minimize.f90:130) and only re-summed after a shrink (:112).
 * ingest_kernel / blend_kernel / shrink_width_kernel / pack_kernel keep searchResults on the device between rounds.
 */
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "nm_common.cuh"
#include "../../include/tiktak_rng.h"

using namespace tknm;

namespace {

__device__ __forceinline__ void warp_argmin(const double* y, int np, int lane, int& idx) {
    double v = DBL_MAX; int bi = 0x7fffffff;
    bool has = false;
    for (int i = lane; i < np; i += 32) {
        const double yi = y[i];
        if (!has || yi < v) { v = yi; bi = i; has = true; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (i2 != 0x7fffffff && (bi == 0x7fffffff || v2 < v || (v2 == v && i2 < bi))) { v = v2; bi = i2; }
    }
    idx = bi;
}
/* argmax with y[skip] read as `subst` (amoeba's y(ihi)=y(ilo) trick, minimize.f90:80-83); skip<0: none */
__device__ __forceinline__ void warp_argmax(const double* y, int np, int lane, int skip, double subst, int& idx) {
    double v = -DBL_MAX; int bi = 0x7fffffff;
    bool has = false;
    for (int i = lane; i < np; i += 32) {
        const double yi = (i == skip) ? subst : y[i];
        if (!has || yi > v) { v = yi; bi = i; has = true; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (i2 != 0x7fffffff && (bi == 0x7fffffff || v2 > v || (v2 == v && i2 < bi))) { v = v2; bi = i2; }
    }
    idx = bi;
}

template <class Obj>
__global__ void __launch_bounds__(128) nm_kernel(const TkTablesG T, const TkScalars sc, const NmArgs a) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int nx = sc.nx, np = nx + 1;
    const int nt = Obj::nterms(nx);
    /* block-shared copies of the bounds and plugin constants */
    double* sbl = smem;
    double* sbu = sbl + nx;
    double* saux = sbu + nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    __syncthreads();
    const int r = blockIdx.x * wpb + warp;
    if (r >= a.n_run) return;
    const int slot = a.rank + r * a.world;
    const int per_warp = np * nx + np + nx + 4 * nx + 4 + 8 * nx;
    double* p = smem + 3 * nx + (size_t)warp * per_warp; /* p[i*nx + j] */
    double* y = p + np * nx;
    double* psum = y + np;
    double* cand = psum + nx; /* cand[c*nx + j], c = R, E, Cout, Cin */
    double* ycand = cand + 4 * nx;
    double* ta = ycand + 4;   /* per-term partials of the four trial points */
    double* tb = ta + 4 * nx;
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, sbl, sbu, saux);

    /* runAmoeba :622-627 -- initial simplex around the blended start */
    const bool shr = a.shr_have && a.shr_have[0] && ((float)(a.first_k + slot) / (float)a.K > 0.5f);
    const double* wsrc = shr ? a.wshr : a.width;
    for (int e = lane; e < np * nx; e += 32) {
        const int ia = e / nx, i = e - ia * nx;
        double t = TK_ADD(a.starts[(size_t)i * a.n_k + slot], TK_DIV(TK_MUL(a.simplex[(size_t)ia * nx + i], wsrc[i]), a.scale));
        t = fmax(fmin(t, T.bu[i]), T.bl[i]);
        p[e] = t;
    }
    __syncwarp();
    for (int ia = lane; ia < np; ia += 32) y[ia] = tk_objfun<Obj>(ctx, TkMemX{p + (size_t)ia * nx, 1});
    __syncwarp();
    /* amoeba :49 psum = sum(p, dim=1), index order */
    for (int j = lane; j < nx; j += 32) {
        double s = 0.0;
        for (int i = 0; i < np; i++) s = TK_ADD(s, p[i * nx + j]);
        psum[j] = s;
    }
    __syncwarp();

    int iter = 0, ilo = 0;
    for (;;) {
        int ihi, inhi;
        warp_argmin(y, np, lane, ilo);
        warp_argmax(y, np, lane, -1, 0.0, ihi);
        const double ylo = y[ilo], yhi = y[ihi];
        warp_argmax(y, np, lane, ihi, ylo, inhi);
        const double ynhi = y[inhi];
        const double rtol = TK_DIV(TK_MUL(2.0, fabs(TK_SUB(yhi, ylo))), TK_ADD(TK_ADD(fabs(yhi), fabs(ylo)), a.tiny));
        if (rtol < a.ftol || iter >= a.itmax) break; /* :85-96 */

        /* the four possible trial points of this iteration (amotry :124-126, unfused) */
        for (int j = lane; j < nx; j += 32) {
            const double ps = psum[j], ph = p[ihi * nx + j];
            const double R = TK_SUB(TK_MUL(ps, a.fac1r), TK_MUL(ph, a.fac2r));
            const double ps1 = TK_ADD(TK_SUB(ps, ph), R); /* psum after R is accepted (:130) */
            cand[j] = R;
            cand[nx + j] = TK_SUB(TK_MUL(ps1, a.fac1e), TK_MUL(R, a.fac2e));
            cand[2 * nx + j] = TK_SUB(TK_MUL(ps1, a.fac1c), TK_MUL(R, a.fac2c));
            cand[3 * nx + j] = TK_SUB(TK_MUL(ps, a.fac1c), TK_MUL(ph, a.fac2c));
        }
        __syncwarp();
        /* objFun at the four trial points: the independent per-coordinate `term`s are spread over the
         * lanes, the in-order `fold` runs on one lane per point (tiktak_obj.cuh) -- same roundings as a
         * single-thread evaluation.  A point outside the bounds takes the full getPenalty path. */
        unsigned oob = 0;
        for (int e = lane; e < 4 * nx; e += 32) {
            const int c = e / nx, i = e - c * nx;
            const double xv = cand[e];
            if (xv < sbl[i] || xv > sbu[i]) oob |= 1u << c;
        }
        oob = __reduce_or_sync(0xffffffffu, oob);
        for (int e = lane; e < 4 * nt; e += 32) {
            const int c = e / nt, i = e - c * nt;
            if ((oob >> c) & 1u) continue;
            double ta_, tb_;
            Obj::term(ctx, TkMemX{cand + (size_t)c * nx, 1}, i, ta_, tb_);
            ta[c * nx + i] = ta_;
            tb[c * nx + i] = tb_;
        }
        __syncwarp();
        if (lane < 4) {
            double yv;
            if ((oob >> lane) & 1u) {
                yv = tk_objfun<Obj>(ctx, TkMemX{cand + (size_t)lane * nx, 1});
            } else {
                double s_, p_;
                Obj::fold_init(ctx, s_, p_);
                for (int i = 0; i < nt; i++) Obj::fold(s_, p_, ta[lane * nx + i], tb[lane * nx + i]);
                yv = tk_objfun_finish(ctx, Obj::finish(ctx, s_, p_), 0.0);
            }
            ycand[lane] = yv;
        }
        __syncwarp();
        const double yR = ycand[0];
        const bool accR = yR < yhi; /* :128 */
        if (accR) {
            for (int j = lane; j < nx; j += 32) {
                const double R = cand[j];
                psum[j] = TK_ADD(TK_SUB(psum[j], p[ihi * nx + j]), R);
                p[ihi * nx + j] = R;
            }
            if (lane == 0) y[ihi] = yR;
        }
        iter++;
        __syncwarp();
        if (yR <= ylo) { /* :99 expansion */
            double yE = ycand[1];
            if (!accR) { /* only if every y was equal; not speculated -- evaluate from the live state */
                for (int j = lane; j < nx; j += 32)
                    cand[nx + j] = TK_SUB(TK_MUL(psum[j], a.fac1e), TK_MUL(p[ihi * nx + j], a.fac2e));
                __syncwarp();
                if (lane == 0) ycand[1] = tk_objfun<Obj>(ctx, TkMemX{cand + nx, 1});
                __syncwarp();
                yE = ycand[1];
            }
            iter++;
            if (yE < y[ihi]) {
                __syncwarp();
                for (int j = lane; j < nx; j += 32) {
                    const double E = cand[nx + j];
                    psum[j] = TK_ADD(TK_SUB(psum[j], p[ihi * nx + j]), E);
                    p[ihi * nx + j] = E;
                }
                if (lane == 0) y[ihi] = yE;
            }
            __syncwarp();
        } else if (yR >= ynhi) { /* :102 contraction */
            const double ysave = accR ? yR : yhi; /* y(ihi) at :103 */
            const int cc = accR ? 2 : 3;
            const double yC = ycand[cc];
            iter++;
            if (yC < ysave) {
                for (int j = lane; j < nx; j += 32) {
                    const double C = cand[cc * nx + j];
                    psum[j] = TK_ADD(TK_SUB(psum[j], p[ihi * nx + j]), C);
                    p[ihi * nx + j] = C;
                }
                if (lane == 0) y[ihi] = yC;
                __syncwarp();
            } else { /* :106-113 shrink toward the best vertex, re-evaluate, re-sum */
                for (int e = lane; e < np * nx; e += 32) {
                    const int i = e / nx, j = e - i * nx;
                    if (i != ilo) p[e] = TK_MUL(0.5, TK_ADD(p[e], p[ilo * nx + j]));
                }
                __syncwarp();
                for (int i = lane; i < np; i += 32)
                    if (i != ilo) y[i] = tk_objfun<Obj>(ctx, TkMemX{p + (size_t)i * nx, 1});
                iter += nx;
                __syncwarp();
                for (int j = lane; j < nx; j += 32) {
                    double s = 0.0;
                    for (int i = 0; i < np; i++) s = TK_ADD(s, p[i * nx + j]);
                    psum[j] = s;
                }
                __syncwarp();
            }
        }
    }
    /* runAmoeba :631-632 returns the first minimal vertex = p(ilo) (amoeba swapped it into row 1);
     * completeSearch :582 re-evaluates objFun there: same point, same value. */
    for (int j = lane; j < nx; j += 32) a.out_x[(size_t)j * a.n_k + slot] = p[ilo * nx + j];
    if (lane == 0) {
        a.out_f[slot] = y[ilo];
        a.out_evals[slot] = np + iter + 1;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Variant for nx <= 31 and nterms <= nx: ONE RESTART PER BLOCK OF FOUR WARPS.
 *
 * amoeba's iteration is a chain of dependent steps (worst vertex -> reflection -> objFun -> expansion OR
 * contraction -> objFun -> ...), so a restart is latency-bound; the kernel shortens the chain instead of
 * widening it:
 *   - the four trial points an iteration can need (reflection R, expansion from R, contraction after an
 *     accepted R, contraction after a rejected R) are evaluated AT ONCE, one per warp (one warp per SM
 *     sub-partition): lane j computes `term` j, lane 0 folds the terms in index order -- the roundings of a
 *     single-thread evaluation (tiktak_obj.cuh) -- and the four values meet in shared memory behind the one
 *     block barrier of the iteration.  Which of them count, and the `iter` bookkeeping, follow
 *     minimize.f90:97-114 exactly, so the trajectory is the sequential one.
 *   - every warp carries the full simplex state redundantly (own copy of p in shared memory, psum_j in lane
 *     j), so nothing else is exchanged and all four warps take the same decisions.
 *   - the vertices are kept SORTED by (y ascending, vertex index descending) across lanes 0..nx: yhi and
 *     ihi = MAXLOC's first occurrence are lane nx, ynhi (y(inhi) after the y(ihi)=y(ilo) substitution,
 *     :80-83) is lane nx-1, ylo is lane 0 -- no reductions; replacing the worst vertex is one ballot +
 *     shuffle-up insertion.  MINLOC's first occurrence (lowest vertex index among equal minima) is only
 *     needed by the shrink and at the end and is resolved there.
 *   - out-of-bounds trial points (rare) take the literal getPenalty/dfovec path on one lane.
 * NXC > 0 compiles nx in (unrolled folds, constant index arithmetic); NXC = 0 reads nx at run time.
 * --------------------------------------------------------------------------------------------- */
struct TkLaneX { /* accessor for plugins whose term i reads x[i] only: the lane's own coordinate */
    double v;
    __device__ __forceinline__ double operator[](int) const { return v; }
};

/* sort the np (y, vertex) pairs held one per lane; every lane returns the pair of its sorted position */
__device__ __forceinline__ void nm_full_sort(double& y, int& row, int np, int lane, double* scr_y, int* scr_r) {
    int rank = 0;
    for (int k = 0; k < np; k++) {
        const double yk = __shfl_sync(0xffffffffu, y, k);
        const int rk = __shfl_sync(0xffffffffu, row, k);
        rank += (k != lane && nm_before(yk, rk, y, row)) ? 1 : 0;
    }
    __syncwarp();
    if (lane < np) { scr_y[rank] = y; scr_r[rank] = row; }
    __syncwarp();
    if (lane < np) { y = scr_y[lane]; row = scr_r[lane]; }
    __syncwarp();
}

/* per-warp scratch of the warp-cooperative objFun */
struct NmScratch {
    double* ta;   /* [ts] per-term partials a_i */
    double* tb;   /* [ts] per-term partials b_i */
    double* pn;   /* [ts] per-coordinate penalty terms (out-of-bounds points only) */
    double* cand; /* [ts] the point, for plugins whose terms read neighbouring coordinates */
};

/* objFun (objective.f90:125-142) at the point whose coordinate j sits in lane j (xj), by one warp; the value
 * is returned on lane 0.  Lane i evaluates `term` i, lane 0 folds them in index order: the roundings of a
 * single-thread evaluation.  A point outside the bounds takes the literal getPenalty sum (:51-74: per-
 * coordinate terms in parallel, summed in index order on lane 0) and dfovec's penalty branch (:92-95). */
template <class Obj, int NXC>
__device__ __forceinline__ double nm_eval_warp(const tk_obj_ctx& ctx, const NmScratch& sc, double xj, double blj, double buj,
                                               int lane, int nx, int nt) {
    const bool has_j = lane < nx, has_t = lane < nt;
    const bool oob = __any_sync(0xffffffffu, has_j && (xj < blj || xj > buj));
    double ta_ = 0.0, tb_ = 0.0;
    if (Obj::kTermLocal) {
        if (has_t) Obj::term(ctx, TkLaneX{xj}, lane, ta_, tb_);
    } else {
        __syncwarp();
        if (has_j) sc.cand[lane] = xj;
        __syncwarp();
        if (has_t) Obj::term(ctx, TkMemX{sc.cand, 1}, lane, ta_, tb_);
    }
    if (has_t) { sc.ta[lane] = ta_; sc.tb[lane] = tb_; }
    if (oob && has_j) { /* getPenalty :57-66, coordinate `lane` */
        const double a = TK_DIV(fmax(0.0, TK_SUB(blj, xj)), fmax(0.1, fabs(blj)));
        const double b = TK_DIV(fmax(0.0, TK_SUB(xj, buj)), fmax(0.1, fabs(buj)));
        sc.pn[lane] = TK_ADD(TK_MUL(ctx.penaltyc, TK_MUL(a, a)), TK_MUL(ctx.penaltyc, TK_MUL(b, b)));
    }
    __syncwarp();
    double yc = 0.0;
    if (lane == 0) {
        double pen = 0.0;
        if (oob) {
            double penalty = 0.0;
            for (int i = 0; i < nx; i++) penalty = TK_ADD(penalty, sc.pn[i]);
            if (penalty > TK_MYEPS) pen = fmin(ctx.max_penalty, TK_ADD(ctx.penalty0, penalty));
        }
        double f;
        if (pen > TK_MYEPS) {
            f = TK_DIV(pen, ctx.sqrt_nm); /* objective.f90:93-94 */
        } else {
            double s_, p_;
            Obj::fold_init(ctx, s_, p_);
            const double2* qa = reinterpret_cast<const double2*>(sc.ta);
            const double2* qb = reinterpret_cast<const double2*>(sc.tb);
            if (NXC) {
                double2 va[(NXC + 1) / 2 + 1], vb[(NXC + 1) / 2 + 1];
#pragma unroll
                for (int i = 0; i < (NXC + 1) / 2; i++) { va[i] = qa[i]; vb[i] = qb[i]; }
#pragma unroll
                for (int i = 0; i < (NXC + 1) / 2; i++) {
                    if (2 * i < nt) Obj::fold(s_, p_, va[i].x, vb[i].x);
                    if (2 * i + 1 < nt) Obj::fold(s_, p_, va[i].y, vb[i].y);
                }
            } else {
                int i = 0;
                for (; i + 2 <= nt; i += 2) {
                    const double2 va = qa[i >> 1], vb = qb[i >> 1];
                    Obj::fold(s_, p_, va.x, vb.x);
                    Obj::fold(s_, p_, va.y, vb.y);
                }
                if (i < nt) Obj::fold(s_, p_, sc.ta[i], sc.tb[i]);
            }
            f = Obj::finish(ctx, s_, p_);
        }
        yc = tk_objfun_finish(ctx, f, pen);
    }
    __syncwarp();
    return yc;
}

#ifndef TK_NMQ_MINB
#define TK_NMQ_MINB 0 /* ptxas decides (72 registers, 6-7 resident blocks per SM); 8 (64 registers) measured slower: C2 4.8 against 4.2 ms, C3 39.7 against 35.9 ms */
#endif
/* One restart by the four warps of a block (runAmoeba :622-627 + amoeba): shared by the per-launch kernel (one block per
 * restart of the launch) and the asynchronous-instances kernel (a block is an instance and runs restart after restart).
 * start(i): coordinate i of the blended start; tick(trips, shrinks): called by every thread once per loop trip (the
 * asynchronous kernel publishes its clock there).  Every warp ends with the same sorted list: ylo, the first minimal
 * vertex ilo (row of this warp's copy p of the simplex), amoeba's evaluation count iter, loop trips and shrink steps. */
struct NmQuadOut {
    double ylo;
    int ilo, iter, trips, shrinks;
    double* p;
};
template <class Obj, int NXC, class StartF, class TickF>
__device__ __forceinline__ NmQuadOut nm_quad_search(const tk_obj_ctx& ctx, const NmArgs& a, int nx, bool shr, double* sbl, double* sbu,
                                                    double* yv, double* yx, double* mine, StartF start, TickF tick) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int np = nx + 1;
    const int nt = Obj::nterms(nx);
    const int ts = (nx + 2) & ~1; /* even stride >= np */
    NmScratch scr;
    scr.ta = mine;
    scr.tb = mine + ts;
    scr.pn = mine + 2 * ts;
    scr.cand = mine + 3 * ts;
    int* scr_r = reinterpret_cast<int*>(mine + 4 * ts); /* sort scratch (with scr.cand) */
    double* p = mine + 5 * ts;                  /* p[i*nx + j]: this warp's copy of the simplex */
    const bool has_j = lane < nx, has_i = lane < np;
    const int j = has_j ? lane : 0;
    const double blj = sbl[j], buj = sbu[j];

    /* runAmoeba :622-627 -- initial simplex around the blended start */
    const double* wsrc = shr ? a.wshr : a.width;
    for (int e = lane; e < np * nx; e += 32) {
        const int ia = e / nx, i = e - ia * nx;
        double t = TK_ADD(start(i), TK_DIV(TK_MUL(a.simplex[(size_t)ia * nx + i], wsrc[i]), a.scale));
        t = fmax(fmin(t, sbu[i]), sbl[i]);
        p[e] = t;
    }
    __syncwarp();
    /* objFun at the vertices: dealt over the four warps, collected through yv */
    for (int v = w; v < np; v += 4) {
        const double yc = nm_eval_warp<Obj, NXC>(ctx, scr, p[v * nx + j], blj, buj, lane, nx, nt);
        if (lane == 0) yv[v] = yc;
    }
    __syncthreads();
    double ys = has_i ? yv[lane] : 0.0; /* y of the vertex at sorted position `lane` */
    int row = lane;                     /* ... and which vertex (row of p) that is */
    nm_full_sort(ys, row, np, lane, scr.cand, scr_r);
    double psum = 0.0; /* psum(lane), amoeba :49 in index order */
    if (has_j)
        for (int i = 0; i < np; i++) psum = TK_ADD(psum, p[i * nx + j]);

    int iter = 0, trips = 0, shrinks = 0;
    unsigned par = 0;
    double ylo;
    for (;;) {
        ylo = __shfl_sync(FULL, ys, 0);
        const double yhi = __shfl_sync(FULL, ys, nx);
        const double ynhi = __shfl_sync(FULL, ys, nx - 1);
        const int ihi = __shfl_sync(FULL, row, nx);
        /* the four possible trial points (amotry :124-126, unfused); this warp evaluates number w */
        const double ph = p[ihi * nx + j];
        const double R = TK_SUB(TK_MUL(psum, a.fac1r), TK_MUL(ph, a.fac2r));
        const double ps1 = TK_ADD(TK_SUB(psum, ph), R); /* psum after R is accepted (:130) */
        const double E = TK_SUB(TK_MUL(ps1, a.fac1e), TK_MUL(R, a.fac2e));
        const double Co = TK_SUB(TK_MUL(ps1, a.fac1c), TK_MUL(R, a.fac2c));
        const double Ci = TK_SUB(TK_MUL(psum, a.fac1c), TK_MUL(ph, a.fac2c));
        /* rtol < ftol (:84-85); the division is only formed when the product test is within 2^-40 */
        const double num = TK_MUL(2.0, fabs(TK_SUB(yhi, ylo)));
        const double den = TK_ADD(TK_ADD(fabs(yhi), fabs(ylo)), a.tiny);
        const double thr = TK_MUL(a.ftol, den);
        bool conv = num < TK_MUL(thr, 0.99999999999909051);
        if (!conv && !(num > TK_MUL(thr, 1.00000000000090949))) conv = TK_DIV(num, den) < a.ftol;
        if (conv || iter >= a.itmax) break;
        trips++;
        tick(trips, shrinks);

        const double xc = (w == 0) ? R : (w == 1) ? E : (w == 2) ? Co : Ci;
        const double yc = nm_eval_warp<Obj, NXC>(ctx, scr, xc, blj, buj, lane, nx, nt);
        if (lane == 0) yx[par * 4 + w] = yc;
        __syncthreads(); /* the one block barrier of the iteration; yx is double-buffered by parity */
        const double2 y01 = reinterpret_cast<const double2*>(yx + par * 4)[0];
        const double2 y23 = reinterpret_cast<const double2*>(yx + par * 4)[1];
        par ^= 1u;
        const double yR = y01.x;
        const bool accR = yR < yhi; /* :128 */
        double pnew = ph, ynew = yhi;
        bool moved = accR;
        if (accR) { psum = ps1; pnew = R; ynew = yR; }
        iter++;
        if (yR <= ylo) { /* :99 expansion */
            double yE = y01.y;
            double Ev = E;
            if (!accR) { /* only when every y is equal; not speculated -- evaluated from the live state */
                Ev = TK_SUB(TK_MUL(psum, a.fac1e), TK_MUL(ph, a.fac2e));
                yE = __shfl_sync(FULL, nm_eval_warp<Obj, NXC>(ctx, scr, Ev, blj, buj, lane, nx, nt), 0);
            }
            iter++;
            if (yE < ynew) {
                psum = TK_ADD(TK_SUB(psum, pnew), Ev);
                pnew = Ev; ynew = yE; moved = true;
            }
        } else if (yR >= ynhi) { /* :102 contraction */
            const double yC = accR ? y23.x : y23.y;
            const double Cv = accR ? Co : Ci;
            iter++;
            if (yC < ynew) {
                psum = TK_ADD(TK_SUB(psum, pnew), Cv);
                pnew = Cv; ynew = yC; moved = true;
            } else { /* :106-113 shrink toward the best vertex, re-evaluate, re-sum */
                shrinks++;
                if (moved && has_j) p[ihi * nx + j] = pnew;
                if (lane == nx) ys = ynew;
                moved = false;
                /* MINLOC: the lowest vertex index among the minima.  ynew > ylo here, so lane nx is not one. */
                const int ilo = (int)__reduce_min_sync(FULL, (has_i && ys == ylo) ? (unsigned)row : 0xffffffffu);
                __syncwarp();
                if (has_j) {
                    const double pl = p[ilo * nx + j];
                    for (int i = 0; i < np; i++)
                        if (i != ilo) p[i * nx + j] = TK_MUL(0.5, TK_ADD(p[i * nx + j], pl));
                }
                __syncwarp();
                for (int v = w; v < np; v += 4) {
                    if (v == ilo) continue;
                    const double yc2 = nm_eval_warp<Obj, NXC>(ctx, scr, p[v * nx + j], blj, buj, lane, nx, nt);
                    if (lane == 0) yv[v] = yc2;
                }
                __syncthreads();
                if (has_i && row != ilo) ys = yv[row];
                iter += nx;
                nm_full_sort(ys, row, np, lane, scr.cand, scr_r);
                psum = 0.0;
                if (has_j)
                    for (int i = 0; i < np; i++) psum = TK_ADD(psum, p[i * nx + j]);
                __syncthreads(); /* yv may be rewritten by the next shrink only after every warp has read it */
            }
        }
        if (moved) { /* vertex ihi takes (pnew, ynew): remove lane nx, insert at its sorted position */
            if (has_j) p[ihi * nx + j] = pnew;
            const int pos = __popc(__ballot_sync(FULL, lane < nx && nm_before(ys, row, ynew, ihi)));
            const double yu = __shfl_up_sync(FULL, ys, 1);
            const int ru = __shfl_up_sync(FULL, row, 1);
            if (lane > pos && lane <= nx) { ys = yu; row = ru; }
            if (lane == pos) { ys = ynew; row = ihi; }
        }
        __syncwarp();
    }
    NmQuadOut o;
    o.ylo = ylo;
    o.ilo = (int)__reduce_min_sync(FULL, (has_i && ys == ylo) ? (unsigned)row : 0xffffffffu);
    o.iter = iter; o.trips = trips; o.shrinks = shrinks; o.p = p;
    return o;
}

template <class Obj, int NXC>
__global__ void __launch_bounds__(128, TK_NMQ_MINB) nm_quad_kernel(const TkTablesG T, const TkScalars sc, const NmArgs a) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nx = NXC ? NXC : sc.nx, np = nx + 1;
    const int ts = (nx + 2) & ~1; /* even stride >= np */
    /* block-shared */
    double* sbl = smem;
    double* sbu = sbl + ts;
    double* saux = sbu + ts;
    double* yv = saux + ts;                     /* yv[vertex]: values of re-evaluated vertices (start, shrink) */
    double* yx = yv + ts;                       /* yx[parity*4 + c]: the four trial values of an iteration */
    double* mine = yx + 8 + (size_t)w * (5 * ts + (size_t)np * nx); /* per warp: term scratch, sort scratch, simplex */
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    __syncthreads();
    const int slot = a.rank + (int)blockIdx.x * a.world;
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, sbl, sbu, saux);
    const bool shr = a.shr_have && a.shr_have[0] && ((float)(a.first_k + slot) / (float)a.K > 0.5f);
    const NmQuadOut o = nm_quad_search<Obj, NXC>(ctx, a, nx, shr, sbl, sbu, yv, yx, mine,
                                                 [&](int i) { return a.starts[(size_t)i * a.n_k + slot]; }, [](int, int) {});
    /* runAmoeba :631-632 returns the first minimal vertex; completeSearch :582 re-evaluates objFun there:
     * same point, same value. */
    if (w == 0) {
        if (lane < nx) a.out_x[(size_t)lane * a.n_k + slot] = o.p[o.ilo * nx + lane];
        if (lane == 0) {
            a.out_f[slot] = o.ylo;
            a.out_evals[slot] = np + o.iter + 1;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Asynchronous instances (DESIGN.md "asynchronous instances").  The reference's own parallel mode is P independent
 * processes: an instance that finishes restart k' appends its row to searchResults.dat and takes the next restart
 * number; the new start is blended toward the best row of the file AS IT IS THEN (TiktakGlobalSearch.f90:546-600,
 * 683-720).  Nobody waits for a round to end.  Here a block is an instance.  To keep the run reproducible, "then" is
 * a VIRTUAL time that advances by one per loop trip of amoeba (a trip costs the four warps the same real time
 * whatever it counts as evaluations; the initial simplex and a shrink step count c0 = ceil((nx+1)/4) trips):
 *   * an instance that finishes at virtual time F takes restart number first_k + P + #(finish events before (F, id)),
 *     i.e. restarts are handed out in the order of the virtual finish times, ties by instance number;
 *   * its start is blended toward the best row among the rows that finished at virtual times <= F
 *     (lowest value; ties by (finish time, instance, k) -- the order of the file);
 * which is exactly what a sequential event simulation of P instances computes (the oracle's local_async).  On the
 * device an instance may only act at virtual time F when no other instance can still produce a finish event <= F:
 * every block publishes a lower bound of its next finish time (clk), and waits until all the others' exceed F.
 * Real time per trip is nearly the same on every SM, so the wait is short; the instance with the smallest F never
 * waits for long (every running clock grows), hence no deadlock -- PROVIDED all P blocks are resident: the launcher
 * asks the occupancy calculator and launches cooperatively.
 * The block itself writes the row (completeSearch :582-597 through the f40.20 record), so there is no ingest kernel.
 * --------------------------------------------------------------------------------------------- */
__global__ void nm_async_init_kernel(const NmAsync q, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= q.K) {
        const bool known = i < q.first_k && q.valid[i];
        TkAsyncRow r;
        r.fv = known ? q.res[(size_t)i * (q.nx + 1)] : 0.0;
        r.fin = known ? 0 : -1;
        r.slt = -1;
        q.ev[i] = r;
        if (i >= q.first_k && i <= q.last_k) q.valid[i] = 0;
    }
    if (i < P) q.clk[i] = q.cb + q.c0;
    if (i == P) { q.clk[P] = min(q.last_k, q.first_k + P - 1); q.clk[P + 1] = 0; /* commits so far (free-running) */ }
    if (i < TK_GVT_COPIES) q.clk[P + 32 + i * 32] = 0; /* the time keeper's words */
    if (i == 0) *q.err = 0;
}

/* FREE: free-running instances (a template parameter: the reproducible form keeps its 84 registers = six blocks per SM) */
template <class Obj, int NXC, bool FREE>
__global__ void __launch_bounds__(128, TK_NMQ_MINB) nm_quad_async_kernel(const TkTablesG T, const TkScalars sc, const NmArgs a, const NmAsync q) {
    extern __shared__ double smem[];
    __shared__ double r_f[4];
    __shared__ int r_t[4], r_s[4], r_k[4], r_cnt[4], r_ev[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
    const int nx = NXC ? NXC : sc.nx, np = nx + 1;
    const int ts = (nx + 2) & ~1;
    double* sbl = smem;
    double* sbu = sbl + ts;
    double* saux = sbu + ts;
    double* yv = saux + ts;
    double* yx = yv + ts;
    double* st = yx + 8;                        /* the blended start of the current restart */
    double* mine = st + ts + (size_t)w * (5 * ts + (size_t)np * nx);
    for (int i = tid; i < nx; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    __syncthreads();
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, sbl, sbu, saux);
    const int s = (int)blockIdx.x, P = (int)gridDim.x - 1;
    const int c0 = q.c0, cb = q.cb;
    /* The last block is the time keeper: one warp folds the P clocks into their minimum (the global virtual time, a lower bound of
     * every finish event still to come) and publishes it in TK_GVT_COPIES words on different lines.  Instances wait on ONE word:
     * hundreds of blocks polling all P clocks saturate the L2 slice that holds them, and the runners' ticks queue up behind the
     * polls (measured: C3 with 592 instances 155 ms, of which 90 % waiting). */
    if (s == P) {
        if (w != 0) return;
        int* gvt = q.clk + P + 32;
        for (;;) {
            int m = 0x7fffffff;
            for (int i = lane; i < P; i += 32) m = min(m, ld_vol(q.clk + i));
            m = (int)__reduce_min_sync(0xffffffffu, (unsigned)m);
            if (lane < TK_GVT_COPIES) st_vol(gvt + lane * 32, m);
            if (m == 0x7fffffff) return;
            __nanosleep(100);
        }
    }
    long long dbg_wait = 0, dbg_scan = 0, dbg_run = 0, dbg_n = 0, dbg_polls = 0, dbg_trips = 0;
    int k = q.first_k + s, Tv = 0; /* the first P restarts start at virtual time 0 and see the rows known before the launch */
    for (;;) {
        const long long t_scan0 = clock64();
        /* one pass over the rows handed out so far: finish events in front of (Tv, s) -> which restart this instance takes; the best
         * visible row.  (Plain L2 loads: every record with a finish time <= Tv was complete before the wait below ended.) */
        int cnt_ev = 0, cnt_vis = 0, bt = 0, bs = 0, bk = -1;
        double bf = 0.0;
        const int hi = min(q.last_k, __ldcg(q.clk + P));
        for (int jj = tid; jj <= hi; jj += 128) {
            const int4 raw = __ldcg(reinterpret_cast<const int4*>(q.ev + jj));
            const int tj = raw.z, sj = raw.w;
            if (tj < 0 || (!FREE && tj > Tv)) continue;
            if (jj >= q.first_k && (tj < Tv || sj < s)) cnt_ev++;
            const double fj = __hiloint2double(raw.y, raw.x);
            if (!(fj == fj)) continue; /* a NaN value is no row (ingest_kernel) */
            cnt_vis++;
            if (bk < 0 || nm_row_before(fj, tj, sj, jj, bf, bt, bs, bk)) { bf = fj; bt = tj; bs = sj; bk = jj; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double f2 = __shfl_xor_sync(0xffffffffu, bf, o);
            const int t2 = __shfl_xor_sync(0xffffffffu, bt, o), s2 = __shfl_xor_sync(0xffffffffu, bs, o), k2 = __shfl_xor_sync(0xffffffffu, bk, o);
            if (k2 >= 0 && (bk < 0 || nm_row_before(f2, t2, s2, k2, bf, bt, bs, bk))) { bf = f2; bt = t2; bs = s2; bk = k2; }
            cnt_ev += __shfl_xor_sync(0xffffffffu, cnt_ev, o);
            cnt_vis += __shfl_xor_sync(0xffffffffu, cnt_vis, o);
        }
        __syncthreads(); /* (the r_* of the previous restart have been read by everybody) */
        if (lane == 0) { r_f[w] = bf; r_t[w] = bt; r_s[w] = bs; r_k[w] = bk; r_cnt[w] = cnt_vis; r_ev[w] = cnt_ev; }
        __syncthreads();
        bf = r_f[0]; bt = r_t[0]; bs = r_s[0]; bk = r_k[0]; cnt_vis = r_cnt[0]; cnt_ev = r_ev[0];
        for (int o = 1; o < 4; o++) {
            if (r_k[o] >= 0 && (bk < 0 || nm_row_before(r_f[o], r_t[o], r_s[o], r_k[o], bf, bt, bs, bk))) { bf = r_f[o]; bt = r_t[o]; bs = r_s[o]; bk = r_k[o]; }
            cnt_vis += r_cnt[o]; cnt_ev += r_ev[o];
        }
        if (FREE) { /* the next number from the counter (clk[P]: highest number handed out) */
            if (Tv > 0) {
                __syncthreads();
                if (tid == 0) r_ev[0] = atomicAdd(q.clk + P, 1) + 1;
                __syncthreads();
                k = r_ev[0];
            }
        } else if (Tv > 0) k = q.first_k + P + cnt_ev; /* P restarts were taken at time 0, one more at every finish event */
        if (k > q.last_k) {
            if (tid == 0) { st_vol(q.clk + s, 0x7fffffff); if (q.dbg) { q.dbg[s * 5] = dbg_wait; q.dbg[s * 5 + 1] = dbg_scan; q.dbg[s * 5 + 2] = dbg_run; q.dbg[s * 5 + 3] = dbg_n; q.dbg[s * 5 + 4] = dbg_trips; } }
            return;
        }
        if (tid == 0 && Tv > 0 && !FREE) atomicMax(q.clk + P, k);
        if (tid == 0) q.base[k] = (cnt_vis > 1 && bk >= 0) ? bk : -1;
        /* getModifiedParam (:683-720), unfused; the searchStart.dat row */
        for (int d = tid; d < nx; d += 128) {
            const double ev = q.xs[(size_t)(d + 1) * q.K + (k - 1)];
            double out = ev;
            if (cnt_vis > 1 && bk >= 0) { /* count <= 1: no finished local minimisation yet (:701) */
                const double wgt = q.w11[k];
                out = TK_ADD(TK_MUL(wgt, q.res[(size_t)bk * (nx + 1) + 1 + d]), TK_MUL(TK_SUB(1.0, wgt), ev));
            }
            st[d] = out;
            q.arch_starts[(size_t)d * q.K + (k - 1)] = out;
        }
        __syncthreads();
        const int T0 = Tv;
        const long long t_run0 = clock64();
        dbg_scan += t_run0 - t_scan0;
        const NmQuadOut o = nm_quad_search<Obj, NXC>(ctx, a, nx, false, sbl, sbu, yv, yx, mine, [&](int i) { return st[i]; },
                                                     [&](int trips, int shrinks) { if (tid == 0) st_vol(q.clk + s, T0 + cb + c0 + trips + shrinks * c0); });
        int F = T0 + cb + c0 + o.trips + o.shrinks * c0;
        /* completeSearch :582-597: the row of restart k (through the f40.20 record when the text round trip is on) */
        if (w == 0) {
            double* dst = q.res + (size_t)k * (nx + 1);
            const double frt = q.roundtrip ? tk_f4020_roundtrip(o.ylo) : o.ylo;
            if (lane < nx) { const double xv = o.p[o.ilo * nx + lane]; dst[1 + lane] = q.roundtrip ? tk_f4020_roundtrip(xv) : xv; }
            if (lane == 0) {
                dst[0] = frt;
                q.ev[k].fv = frt;
                q.ev[k].slt = s;
                q.valid[k] = (frt == frt) ? 1 : 0;
                q.arch_evals[k] = np + o.iter + 1;
            }
            __syncwarp();
            if (FREE) F = 1 + (lane == 0 ? atomicAdd(q.clk + P + 1, 1) : 0); /* the order of the commits is the order of the file */
            F = __shfl_sync(0xffffffffu, F, 0);
            if (lane == 0) {
                __threadfence();
                st_vol(&q.ev[k].fin, F);            /* the finish event exists from here on ... */
                __threadfence();
                st_vol(q.clk + s, F + cb + c0);     /* ... and this instance's next one cannot come before F + cb + c0 */
            }
        }
        const long long t_wait0 = clock64();
        dbg_run += t_wait0 - t_run0; dbg_n++; dbg_trips += c0 + o.trips + o.shrinks * c0;
        /* act at virtual time F only when nobody else can still finish at or before it */
        if (FREE) { /* nobody waits for anybody */
            __syncthreads();
            __threadfence();
            Tv = 1;
            continue;
        }
        if (w == 0) { /* one warp polls one word; the others park at the barrier */
            const int* gvt = q.clk + P + 32 + (s % TK_GVT_COPIES) * 32;
            unsigned ns = 100;
            for (;;) {
                dbg_polls++;
                if (ld_vol(gvt) > F) break;
                if (clock64() - t_wait0 > TK_ASYNC_PATIENCE) { if (lane == 0) atomicExch(q.err, 1); break; } /* never in a healthy run */
                __nanosleep(ns);
                if (ns < 800) ns += ns;
            }
        }
        __syncthreads();
        if (ld_vol(q.err)) { if (tid == 0) st_vol(q.clk + s, 0x7fffffff); return; }
        __threadfence();
        dbg_wait += clock64() - t_wait0;
        Tv = F;
    }
}

/* ---------------------------------------------------------------------------------------------
 * The same four-warp scheme for 32 <= nx <= 127 (CJ = ceil((nx+1)/32) coordinates / vertices per lane):
 * lane l owns coordinates j = l + 32c, the sorted (y, vertex) list lives in the warp's shared memory
 * (insertion = CJ ballots + a shift by one), the simplex p is ONE shared copy per block (every warp stores
 * the same values), and the in-order fold of the nt terms on lane 0 is the longest dependent chain of the
 * iteration (nt DADDs -- the price of the reference's index-order sum).
 * --------------------------------------------------------------------------------------------- */
template <class Obj, int CJ>
__device__ __forceinline__ double nm_eval_warp_wide(const tk_obj_ctx& ctx, const NmScratch& sc, const double (&xj)[CJ],
                                                    const double (&blj)[CJ], const double (&buj)[CJ], int lane, int nx, int nt) {
    bool ob = false;
#pragma unroll
    for (int c = 0; c < CJ; c++) ob = ob || ((lane + 32 * c < nx) && (xj[c] < blj[c] || xj[c] > buj[c]));
    const bool oob = __any_sync(0xffffffffu, ob);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CJ; c++)
        if (lane + 32 * c < nx) sc.cand[lane + 32 * c] = xj[c];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CJ; c++) {
        const int i = lane + 32 * c;
        if (i < nt) {
            double ta_, tb_;
            Obj::term(ctx, TkMemX{sc.cand, 1}, i, ta_, tb_);
            sc.ta[i] = ta_; sc.tb[i] = tb_;
        }
        if (oob && i < nx) { /* getPenalty :57-66, coordinate i */
            const double a = TK_DIV(fmax(0.0, TK_SUB(blj[c], xj[c])), fmax(0.1, fabs(blj[c])));
            const double b = TK_DIV(fmax(0.0, TK_SUB(xj[c], buj[c])), fmax(0.1, fabs(buj[c])));
            sc.pn[i] = TK_ADD(TK_MUL(ctx.penaltyc, TK_MUL(a, a)), TK_MUL(ctx.penaltyc, TK_MUL(b, b)));
        }
    }
    __syncwarp();
    double yc = 0.0;
    if (lane == 0) {
        double pen = 0.0;
        if (oob) {
            double penalty = 0.0;
            for (int i = 0; i < nx; i++) penalty = TK_ADD(penalty, sc.pn[i]);
            if (penalty > TK_MYEPS) pen = fmin(ctx.max_penalty, TK_ADD(ctx.penalty0, penalty));
        }
        double f;
        if (pen > TK_MYEPS) {
            f = TK_DIV(pen, ctx.sqrt_nm); /* objective.f90:93-94 */
        } else {
            double s_, p_;
            Obj::fold_init(ctx, s_, p_);
            const double2* qa = reinterpret_cast<const double2*>(sc.ta);
            const double2* qb = reinterpret_cast<const double2*>(sc.tb);
            int i = 0;
            for (; i + 8 <= nt; i += 8) { /* loads first, then the in-order chain */
                const double2 a0 = qa[(i >> 1)], a1 = qa[(i >> 1) + 1], a2 = qa[(i >> 1) + 2], a3 = qa[(i >> 1) + 3];
                const double2 b0 = qb[(i >> 1)], b1 = qb[(i >> 1) + 1], b2 = qb[(i >> 1) + 2], b3 = qb[(i >> 1) + 3];
                Obj::fold(s_, p_, a0.x, b0.x); Obj::fold(s_, p_, a0.y, b0.y);
                Obj::fold(s_, p_, a1.x, b1.x); Obj::fold(s_, p_, a1.y, b1.y);
                Obj::fold(s_, p_, a2.x, b2.x); Obj::fold(s_, p_, a2.y, b2.y);
                Obj::fold(s_, p_, a3.x, b3.x); Obj::fold(s_, p_, a3.y, b3.y);
            }
            for (; i < nt; i++) Obj::fold(s_, p_, sc.ta[i], sc.tb[i]);
            f = Obj::finish(ctx, s_, p_);
        }
        yc = tk_objfun_finish(ctx, f, pen);
    }
    __syncwarp();
    return yc;
}

/* sort the np (y, vertex) pairs of the warp's list in shared memory (rank by counting, then scatter) */
template <int CJ>
__device__ __forceinline__ void nm_full_sort_wide(double* ys, int* rows, int np, int lane, double* scr_y, int* scr_r) {
    __syncwarp();
    double my[CJ]; int mr[CJ], rk[CJ];
#pragma unroll
    for (int c = 0; c < CJ; c++) {
        const int e = lane + 32 * c;
        my[c] = (e < np) ? ys[e] : 0.0; mr[c] = (e < np) ? rows[e] : 0; rk[c] = 0;
    }
    for (int k = 0; k < np; k++) {
        const double yk = ys[k]; const int rkk = rows[k];
#pragma unroll
        for (int c = 0; c < CJ; c++) rk[c] += (k != lane + 32 * c && nm_before(yk, rkk, my[c], mr[c])) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CJ; c++)
        if (lane + 32 * c < np) { scr_y[rk[c]] = my[c]; scr_r[rk[c]] = mr[c]; }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CJ; c++) {
        const int e = lane + 32 * c;
        if (e < np) { ys[e] = scr_y[e]; rows[e] = scr_r[e]; }
    }
    __syncwarp();
}

/* Roles: warp 0 is the restart's "master" (sorted list, psum, trial points, decisions, simplex updates); all four
 * warps evaluate.  Per iteration the master writes the four trial points to shared memory, barrier A, every warp
 * evaluates one of them, barrier B, the master decides.  With several blocks per SM the kernel is issue-bound, and
 * three of four warps now run only the evaluation (no redundant bookkeeping): ~2.4x fewer instructions per iteration
 * than carrying the state in every warp. */
enum { NMW_EVAL4 = 0, NMW_VERTS = 1, NMW_STOP = 2 };
#ifndef TK_NMW_MB
#define TK_NMW_MB 0 /* resident blocks per SM the wide kernel is compiled for (0: ptxas decides) */
#endif
template <class Obj, int CJ>
__global__ void __launch_bounds__(128, TK_NMW_MB) nm_quad_wide_kernel(const TkTablesG T, const TkScalars sc, const NmArgs a) {
    extern __shared__ double smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool master = w == (int)(blockIdx.x & 3u); /* rotate the master over the SM's four sub-partitions */
    const int nx = sc.nx, np = nx + 1;
    const int nt = Obj::nterms(nx);
    const int ts = (nx + 2) & ~1; /* even stride >= np */
    /* block-shared */
    double* sbl = smem;
    double* sbu = sbl + ts;
    double* saux = sbu + ts;
    double* yv = saux + ts;      /* yv[vertex]: values of re-evaluated vertices (start, shrink) */
    double* xcand = yv + ts;     /* xcand[c*ts + j]: the four trial points of the iteration */
    double* yx = xcand + 4 * ts; /* yx[c]: their values */
    int* ctl = reinterpret_cast<int*>(yx + 4); /* ctl[0] = mode, ctl[1] = vertex skipped by NMW_VERTS */
    double* p = yx + 6;          /* p[i*nx + j]: the simplex (written by the master only) */
    /* master's sorted list + sort scratch */
    double* ys = p + (size_t)np * nx;
    int* rows = reinterpret_cast<int*>(ys + ts);
    double* scr_y = ys + 2 * ts;
    int* scr_r = reinterpret_cast<int*>(ys + 3 * ts);
    /* per warp evaluation scratch */
    double* mine = ys + 4 * ts + (size_t)w * (4 * ts);
    NmScratch scr;
    scr.ta = mine;
    scr.tb = mine + ts;
    scr.pn = mine + 2 * ts;
    scr.cand = mine + 3 * ts;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    const int slot = a.rank + (int)blockIdx.x * a.world;
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, sbl, sbu, saux);
    /* runAmoeba :622-627 -- initial simplex around the blended start */
    const bool shr = a.shr_have && a.shr_have[0] && ((float)(a.first_k + slot) / (float)a.K > 0.5f);
    const double* wsrc = shr ? a.wshr : a.width;
    for (int e = threadIdx.x; e < np * nx; e += blockDim.x) {
        const int ia = e / nx, i = e - ia * nx;
        double t = TK_ADD(a.starts[(size_t)i * a.n_k + slot], TK_DIV(TK_MUL(a.simplex[(size_t)ia * nx + i], wsrc[i]), a.scale));
        t = fmax(fmin(t, T.bu[i]), T.bl[i]);
        p[e] = t;
    }
    if (threadIdx.x == 0) { ctl[0] = NMW_VERTS; ctl[1] = -1; } /* first pass: objFun at every vertex */
    __syncthreads();
    double blj[CJ], buj[CJ], psum[CJ];
    int jj[CJ];
#pragma unroll
    for (int c = 0; c < CJ; c++) {
        jj[c] = (lane + 32 * c < nx) ? lane + 32 * c : 0;
        blj[c] = sbl[jj[c]]; buj[c] = sbu[jj[c]]; psum[c] = 0.0;
    }
    /* master state that lives across iterations */
    int iter = 0, ihi = 0, ilo_keep = -1;
    bool first = true, after_shrink = false;
    double ylo = 0.0, yhi = 0.0, ynhi = 0.0;
    double ph[CJ], R[CJ], ps1[CJ], E[CJ], Co[CJ], Ci[CJ];
#pragma unroll
    for (int c = 0; c < CJ; c++) { ph[c] = R[c] = ps1[c] = E[c] = Co[c] = Ci[c] = 0.0; }

    for (;;) {
        /* ---- barrier A was prepared by the master (or by the prologue above) ---- */
        __syncthreads();
        const int mode = ctl[0];
        if (mode == NMW_STOP) break;
        if (mode == NMW_EVAL4) {
            double xc[CJ];
#pragma unroll
            for (int c = 0; c < CJ; c++) xc[c] = xcand[w * ts + jj[c]];
            const double yc = nm_eval_warp_wide<Obj, CJ>(ctx, scr, xc, blj, buj, lane, nx, nt);
            if (lane == 0) yx[w] = yc;
        } else { /* NMW_VERTS: the vertices dealt over the warps */
            const int skip = ctl[1];
            for (int v = w; v < np; v += 4) {
                if (v == skip) continue;
                double xv[CJ];
#pragma unroll
                for (int c = 0; c < CJ; c++) xv[c] = p[v * nx + jj[c]];
                const double yc = nm_eval_warp_wide<Obj, CJ>(ctx, scr, xv, blj, buj, lane, nx, nt);
                if (lane == 0) yv[v] = yc;
            }
        }
        __syncthreads(); /* ---- barrier B: values ready ---- */
        if (!master) continue;

        /* ================= master ================= */
        if (mode == NMW_VERTS) {
            if (first) {
                for (int e = lane; e < np; e += 32) { ys[e] = yv[e]; rows[e] = e; }
                first = false;
            } else { /* after a shrink: every vertex but ilo was re-evaluated */
                for (int e = lane; e < np; e += 32)
                    if (rows[e] != ilo_keep) ys[e] = yv[rows[e]];
                iter += nx;
            }
            nm_full_sort_wide<CJ>(ys, rows, np, lane, scr_y, scr_r);
#pragma unroll
            for (int c = 0; c < CJ; c++) { /* psum, amoeba :49 / :112 in index order */
                double s_ = 0.0;
                for (int i = 0; i < np; i++) s_ = TK_ADD(s_, p[i * nx + jj[c]]);
                psum[c] = s_;
            }
            after_shrink = false;
        } else {
            /* decisions of the iteration (minimize.f90:97-114) */
            const double yR = yx[0];
            const bool accR = yR < yhi; /* :128 */
            double pnew[CJ], ynew = yhi;
            bool moved = accR;
#pragma unroll
            for (int c = 0; c < CJ; c++) { pnew[c] = accR ? R[c] : ph[c]; if (accR) psum[c] = ps1[c]; }
            if (accR) ynew = yR;
            iter++;
            if (yR <= ylo) { /* :99 expansion */
                double yE = yx[1];
                if (!accR) { /* only when every y is equal; not speculated -- evaluated from the live state */
#pragma unroll
                    for (int c = 0; c < CJ; c++) E[c] = TK_SUB(TK_MUL(psum[c], a.fac1e), TK_MUL(ph[c], a.fac2e));
                    yE = __shfl_sync(FULL, nm_eval_warp_wide<Obj, CJ>(ctx, scr, E, blj, buj, lane, nx, nt), 0);
                }
                iter++;
                if (yE < ynew) {
#pragma unroll
                    for (int c = 0; c < CJ; c++) { psum[c] = TK_ADD(TK_SUB(psum[c], pnew[c]), E[c]); pnew[c] = E[c]; }
                    ynew = yE; moved = true;
                }
            } else if (yR >= ynhi) { /* :102 contraction */
                const double yC = accR ? yx[2] : yx[3];
                iter++;
                if (yC < ynew) {
#pragma unroll
                    for (int c = 0; c < CJ; c++) {
                        const double Cv = accR ? Co[c] : Ci[c];
                        psum[c] = TK_ADD(TK_SUB(psum[c], pnew[c]), Cv); pnew[c] = Cv;
                    }
                    ynew = yC; moved = true;
                } else { /* :106-113 shrink toward the best vertex; the re-evaluation is the next pass */
                    if (moved) {
#pragma unroll
                        for (int c = 0; c < CJ; c++)
                            if (lane + 32 * c < nx) p[ihi * nx + jj[c]] = pnew[c];
                    }
                    if (lane == 0) ys[nx] = ynew;
                    moved = false;
                    __syncwarp();
                    unsigned cand_lo = 0xffffffffu; /* MINLOC: the lowest vertex index among the minima */
                    for (int e = lane; e < np; e += 32)
                        if (ys[e] == ylo) cand_lo = min(cand_lo, (unsigned)rows[e]);
                    const int ilo = (int)__reduce_min_sync(FULL, cand_lo);
#pragma unroll
                    for (int c = 0; c < CJ; c++) {
                        if (lane + 32 * c >= nx) continue;
                        const double pl = p[ilo * nx + jj[c]];
                        for (int i = 0; i < np; i++)
                            if (i != ilo) p[i * nx + jj[c]] = TK_MUL(0.5, TK_ADD(p[i * nx + jj[c]], pl));
                    }
                    ilo_keep = ilo;
                    after_shrink = true;
                }
            }
            if (moved) { /* vertex ihi takes (pnew, ynew): drop position nx, insert at its sorted position */
                int pos = 0;
                double ye[CJ]; int re[CJ];
#pragma unroll
                for (int c = 0; c < CJ; c++) {
                    const int e = lane + 32 * c;
                    if (e < nx) p[ihi * nx + e] = pnew[c];
                    ye[c] = (e < nx) ? ys[e] : 0.0; re[c] = (e < nx) ? rows[e] : 0;
                    pos += __popc(__ballot_sync(FULL, e < nx && nm_before(ye[c], re[c], ynew, ihi)));
                }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CJ; c++) {
                    const int e = lane + 32 * c;
                    if (e < nx && e >= pos) { ys[e + 1] = ye[c]; rows[e + 1] = re[c]; }
                }
                if (lane == 0) { ys[pos] = ynew; rows[pos] = ihi; }
            }
        }
        __syncwarp();
        /* ---- prepare the next pass ---- */
        if (after_shrink) {
            if (lane == 0) { ctl[0] = NMW_VERTS; ctl[1] = ilo_keep; }
            continue;
        }
        ylo = ys[0]; yhi = ys[nx]; ynhi = ys[nx - 1];
        ihi = rows[nx];
        if (nm_converged(a, ylo, yhi) || iter >= a.itmax) {
            if (lane == 0) ctl[0] = NMW_STOP;
            continue;
        }
        /* the four possible trial points (amotry :124-126, unfused) */
#pragma unroll
        for (int c = 0; c < CJ; c++) {
            ph[c] = p[ihi * nx + jj[c]];
            R[c] = TK_SUB(TK_MUL(psum[c], a.fac1r), TK_MUL(ph[c], a.fac2r));
            ps1[c] = TK_ADD(TK_SUB(psum[c], ph[c]), R[c]); /* psum after R is accepted (:130) */
            E[c] = TK_SUB(TK_MUL(ps1[c], a.fac1e), TK_MUL(R[c], a.fac2e));
            Co[c] = TK_SUB(TK_MUL(ps1[c], a.fac1c), TK_MUL(R[c], a.fac2c));
            Ci[c] = TK_SUB(TK_MUL(psum[c], a.fac1c), TK_MUL(ph[c], a.fac2c));
            if (lane + 32 * c < nx) {
                xcand[jj[c]] = R[c]; xcand[ts + jj[c]] = E[c]; xcand[2 * ts + jj[c]] = Co[c]; xcand[3 * ts + jj[c]] = Ci[c];
            }
        }
        if (lane == 0) ctl[0] = NMW_EVAL4;
    }
    /* runAmoeba :631-632 returns the first minimal vertex; completeSearch :582 re-evaluates objFun there:
     * same point, same value. */
    if (master) {
        unsigned cand_lo = 0xffffffffu;
        for (int e = lane; e < np; e += 32)
            if (ys[e] == ylo) cand_lo = min(cand_lo, (unsigned)rows[e]);
        const int ilo = (int)__reduce_min_sync(FULL, cand_lo);
        for (int jx = lane; jx < nx; jx += 32) a.out_x[(size_t)jx * a.n_k + slot] = p[ilo * nx + jx];
        if (lane == 0) {
            a.out_f[slot] = ylo;
            a.out_evals[slot] = np + iter + 1;
        }
    }
}

/* getSortedResults/getBestLine (:781-823) over the device-resident result rows.
 * res: rows 0..K, each [fval, x(nx)]; best: [0]=fval [1]=row [2]=#rows, [3..] = x */
__global__ void best_kernel(const double* res, const int* valid, int nrows, int nx, double* best) {
    __shared__ double sv[256];
    __shared__ int si[256];
    __shared__ int sc[256];
    double v = DBL_MAX; int bi = 0x7fffffff, cnt = 0;
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        if (!valid[r]) continue;
        cnt++;
        const double f = res[(size_t)r * (nx + 1)];
        if (bi == 0x7fffffff || f < v) { v = f; bi = r; }
    }
    sv[threadIdx.x] = v; si[threadIdx.x] = bi; sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = sv[threadIdx.x + o]; const int i2 = si[threadIdx.x + o];
            if (i2 != 0x7fffffff && (si[threadIdx.x] == 0x7fffffff || v2 < sv[threadIdx.x] ||
                                     (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x]))) {
                sv[threadIdx.x] = v2; si[threadIdx.x] = i2;
            }
            sc[threadIdx.x] += sc[threadIdx.x + o];
        }
        __syncthreads();
    }
    const int b = si[0];
    if (threadIdx.x == 0) { best[0] = sv[0]; best[1] = (b == 0x7fffffff) ? -1.0 : (double)b; best[2] = (double)sc[0]; }
    if (b != 0x7fffffff)
        for (int d = threadIdx.x; d < nx; d += blockDim.x) best[3 + d] = res[(size_t)b * (nx + 1) + 1 + d];
}

/* getModifiedParam (:683-720): start = w11*best + (1-w11)*x_starts(k), unfused; w11 is the
 * single-precision weight widened to double (computed on the host, h->w11). */
__global__ void blend_kernel(const double* xs, int K, int nx, int first_k, int n_k, const double* w11, int w_idx,
                             const double* best, double* starts, double* arch_starts) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_k * nx) return;
    const int d = e / n_k, s = e - d * n_k, k = first_k + s; /* k is 1-based */
    const double ev = xs[(size_t)(d + 1) * K + (k - 1)];
    double out = ev;
    if (best[2] > 1.0) { /* count <= 1: no finished local minimisation yet (:701) */
        const double w = w11[w_idx > 0 ? w_idx : k]; /* SolveMissingSearch calls getModifiedParam with whichPoint = p_maxpoints (:1121) */
        out = TK_ADD(TK_MUL(w, best[3 + d]), TK_MUL(TK_SUB(1.0, w), ev));
    }
    starts[(size_t)d * n_k + s] = out;
    arch_starts[(size_t)d * K + (k - 1)] = out; /* the searchStart.dat row of restart k (:557) */
}

/* getLotteryPoint (TiktakGlobalSearch.f90:722-766), selectType = 1: the base point of restart k is the result row of
 * rank i (rows ordered by (fval, k)) with probability proportional to 1/i among the best numPoints rows.
 * cum = the reference's cumulative probOfPoint (computed on the host in its arithmetic), u = tk_uniform01(seed
 * 123456, k) in place of RANDOM_NUMBER; the first i with u < cum(i) wins (the last row if rounding leaves none). */
__global__ void lottery_collect_kernel(const double* res, const int* valid, int nrows, int nx, double* f, unsigned int* kk,
                                       unsigned int* n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows || !valid[r]) return;
    const unsigned int pos = atomicAdd(n, 1u);
    f[pos] = res[(size_t)r * (nx + 1)];
    kk[pos] = (unsigned int)r;
}
__global__ void blend_lottery_kernel(const double* xs, int K, int nx, int first_k, int n_k, const double* w11, int w_idx, int count,
                                     int num_points, const double* cum, const unsigned int* sorted_k, const double* res,
                                     double* starts, double* arch_starts, int* lot_point) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_k * nx) return;
    const int d = e / n_k, s = e - d * n_k, k = first_k + s;
    const double ev = xs[(size_t)(d + 1) * K + (k - 1)];
    double out = ev;
    int lot = 0;
    if (count > 1) { /* :701 */
        const double u = tk_uniform01(TK_LOTTERY_SEED, TK_LOTTERY_STREAM, (uint64_t)k);
        int lo = 0, hi = num_points - 1; /* first i with u < cum[i]; cum is non-decreasing */
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (u < cum[mid]) hi = mid; else lo = mid + 1;
        }
        lot = (int)sorted_k[lo];
        const double w = w11[w_idx > 0 ? w_idx : k];
        out = TK_ADD(TK_MUL(w, res[(size_t)lot * (nx + 1) + 1 + d]), TK_MUL(TK_SUB(1.0, w), ev));
    }
    starts[(size_t)d * n_k + s] = out;
    arch_starts[(size_t)d * K + (k - 1)] = out;
    if (d == 0) lot_point[k] = lot;
}

/* One round's outcomes -> the device-resident searchResults table (completeSearch :582-597 without the
 * file): row k = [fval, x] through the f40.20 record when text_roundtrip is on, the evaluation count, and
 * the running best row (getBestLine :781-802: lowest fval, first row among equals) -- so the next round's
 * blend needs no re-read and no re-sort.  Source: this GPU's slot-indexed outputs (world == 1), or the
 * all-gathered packed rows [rank][n_max][nx+2] = f, x, evals (slot s was run by rank s % world). */
struct IngestArgs {
    int first_k, n_k, nx, world, n_max, roundtrip;
    int round_size; /* restarts per round inside the launch (== n_k: one round) */
    int* accepted;  /* out: restarts ingested (may be NULL) */
    const double* pack;
    const double* out_f;
    const double* out_x;
    const int* out_evals;
    double* res;
    int* valid;
    int* arch_evals;
    double* best;
};
/* The launch may hold SEVERAL rounds (g.n_k restarts, rounds of g.round_size): rounds after the first were started
 * speculatively against the best row known when the launch was queued.  They are ingested in order for as long as
 * that assumption held, i.e. until a round changes the best row (or brings the row count above 1, which switches
 * the blend on, :701); `accepted` = restarts ingested.  The rest of the launch is discarded by the caller and run
 * again against the new best row -- the accepted sequence is exactly the one of strictly sequential rounds. */
__global__ void __launch_bounds__(256) ingest_kernel(const IngestArgs g) {
    __shared__ double sv[256];
    __shared__ int sk[256];
    __shared__ int sc[256];
    __shared__ int take, stop;
    const int nx = g.nx;
    int r0 = 0;
    for (; r0 < g.n_k; r0 += g.round_size) {
        const int r1 = min(g.n_k, r0 + g.round_size);
        double v = DBL_MAX; int bk = 0x7fffffff, cnt = 0;
        for (int s = r0 + threadIdx.x; s < r1; s += blockDim.x) {
            const double* row = g.pack ? g.pack + ((size_t)(s % g.world) * g.n_max + s / g.world) * (nx + 2) : nullptr;
            const double f = row ? row[0] : g.out_f[s];
            if (!(f == f)) continue;
            const int k = g.first_k + s;
            const double frt = g.roundtrip ? tk_f4020_roundtrip(f) : f;
            double* dst = g.res + (size_t)k * (nx + 1);
            dst[0] = frt;
            for (int d = 0; d < nx; d++) {
                const double xv = row ? row[1 + d] : g.out_x[(size_t)d * g.n_k + s];
                dst[1 + d] = g.roundtrip ? tk_f4020_roundtrip(xv) : xv;
            }
            g.valid[k] = 1;
            g.arch_evals[k] = row ? (int)row[nx + 1] : g.out_evals[s];
            cnt++;
            if (bk == 0x7fffffff || frt < v) { v = frt; bk = k; }
        }
        sv[threadIdx.x] = v; sk[threadIdx.x] = bk; sc[threadIdx.x] = cnt;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const double v2 = sv[threadIdx.x + o]; const int k2 = sk[threadIdx.x + o];
                if (k2 != 0x7fffffff && (sk[threadIdx.x] == 0x7fffffff || v2 < sv[threadIdx.x] ||
                                         (v2 == sv[threadIdx.x] && k2 < sk[threadIdx.x]))) {
                    sv[threadIdx.x] = v2; sk[threadIdx.x] = k2;
                }
                sc[threadIdx.x] += sc[threadIdx.x + o];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const bool none = g.best[1] < 0.0;
            const double rows_before = none ? 0.0 : g.best[2];
            take = (sk[0] != 0x7fffffff) && (none || sv[0] < g.best[0]);
            g.best[2] = rows_before + (double)sc[0];
            if (take) { g.best[0] = sv[0]; g.best[1] = (double)sk[0]; }
            stop = take || (rows_before <= 1.0 && sc[0] > 0);
        }
        __syncthreads();
        if (take)
            for (int d = threadIdx.x; d < nx; d += blockDim.x) g.best[3 + d] = g.res[(size_t)sk[0] * (nx + 1) + 1 + d];
        const int stop_now = stop;
        __syncthreads();
        if (stop_now) { r0 = r1; break; }
    }
    if (threadIdx.x == 0 && g.accepted) g.accepted[0] = min(r0, g.n_k);
}

/* this rank's slots of the round, packed for the all-gather: row q = slot rank + q*world = [f, x, evals] */
__global__ void pack_kernel(int n_k, int nx, int rank, int world, const double* out_f, const double* out_x,
                            const int* out_evals, double* pack) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = rank + q * world;
    if (s >= n_k) return;
    double* row = pack + (size_t)q * (nx + 2);
    row[0] = out_f[s];
    for (int d = 0; d < nx; d++) row[1 + d] = out_x[(size_t)d * n_k + s];
    row[nx + 1] = (double)out_evals[s];
}

/* nm_shrink_mode = 1 (README.md:36, SURVEY finding 3): per-coordinate spread of the (up to) 10 best local
 * minima known so far, ordered by (fval, k); wshr = max - min where positive, else the Sobol range width;
 * have = at least two finished restarts.  One block; 10 argmin passes over the result rows. */
__global__ void __launch_bounds__(256) shrink_width_kernel(const double* res, const int* valid, int K, int nx,
                                                           const double* width, double* wshr, int* have) {
    __shared__ double sv[256];
    __shared__ int sk[256];
    __shared__ int topk[10];
    __shared__ int T;
    double pf = 0.0; int pk = -1;
    if (threadIdx.x == 0) T = 0;
    for (int t = 0; t < 10; t++) {
        double v = DBL_MAX; int bk = 0x7fffffff;
        for (int k = 1 + threadIdx.x; k <= K; k += blockDim.x) {
            if (!valid[k]) continue;
            const double f = res[(size_t)k * (nx + 1)];
            if (pk >= 0 && !(f > pf || (f == pf && k > pk))) continue;
            if (bk == 0x7fffffff || f < v) { v = f; bk = k; }
        }
        sv[threadIdx.x] = v; sk[threadIdx.x] = bk;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const double v2 = sv[threadIdx.x + o]; const int k2 = sk[threadIdx.x + o];
                if (k2 != 0x7fffffff && (sk[threadIdx.x] == 0x7fffffff || v2 < sv[threadIdx.x] ||
                                         (v2 == sv[threadIdx.x] && k2 < sk[threadIdx.x]))) {
                    sv[threadIdx.x] = v2; sk[threadIdx.x] = k2;
                }
            }
            __syncthreads();
        }
        const int fk = sk[0];
        const double ff = sv[0];
        __syncthreads();
        if (fk == 0x7fffffff) break;
        if (threadIdx.x == 0) { topk[t] = fk; T = t + 1; }
        pf = ff; pk = fk;
    }
    __syncthreads();
    if (threadIdx.x == 0) have[0] = (T >= 2) ? 1 : 0;
    for (int d = threadIdx.x; d < nx; d += blockDim.x) {
        double w = width[d];
        if (T >= 2) {
            double mn = res[(size_t)topk[0] * (nx + 1) + 1 + d], mx = mn;
            for (int t = 1; t < T; t++) {
                const double x = res[(size_t)topk[t] * (nx + 1) + 1 + d];
                mn = fmin(mn, x); mx = fmax(mx, x);
            }
            const double ww = TK_SUB(mx, mn);
            if (ww > 0.0) w = ww;
        }
        wshr[d] = w;
    }
}

} /* namespace */

int tk_launch_best(tiktak_handle* h) {
    best_kernel<<<1, 256, 0, h->stream>>>(h->d_res, h->d_res_valid, h->cfg.maxpoints_plus1 + 1, h->cfg.nx, h->d_best);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_launch_blend(tiktak_handle* h, int first_k, int n_k, int w_idx) {
    const int tot = n_k * h->cfg.nx;
    blend_kernel<<<(tot + 127) / 128, 128, 0, h->stream>>>(h->d_xstarts, h->cfg.maxpoints_plus1, h->cfg.nx, first_k, n_k,
                                                            h->d_w11, w_idx, h->d_best, h->d_starts, h->d_arch_starts);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_sort_candidates(tiktak_handle* h, double* d_f, unsigned int* d_i, int64_t n, double* d_sf, unsigned int* d_si);

/* the blend of a round under selectType = 1 (lottery); `count` = result rows known to the handle */
int tk_launch_blend_lottery(tiktak_handle* h, int first_k, int n_k, int count, int w_idx) {
    const int nx = h->cfg.nx, K = h->cfg.maxpoints_plus1;
    if (!h->d_lot_f) {
        TK_CUDA(cudaMalloc(&h->d_lot_f, (size_t)(K + 1) * 8)); TK_CUDA(cudaMalloc(&h->d_lot_sf, (size_t)(K + 1) * 8));
        TK_CUDA(cudaMalloc(&h->d_lot_k, (size_t)(K + 1) * 4)); TK_CUDA(cudaMalloc(&h->d_lot_sk, (size_t)(K + 1) * 4));
        TK_CUDA(cudaMalloc(&h->d_lot_cum, (size_t)(K + 1) * 8)); TK_CUDA(cudaMalloc(&h->d_lot_point, (size_t)(K + 1) * 4));
        TK_CUDA(cudaMemsetAsync(h->d_lot_point, 0, (size_t)(K + 1) * 4, h->stream));
    }
    const int num_points = std::max(1, std::min(count, (int)h->cfg.lottery_points));
    if (count > 1) {
        TK_CUDA(cudaMemsetAsync(h->d_cand_n, 0, sizeof(unsigned int), h->stream));
        lottery_collect_kernel<<<(K + 1 + 255) / 256, 256, 0, h->stream>>>(h->d_res, h->d_res_valid, K + 1, nx, h->d_lot_f, h->d_lot_k, h->d_cand_n);
        h->launches++;
        int rc = tk_sort_candidates(h, h->d_lot_f, h->d_lot_k, count, h->d_lot_sf, h->d_lot_sk);
        if (rc) return rc;
        if (num_points != h->lot_cum_n) { /* probOfPoint, :742-754, in the reference's arithmetic */
            std::vector<double>& prob = h->lot_cum;
            prob.assign(num_points, 0.0);
            const long long sumVal = ((long long)num_points * (num_points + 1)) / 2;
            for (int i = 1; i <= num_points; i++) prob[i - 1] = (double)sumVal / (double)i;
            double sum2 = 0.0;
            for (int i = 0; i < num_points; i++) sum2 = sum2 + prob[i];
            for (int i = 0; i < num_points; i++) prob[i] = prob[i] / sum2;
            for (int i = 1; i < num_points; i++) prob[i] = prob[i - 1] + prob[i];
            TK_CUDA(cudaStreamSynchronize(h->stream)); /* an earlier round may still be reading d_lot_cum / lot_cum */
            TK_CUDA(cudaMemcpyAsync(h->d_lot_cum, prob.data(), (size_t)num_points * 8, cudaMemcpyHostToDevice, h->stream));
            h->lot_cum_n = num_points;
        }
    }
    const int tot = n_k * nx;
    blend_lottery_kernel<<<(tot + 127) / 128, 128, 0, h->stream>>>(h->d_xstarts, K, nx, first_k, n_k, h->d_w11, w_idx, count, num_points,
                                                                    h->d_lot_cum, h->d_lot_sk, h->d_res, h->d_starts,
                                                                    h->d_arch_starts, h->d_lot_point);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_launch_ingest(tiktak_handle* h, int first_k, int n_k, int round_size, const double* pack, int* d_accepted) {
    IngestArgs g;
    g.first_k = first_k; g.n_k = n_k; g.nx = h->cfg.nx; g.world = h->cfg.world;
    g.n_max = (n_k + h->cfg.world - 1) / h->cfg.world;
    g.roundtrip = h->cfg.text_roundtrip;
    g.round_size = round_size < 1 ? n_k : round_size;
    g.accepted = d_accepted;
    g.pack = pack; g.out_f = h->d_out_f; g.out_x = h->d_out_x; g.out_evals = h->d_out_evals;
    g.res = h->d_res; g.valid = h->d_res_valid; g.arch_evals = h->d_arch_evals; g.best = h->d_best;
    ingest_kernel<<<1, 256, 0, h->stream>>>(g);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

/* how many restarts of the local-search kernel are resident at once on this GPU (one wave) */
bool tk_nm_use_w4(tiktak_handle* h);
int tk_nm_w4_capacity(tiktak_handle* h);
int tk_nm_wave_capacity(tiktak_handle* h) {
    const int nx = h->cfg.nx;
    int per_sm = 1;
    if (h->cfg.objective_id == TIKTAK_OBJ_SMM) return h->sm_count;
    if (tk_nm_use_w4(h)) return std::max(1, tk_nm_w4_capacity(h));
    tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        if (nx <= 31 && Obj::nterms(nx) <= nx) {
            const int ts = (nx + 2) & ~1;
            const size_t shmem = (size_t)(4 * ts + 8 + 4 * (5 * ts + (nx + 1) * nx)) * sizeof(double);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nm_quad_kernel<Obj, 0>, 128, shmem);
        } else if (nx <= 127 && Obj::nterms(nx) <= nx) {
            const int ts = (nx + 2) & ~1;
            const size_t shmem = (size_t)(8 * ts + 6 + (nx + 1) * nx + 4 * ts + 4 * 4 * ts) * sizeof(double);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nm_quad_wide_kernel<Obj, 2>, 128, shmem);
        }
        return TIKTAK_OK;
    });
    if (per_sm < 1) per_sm = 1;
    return per_sm * h->sm_count;
}

int tk_launch_pack(tiktak_handle* h, int n_k, double* pack) {
    const int n_max = (n_k + h->cfg.world - 1) / h->cfg.world;
    pack_kernel<<<(n_max + 127) / 128, 128, 0, h->stream>>>(n_k, h->cfg.nx, h->cfg.rank, h->cfg.world, h->d_out_f, h->d_out_x,
                                                            h->d_out_evals, pack);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_launch_shrink_width(tiktak_handle* h) {
    shrink_width_kernel<<<1, 256, 0, h->stream>>>(h->d_res, h->d_res_valid, h->cfg.maxpoints_plus1, h->cfg.nx, h->d_width,
                                                  h->d_wshr, h->d_shr_have);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

template <class Obj, int NXC>
static int nm_launch_warp(tiktak_handle* h, const TkTablesG& T, const NmArgs& a, size_t shmem) {
    TK_CUDA(cudaFuncSetAttribute(nm_quad_kernel<Obj, NXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    nm_quad_kernel<Obj, NXC><<<(unsigned)a.n_run, 128, shmem, h->stream>>>(T, h->sc, a);
    return TIKTAK_OK;
}

/* best row of the table by (value, virtual finish time, instance, k): getBestLine on the file an asynchronous run would have written */
__global__ void best_async_kernel(const double* res, const int* valid, const TkAsyncRow* ev, int nrows, int nx, double* best) {
    __shared__ double sf[256];
    __shared__ int st_[256], ss[256], sk[256], sc[256];
    double bf = 0.0; int bt = 0, bs = 0, bk = -1, cnt = 0;
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        if (!valid[r]) continue;
        cnt++;
        const double f = res[(size_t)r * (nx + 1)];
        const int t = r == 0 ? 0 : max(ev[r].fin, 0), s_ = r == 0 ? -1 : ev[r].slt;
        if (bk < 0 || nm_row_before(f, t, s_, r, bf, bt, bs, bk)) { bf = f; bt = t; bs = s_; bk = r; }
    }
    sf[threadIdx.x] = bf; st_[threadIdx.x] = bt; ss[threadIdx.x] = bs; sk[threadIdx.x] = bk; sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int i2 = threadIdx.x + o;
            if (sk[i2] >= 0 && (sk[threadIdx.x] < 0 || nm_row_before(sf[i2], st_[i2], ss[i2], sk[i2], sf[threadIdx.x], st_[threadIdx.x], ss[threadIdx.x], sk[threadIdx.x]))) {
                sf[threadIdx.x] = sf[i2]; st_[threadIdx.x] = st_[i2]; ss[threadIdx.x] = ss[i2]; sk[threadIdx.x] = sk[i2];
            }
            sc[threadIdx.x] += sc[i2];
        }
        __syncthreads();
    }
    const int b = sk[0];
    if (threadIdx.x == 0) { best[0] = sf[0]; best[1] = (double)b; best[2] = (double)sc[0]; }
    if (b >= 0)
        for (int d = threadIdx.x; d < nx; d += blockDim.x) best[3 + d] = res[(size_t)b * (nx + 1) + 1 + d];
}

template <class Obj, int NXC>
static int nm_launch_async(tiktak_handle* h, const TkTablesG& T, const NmArgs& a, NmAsync q, size_t shmem, int want, int* used) {
    auto kern = q.free_run ? nm_quad_async_kernel<Obj, NXC, true> : nm_quad_async_kernel<Obj, NXC, false>;
    TK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    int per_sm = 0;
    TK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, shmem));
    const int n = q.last_k - q.first_k + 1;
    const int P = std::max(1, std::min(std::min(want, n), per_sm * h->sm_count - 1)); /* every instance and the time keeper must be resident (they wait for one another) */
    int rc = tk_async_init(h, q, P);
    if (rc) return rc;
    void* args[] = {(void*)&T, (void*)&h->sc, (void*)&a, (void*)&q};
    TK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)P + 1), dim3(128), args, shmem, h->stream));
    h->launches++;
    *used = P;
    return TIKTAK_OK;
}

int tk_launch_nm_vec(tiktak_handle* h, const NmArgs& a);
bool tk_nm_w4_supported(tiktak_handle* h);
int tk_nm_w4_capacity(tiktak_handle* h);
int tk_launch_nm_w4(tiktak_handle* h, const NmArgs& a, const NmQueue& q);

/* which Nelder-Mead form runs: one warp per restart + device queue (kernels_nm_w4.cu) wherever it is built for the
 * configuration; TIKTAK_NM_FORM=quad forces the four-warp form of this file (experiments, cross-check of one form against the other) */
bool tk_nm_use_w4(tiktak_handle* h) {
    if (h->nm_form < 0) {
        const char* e = getenv("TIKTAK_NM_FORM");
        /* measured (B200, round 2): four warps per restart are ~2x faster per evaluation than one warp where the quad kernel
         * exists (nx <= 31); above that the one-warp kernel wins (C4: 97 ms against 254 ms) */
        const bool want_w4 = (e && !strcmp(e, "w4")) || (!(e && !strcmp(e, "quad")) && h->cfg.nx > 31);
        h->nm_form = (want_w4 && tk_nm_w4_supported(h)) ? 1 : 0;
    }
    return h->nm_form == 1;
}

/* commit_round_size > 0 (world == 1, throughput form): the kernel itself commits complete rounds into the results
 * table and stops the launch at the first round that changes the best row; h->d_nmq[NMQ_ACCEPTED] = restarts committed */
int tk_launch_nm(tiktak_handle* h, int first_k, int n_k, int commit_round_size) {
    const int nx = h->cfg.nx, np = nx + 1;
    NmArgs a;
    a.n_k = n_k;
    a.rank = h->cfg.rank;
    a.world = h->cfg.world;
    a.n_run = (n_k > a.rank) ? (n_k - a.rank + a.world - 1) / a.world : 0;
    if (a.n_run <= 0) return TIKTAK_OK;
    a.first_k = first_k;
    a.K = h->cfg.maxpoints_plus1;
    a.starts = h->d_starts;
    a.width = h->d_width;
    a.wshr = h->cfg.nm_shrink_mode == 1 ? h->d_wshr : nullptr;
    a.shr_have = h->cfg.nm_shrink_mode == 1 ? h->d_shr_have : nullptr;
    a.simplex = h->d_simplex;
    a.scale = h->nm_scale;
    a.ftol = h->cfg.nm_ftol;
    a.itmax = h->cfg.nm_itmax;
    const double ndim = (double)nx;
    a.fac1r = (1.0 - (-1.0)) / ndim; a.fac2r = a.fac1r - (-1.0);
    a.fac1e = (1.0 - 2.0) / ndim;    a.fac2e = a.fac1e - 2.0;
    a.fac1c = (1.0 - 0.5) / ndim;    a.fac2c = a.fac1c - 0.5;
    a.tiny = (double)1.0e-10f; /* TINY=1.0e-10 is a single-precision literal (minimize.f90:28) */
    a.out_x = h->d_out_x;
    a.out_f = h->d_out_f;
    a.out_evals = h->d_out_evals;
    if (h->cfg.objective_id == TIKTAK_OBJ_SMM) { /* vector-valued dfovec: one restart per block, block-wide evaluations (kernels_dfnls.cu) */
        if (commit_round_size > 0) { tk_set_error("nm: in-kernel commit needs nm_w4_kernel"); return TIKTAK_ERR_STATE; }
        return tk_launch_nm_vec(h, a);
    }
    if (tk_nm_use_w4(h)) {
        NmQueue q;
        const int K = h->cfg.maxpoints_plus1;
        if (!h->d_nmq) TK_CUDA(cudaMalloc(&h->d_nmq, (size_t)(NMQ_DONE0 + K + 1) * sizeof(int)));
        q.commit = (commit_round_size > 0 && h->cfg.world == 1) ? 1 : 0;
        q.round_size = q.commit ? commit_round_size : n_k;
        q.n_rounds = (n_k + q.round_size - 1) / q.round_size;
        TK_CUDA(cudaMemsetAsync(h->d_nmq, 0, (size_t)(NMQ_DONE0 + q.n_rounds) * sizeof(int), h->stream));
        /* restarts that may run ahead of the committed rounds: a round, or what two warps per SM sub-partition hold */
        q.depth0 = std::max(q.round_size, h->sm_count * 8);
        q.roundtrip = h->cfg.text_roundtrip;
        q.nx = nx;
        q.ctl = h->d_nmq; q.res = h->d_res; q.valid = h->d_res_valid; q.arch_evals = h->d_arch_evals; q.best = h->d_best;
        return tk_launch_nm_w4(h, a, q);
    }
    if (commit_round_size > 0) { tk_set_error("nm: in-kernel commit needs nm_w4_kernel"); return TIKTAK_ERR_STATE; }
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    return tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        int rc = TIKTAK_OK;
        if (nx <= 31 && Obj::nterms(nx) <= nx) { /* one restart per block of four warps */
            const int ts = (nx + 2) & ~1;
            const size_t shmem = (size_t)(4 * ts + 8 + 4 * (5 * ts + np * nx)) * sizeof(double);
            switch (nx) {
                case 4: rc = nm_launch_warp<Obj, 4>(h, T, a, shmem); break;
                case 10: rc = nm_launch_warp<Obj, 10>(h, T, a, shmem); break;
                case 20: rc = nm_launch_warp<Obj, 20>(h, T, a, shmem); break;
                default: rc = nm_launch_warp<Obj, 0>(h, T, a, shmem); break;
            }
        } else if (nx <= 127 && Obj::nterms(nx) <= nx) { /* four warps per restart, CJ coordinates per lane */
            const int ts = (nx + 2) & ~1;
            const size_t shmem = (size_t)(8 * ts + 6 + np * nx + 4 * ts + 4 * 4 * ts) * sizeof(double);
            const int cj = (np + 31) / 32;
            auto go = [&]<int CJ>() -> int {
                TK_CUDA(cudaFuncSetAttribute(nm_quad_wide_kernel<Obj, CJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                TK_CUDA(cudaFuncSetAttribute(nm_quad_wide_kernel<Obj, CJ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                nm_quad_wide_kernel<Obj, CJ><<<(unsigned)a.n_run, 128, shmem, h->stream>>>(T, h->sc, a);
                return TIKTAK_OK;
            };
            rc = cj == 2 ? go.template operator()<2>() : cj == 3 ? go.template operator()<3>() : go.template operator()<4>();
        } else {
            const size_t per_warp = (size_t)(np * nx + np + nx + 4 * nx + 4 + 8 * nx) * sizeof(double);
            const size_t per_block = (size_t)3 * nx * sizeof(double);
            int wpb = 4;
            while (wpb > 1 && per_warp * wpb + per_block > 200 * 1024) wpb >>= 1;
            if (per_warp * wpb + per_block > 227 * 1024) {
                tk_set_error("nx=%d: Nelder-Mead simplex does not fit in shared memory", nx);
                return TIKTAK_ERR_UNSUPPORTED;
            }
            /* spread the restarts over the SMs before stacking warps in a block */
            while (wpb > 1 && (a.n_run + wpb - 1) / wpb < h->sm_count) wpb >>= 1;
            const unsigned grid = (unsigned)((a.n_run + wpb - 1) / wpb);
            TK_CUDA(cudaFuncSetAttribute(nm_kernel<Obj>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            nm_kernel<Obj><<<grid, wpb * 32, per_warp * wpb + per_block, h->stream>>>(T, h->sc, a);
        }
        if (rc) return rc;
        h->launches++;
        TK_CUDA(cudaGetLastError());
        return TIKTAK_OK;
    });
}


/* Asynchronous instances (nm_quad_async_kernel): restarts first_k .. first_k + n_k - 1 by min(instances, resident blocks) instances;
 * rows, evaluation counts, searchStart rows and the best row are on the device when the stream has drained. */
int tk_launch_nm_w4_async(tiktak_handle* h, const NmArgs& a, const NmAsync& aq, int want, int* used);

/* which instance runs: -1 = none built for the configuration; 0 = a block of four warps (nm_quad_async_kernel, nx <= 31), tick = pass
 * of amoeba's loop; 1 = one warp (nm_w4_kernel), tick = objective evaluation.  The four-warp block is the latency form: it wins while
 * the run has fewer restarts in flight than the GPU has room for; with every SM full the one-warp form spends less than half the
 * instructions per counted evaluation and the GPU holds 3.3 times as many instances (C3: 887 blocks 19.4 ms, 2956 warps
 * 14.8 ms).  So: one warp above nx = 31 and whenever the caller asks for at least twice the blocks the four-warp form keeps resident;
 * TIKTAK_ASYNC_FORM=quad|w4 overrides. */
int tk_nm_async_form(tiktak_handle* h, int instances) {
    const int nx = h->cfg.nx;
    if (h->cfg.world != 1 || h->cfg.search_type != 0 || h->cfg.nm_shrink_mode != 0 || h->cfg.objective_id == TIKTAK_OBJ_SMM) return -1;
    bool scalar = false;
    tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int { scalar = Obj::nterms(nx) <= nx; return 0; });
    if (!scalar) return -1;
    const bool quad_ok = nx <= 31, w4_ok = tk_nm_w4_supported(h);
    if (!quad_ok && !w4_ok) return -1;
    const char* e = getenv("TIKTAK_ASYNC_FORM");
    if (e && !strcmp(e, "quad") && quad_ok) return 0;
    if (e && !strcmp(e, "w4") && w4_ok) return 1;
    if (!quad_ok) return 1;
    return (w4_ok && instances >= 12 * h->sm_count /* twice the six blocks per SM */) ? 1 : 0;
}
bool tk_nm_async_supported(tiktak_handle* h) { return tk_nm_async_form(h, 0) >= 0; }

int tk_launch_nm_async(tiktak_handle* h, int first_k, int n_k, int instances, int* used, bool free_run) {
    const int nx = h->cfg.nx, np = nx + 1, K = h->cfg.maxpoints_plus1;
    int form = tk_nm_async_form(h, instances);
    if (form < 0) { tk_set_error("local_async: Nelder-Mead with a scalar objective, nx <= 127, search_type 0, nm_shrink_mode 0, one GPU"); return TIKTAK_ERR_UNSUPPORTED; }
    NmArgs a;
    a.n_k = n_k; a.rank = 0; a.world = 1; a.n_run = n_k; a.first_k = first_k; a.K = K;
    a.starts = nullptr; a.width = h->d_width; a.wshr = nullptr; a.shr_have = nullptr;
    a.simplex = h->d_simplex; a.scale = h->nm_scale; a.ftol = h->cfg.nm_ftol; a.itmax = h->cfg.nm_itmax;
    const double ndim = (double)nx;
    a.fac1r = (1.0 - (-1.0)) / ndim; a.fac2r = a.fac1r - (-1.0);
    a.fac1e = (1.0 - 2.0) / ndim;    a.fac2e = a.fac1e - 2.0;
    a.fac1c = (1.0 - 0.5) / ndim;    a.fac2c = a.fac1c - 0.5;
    a.tiny = (double)1.0e-10f;
    a.out_x = nullptr; a.out_f = nullptr; a.out_evals = nullptr;
    NmAsync q;
    int rc0 = tk_async_setup(h, first_k, n_k, form == 1 ? np + 1 : (np + 3) / 4, &q); /* (one-warp form: c0 = the least a restart evaluates) */
    if (rc0) return rc0;
    q.free_run = free_run ? 1 : 0;
    if (form == 1) {
        if (!getenv("TIKTAK_ASYNC_CB")) q.cb = TIKTAK_ASYNC_PICKUP_TICKS_NM_EVALS;
        int rc = tk_launch_nm_w4_async(h, a, q, instances, used);
        if (rc) return rc;
        TK_CUDA(cudaGetLastError());
        return tk_async_best(h);
    }
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    const int ts = (nx + 2) & ~1;
    const size_t shmem = (size_t)(4 * ts + 8 + ts + 4 * (5 * ts + np * nx)) * sizeof(double);
    int rc = tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        switch (nx) {
            case 4: return nm_launch_async<Obj, 4>(h, T, a, q, shmem, instances, used);
            case 10: return nm_launch_async<Obj, 10>(h, T, a, q, shmem, instances, used);
            case 20: return nm_launch_async<Obj, 20>(h, T, a, q, shmem, instances, used);
            default: return nm_launch_async<Obj, 0>(h, T, a, q, shmem, instances, used);
        }
    });
    if (rc) return rc;
    TK_CUDA(cudaGetLastError());
    return tk_async_best(h);
}

/* the bookkeeping both asynchronous kernels (Nelder-Mead here, DFNLS in kernels_dfnls.cu) share */
int tk_async_setup(tiktak_handle* h, int first_k, int n_k, int c0, NmAsync* q) {
    const int K = h->cfg.maxpoints_plus1;
    if (!h->d_async_ev) TK_CUDA(cudaMalloc(&h->d_async_ev, (size_t)(K + 1) * sizeof(TkAsyncRow)));
    if (!h->d_async_clk) TK_CUDA(cudaMalloc(&h->d_async_clk, (size_t)64 * 1024 * sizeof(int)));
    q->first_k = first_k; q->last_k = first_k + n_k - 1; q->K = K; q->nx = h->cfg.nx; q->roundtrip = h->cfg.text_roundtrip;
    q->xs = h->d_xstarts; q->w11 = h->d_w11; q->res = h->d_res; q->valid = h->d_res_valid; q->arch_evals = h->d_arch_evals;
    q->arch_starts = h->d_arch_starts; q->ev = h->d_async_ev; q->clk = h->d_async_clk;
    q->err = h->d_async_clk + 60 * 1024; /* (the clock array is 64 K ints: far behind the clocks and the time keeper's words) */
    q->dbg = nullptr;
    if (getenv("TIKTAK_ASYNC_DEBUG")) { static long long* d = nullptr; if (!d) cudaMalloc(&d, 5 * 2048 * sizeof(long long)); q->dbg = d; h->async_dbg = d; }
    if (!h->d_async_base) TK_CUDA(cudaMalloc(&h->d_async_base, (size_t)(K + 1) * sizeof(int)));
    q->base = h->d_async_base;
    q->free_run = 0;
    q->c0 = c0; q->cb = TIKTAK_ASYNC_PICKUP_TICKS;
    if (const char* e = getenv("TIKTAK_ASYNC_CB")) q->cb = atoi(e); /* experiment */
    return TIKTAK_OK;
}
int tk_async_init(tiktak_handle* h, const NmAsync& q, int P) {
    nm_async_init_kernel<<<(std::max(q.K + 1, P + 1) + 255) / 256, 256, 0, h->stream>>>(q, P);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}
int tk_async_best(tiktak_handle* h) {
    best_async_kernel<<<1, 256, 0, h->stream>>>(h->d_res, h->d_res_valid, h->d_async_ev, h->cfg.maxpoints_plus1 + 1, h->cfg.nx, h->d_best);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}
