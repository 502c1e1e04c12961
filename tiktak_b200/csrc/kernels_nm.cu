/* kernels_nm.cu -- batched warm-start Nelder-Mead: one restart per warp.
 *
 * Replaces, per restart k: getBestLine/getModifiedParam (TiktakGlobalSearch.f90:683-720, 781-823: the
 * theta-blend toward the best local minimum so far, which the reference obtains by re-reading and
 * re-sorting searchResults.dat), runAmoeba (:602-633: regular simplex from simplex_coordinates2 scaled
 * by range/K^(1/nx), clamped to the bounds, nx+1 evaluations) and amoeba/amotry (minimize.f90:7-135).
 *
 * Mapping: a warp owns a restart.  The simplex (nx+1 vertices x nx), y, psum live in the warp's slice
 * of shared memory; lanes parallelise the coordinate-wise vector updates, warp shuffles give
 * argmin/argmax (first occurrence, like MINLOC/MAXLOC) and the shrink re-evaluates its nx vertices on
 * nx lanes at once.  The sequential part of amoeba -- reflect, then expand OR contract -- is collapsed
 * into one evaluation phase per iteration by evaluating the four possible trial points (reflection R,
 * expansion from R, contraction after an accepted R, contraction after a rejected R) on four lanes at
 * once; which of them count, and the `iter` bookkeeping, follow minimize.f90:97-114 exactly, so the
 * trajectory is the sequential one.  All arithmetic is unfused, in the reference's order; psum is
 * updated incrementally (minimize.f90:130) and only re-summed after a shrink (:112).
 */
#include <float.h>

#include "tk_common.cuh"

namespace {

struct NmArgs {
    int n_run;            /* restarts this launch runs */
    int n_k;              /* restarts of the round (column length of starts/out_x) */
    int rank, world;      /* restart slot = rank + r*world */
    const double* starts; /* (n_k, nx) column-major */
    const double* width;  /* nx: p_range(:,2)-p_range(:,1), or (n_k, nx) column-major if per_restart_width */
    int per_restart_width;
    const double* simplex; /* (nx, nx+1) column-major unit simplex */
    double scale;          /* DBLE(p_maxpoints)**(1/nx) */
    double ftol;
    int itmax;
    double fac1r, fac2r, fac1e, fac2e, fac1c, fac2c; /* amotry's fac1/fac2 for fac = -1, 2, 0.5 */
    double tiny;
    double* out_x; /* (n_k, nx) column-major */
    double* out_f; /* n_k */
    int* out_evals;
};

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
    for (int e = lane; e < np * nx; e += 32) {
        const int ia = e / nx, i = e - ia * nx;
        const double wd = a.per_restart_width ? a.width[(size_t)i * a.n_k + slot] : a.width[i];
        double t = TK_ADD(a.starts[(size_t)i * a.n_k + slot], TK_DIV(TK_MUL(a.simplex[(size_t)ia * nx + i], wd), a.scale));
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
 * Lane-per-coordinate variant for nx <= 31 (one restart per warp, as above): lane j keeps psum_j and the
 * four trial coordinates of the iteration in registers, lane i keeps y_i, so argmin/argmax are two
 * warp REDUX + one ballot on the order-preserving integer image of y (first occurrence = lowest
 * lane), the per-coordinate `term`s of all four trial points are evaluated by lane j back to back
 * (4-way ILP on the cos chains) and only the in-order `fold` runs on one lane per trial point.
 * Same arithmetic and the same decisions as nm_kernel; ~5x lower latency per simplex iteration.
 * --------------------------------------------------------------------------------------------- */
__device__ __forceinline__ unsigned long long nm_order_key(double f) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(f);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ int warp_first_min_key(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return __ffs(__ballot_sync(0xffffffffu, hi == mh && lo == ml)) - 1;
}

template <class Obj>
__global__ void __launch_bounds__(128) nm_lane_kernel(const TkTablesG T, const TkScalars sc, const NmArgs a) {
    extern __shared__ double smem[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int nx = sc.nx, np = nx + 1;
    const int nt = Obj::nterms(nx);
    double* sbl = smem;
    double* sbu = sbl + nx;
    double* saux = sbu + nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    __syncthreads();
    const int r = blockIdx.x * wpb + warp;
    if (r >= a.n_run) return;
    const int slot = a.rank + r * a.world;
    const int per_warp = np * nx + 4 * nx + 8 * nx;
    double* p = smem + 3 * nx + (size_t)warp * per_warp; /* p[i*nx + j] */
    double* cand = p + np * nx;                           /* cand[c*nx + j] */
    double* ta = cand + 4 * nx;
    double* tb = ta + 4 * nx;
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, sbl, sbu, saux);
    const bool has_j = lane < nx, has_i = lane < np;
    const int j = has_j ? lane : 0;
    const double blj = sbl[j], buj = sbu[j];

    /* runAmoeba :622-627 */
    for (int e = lane; e < np * nx; e += 32) {
        const int ia = e / nx, i = e - ia * nx;
        const double wd = a.per_restart_width ? a.width[(size_t)i * a.n_k + slot] : a.width[i];
        double t = TK_ADD(a.starts[(size_t)i * a.n_k + slot], TK_DIV(TK_MUL(a.simplex[(size_t)ia * nx + i], wd), a.scale));
        t = fmax(fmin(t, sbu[i]), sbl[i]);
        p[e] = t;
    }
    __syncwarp();
    double yl = 0.0; /* y(lane) */
    if (has_i) yl = tk_objfun<Obj>(ctx, TkMemX{p + (size_t)lane * nx, 1});
    double psum = 0.0; /* psum(lane) */
    if (has_j)
        for (int i = 0; i < np; i++) psum = TK_ADD(psum, p[i * nx + j]);

    int item_c[4], item_i[4]; /* this lane's (trial point, term) items; c = 4 marks "none" */
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int e = lane + 32 * u;
        item_c[u] = (e < 4 * nt) ? e / nt : 4;
        item_i[u] = (e < 4 * nt) ? e - item_c[u] * nt : 0;
    }
    int iter = 0, ilo = 0;
    for (;;) {
        const unsigned long long key = nm_order_key(yl);
        ilo = warp_first_min_key(has_i ? key : ~0ULL);
        const int ihi = warp_first_min_key(has_i ? ~key : ~0ULL);
        const double ylo = __shfl_sync(FULL, yl, ilo), yhi = __shfl_sync(FULL, yl, ihi);
        const unsigned long long klo = nm_order_key(ylo);
        const int inhi = warp_first_min_key(has_i ? ~(lane == ihi ? klo : key) : ~0ULL);
        const double ynhi = __shfl_sync(FULL, yl, inhi);
        /* rtol < ftol (:84-85); the division is only formed when the product test is within 2^-40 */
        const double num = TK_MUL(2.0, fabs(TK_SUB(yhi, ylo)));
        const double den = TK_ADD(TK_ADD(fabs(yhi), fabs(ylo)), a.tiny);
        const double thr = TK_MUL(a.ftol, den);
        bool conv;
        if (num < TK_MUL(thr, 0.99999999999909051)) conv = true;
        else if (num > TK_MUL(thr, 1.00000000000090949)) conv = false;
        else conv = TK_DIV(num, den) < a.ftol;
        if (conv || iter >= a.itmax) break;

        /* trial points (amotry :124-126, unfused) */
        double ph = has_j ? p[ihi * nx + j] : 0.0;
        const double R = TK_SUB(TK_MUL(psum, a.fac1r), TK_MUL(ph, a.fac2r));
        const double ps1 = TK_ADD(TK_SUB(psum, ph), R);
        const double E = TK_SUB(TK_MUL(ps1, a.fac1e), TK_MUL(R, a.fac2e));
        const double Co = TK_SUB(TK_MUL(ps1, a.fac1c), TK_MUL(R, a.fac2c));
        const double Ci = TK_SUB(TK_MUL(psum, a.fac1c), TK_MUL(ph, a.fac2c));
        unsigned oob = 0;
        oob |= __ballot_sync(FULL, has_j && (R < blj || R > buj)) ? 1u : 0u;
        oob |= __ballot_sync(FULL, has_j && (E < blj || E > buj)) ? 2u : 0u;
        oob |= __ballot_sync(FULL, has_j && (Co < blj || Co > buj)) ? 4u : 0u;
        oob |= __ballot_sync(FULL, has_j && (Ci < blj || Ci > buj)) ? 8u : 0u;
        if (has_j) { cand[j] = R; cand[nx + j] = E; cand[2 * nx + j] = Co; cand[3 * nx + j] = Ci; }
        __syncwarp();
        /* the 4*nt independent terms, dealt over all 32 lanes (item e = lane + 32u = c*nt + i); the loop is
         * unrolled so that the (latency-bound) term chains of one lane interleave */
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int c = item_c[u], i = item_i[u];
            if (c < 4 && !((oob >> c) & 1u)) {
                double ta_, tb_;
                Obj::term(ctx, TkMemX{cand + (size_t)c * nx, 1}, i, ta_, tb_);
                ta[c * nx + i] = ta_;
                tb[c * nx + i] = tb_;
            }
        }
        __syncwarp();
        double yc = 0.0;
        if (lane < 4) {
            if ((oob >> lane) & 1u) {
                yc = tk_objfun<Obj>(ctx, TkMemX{cand + (size_t)lane * nx, 1});
            } else {
                double s_, p_;
                Obj::fold_init(ctx, s_, p_);
                const double* qa = ta + lane * nx;
                const double* qb = tb + lane * nx;
                int i = 0;
                for (; i + 4 <= nt; i += 4) { /* loads first, then the in-order chain */
                    const double a0 = qa[i], a1 = qa[i + 1], a2 = qa[i + 2], a3 = qa[i + 3];
                    const double b0 = qb[i], b1 = qb[i + 1], b2 = qb[i + 2], b3 = qb[i + 3];
                    Obj::fold(s_, p_, a0, b0); Obj::fold(s_, p_, a1, b1);
                    Obj::fold(s_, p_, a2, b2); Obj::fold(s_, p_, a3, b3);
                }
                for (; i < nt; i++) Obj::fold(s_, p_, qa[i], qb[i]);
                yc = tk_objfun_finish(ctx, Obj::finish(ctx, s_, p_), 0.0);
            }
        }
        const double yR = __shfl_sync(FULL, yc, 0);
        const bool accR = yR < yhi; /* :128 */
        if (accR) {
            psum = ps1;
            if (has_j) p[ihi * nx + j] = R;
            ph = R;
            if (lane == ihi) yl = yR;
        }
        iter++;
        if (yR <= ylo) { /* :99 expansion */
            double yE = __shfl_sync(FULL, yc, 1);
            double Ev = E;
            if (!accR) { /* unreachable unless all y are equal; evaluated from the live state */
                Ev = TK_SUB(TK_MUL(psum, a.fac1e), TK_MUL(ph, a.fac2e));
                if (has_j) cand[nx + j] = Ev;
                __syncwarp();
                double t_ = 0.0;
                if (lane == 0) t_ = tk_objfun<Obj>(ctx, TkMemX{cand + nx, 1});
                yE = __shfl_sync(FULL, t_, 0);
            }
            iter++;
            const double ycur = accR ? yR : yhi;
            if (yE < ycur) {
                psum = TK_ADD(TK_SUB(psum, ph), Ev);
                if (has_j) p[ihi * nx + j] = Ev;
                if (lane == ihi) yl = yE;
            }
        } else if (yR >= ynhi) { /* :102 contraction */
            const double ysave = accR ? yR : yhi;
            const double yC = __shfl_sync(FULL, yc, accR ? 2 : 3);
            const double Cv = accR ? Co : Ci;
            iter++;
            if (yC < ysave) {
                psum = TK_ADD(TK_SUB(psum, ph), Cv);
                if (has_j) p[ihi * nx + j] = Cv;
                if (lane == ihi) yl = yC;
            } else { /* :106-113 shrink */
                __syncwarp();
                if (has_j) {
                    const double pl = p[ilo * nx + j];
                    for (int i = 0; i < np; i++)
                        if (i != ilo) p[i * nx + j] = TK_MUL(0.5, TK_ADD(p[i * nx + j], pl));
                }
                __syncwarp();
                if (has_i && lane != ilo) yl = tk_objfun<Obj>(ctx, TkMemX{p + (size_t)lane * nx, 1});
                iter += nx;
                psum = 0.0;
                if (has_j)
                    for (int i = 0; i < np; i++) psum = TK_ADD(psum, p[i * nx + j]);
            }
        }
        __syncwarp();
    }
    if (has_j) a.out_x[(size_t)j * a.n_k + slot] = p[ilo * nx + j];
    if (lane == ilo) {
        a.out_f[slot] = yl;
        a.out_evals[slot] = np + iter + 1;
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
__global__ void blend_kernel(const double* xs, int K, int nx, int first_k, int n_k, const double* w11,
                             const double* best, double* starts) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_k * nx) return;
    const int d = e / n_k, s = e - d * n_k, k = first_k + s; /* k is 1-based */
    const double ev = xs[(size_t)(d + 1) * K + (k - 1)];
    double out = ev;
    if (best[2] > 1.0) { /* count <= 1: no finished local minimisation yet (:701) */
        const double w = w11[k];
        out = TK_ADD(TK_MUL(w, best[3 + d]), TK_MUL(TK_SUB(1.0, w), ev));
    }
    starts[(size_t)d * n_k + s] = out;
}

} /* namespace */

int tk_launch_best(tiktak_handle* h) {
    best_kernel<<<1, 256, 0, h->stream>>>(h->d_res, h->d_res_valid, h->cfg.maxpoints_plus1 + 1, h->cfg.nx, h->d_best);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_launch_blend(tiktak_handle* h, int first_k, int n_k) {
    const int tot = n_k * h->cfg.nx;
    blend_kernel<<<(tot + 127) / 128, 128, 0, h->stream>>>(h->d_xstarts, h->cfg.maxpoints_plus1, h->cfg.nx, first_k, n_k,
                                                            h->d_w11, h->d_best, h->d_starts);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_launch_nm(tiktak_handle* h, int first_k, int n_k) {
    const int nx = h->cfg.nx, np = nx + 1;
    NmArgs a;
    a.n_k = n_k;
    a.rank = h->cfg.rank;
    a.world = h->cfg.world;
    a.n_run = (n_k > a.rank) ? (n_k - a.rank + a.world - 1) / a.world : 0;
    if (a.n_run <= 0) return TIKTAK_OK;
    a.starts = h->d_starts;
    a.width = h->use_round_width ? h->d_width_round : h->d_width;
    a.per_restart_width = h->use_round_width ? 1 : 0;
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
    (void)first_k;
    const bool lane_variant = nx <= 31;
    const size_t per_warp = (size_t)(lane_variant ? (np * nx + 12 * nx) : (np * nx + np + nx + 4 * nx + 4 + 8 * nx)) * sizeof(double);
    const size_t per_block = (size_t)3 * nx * sizeof(double);
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb + per_block > 200 * 1024) wpb >>= 1;
    if (per_warp * wpb + per_block > 227 * 1024) {
        tk_set_error("nx=%d: Nelder-Mead simplex does not fit in shared memory", nx);
        return TIKTAK_ERR_UNSUPPORTED;
    }
    /* spread the restarts over the SMs before stacking warps in a block */
    while (wpb > 1 && (a.n_run + wpb - 1) / wpb < h->sm_count) wpb >>= 1;
    const size_t shmem = per_warp * wpb + per_block;
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    return tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        const unsigned grid = (unsigned)((a.n_run + wpb - 1) / wpb);
        if (lane_variant) {
            TK_CUDA(cudaFuncSetAttribute(nm_lane_kernel<Obj>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            nm_lane_kernel<Obj><<<grid, wpb * 32, shmem, h->stream>>>(T, h->sc, a);
        } else {
            TK_CUDA(cudaFuncSetAttribute(nm_kernel<Obj>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            nm_kernel<Obj><<<grid, wpb * 32, shmem, h->stream>>>(T, h->sc, a);
        }
        h->launches++;
        TK_CUDA(cudaGetLastError());
        return TIKTAK_OK;
    });
}
