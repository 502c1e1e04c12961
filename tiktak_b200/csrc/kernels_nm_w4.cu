/* kernels_nm_w4.cu -- Nelder-Mead, one restart per WARP, restarts handed out by a device-side queue.
 *
 * Same reference code as kernels_nm.cu -- runAmoeba (TiktakGlobalSearch.f90:602-633), amoeba/amotry
 * (minimize.f90:7-135), the bookkeeping of completeSearch (:582-597) -- and the same arithmetic and decisions, so a
 * restart's trajectory, minimum and evaluation count are bit for bit those of nm_quad_kernel and of the CPU oracle.
 *
 * The local stage is a sequence of rounds (a round's starts depend on the best row of the rounds before it), a round
 * holds ~1000 restarts and lasts as long as its longest restart, so what counts is the LATENCY of one restart with
 * ~1000 of them in flight -- not instructions per evaluation (measured: a form with one evaluation per step and two
 * restarts per warp needed 1850 cycles per evaluation; nm_quad_kernel needs 930 per iteration of ~1.6 evaluations, but
 * spends four warps, a block barrier and a shared-memory exchange on it and holds only 888 restarts).  This kernel
 * keeps nm_quad_kernel's idea -- the four trial points an iteration can need (reflection R, expansion from R,
 * contraction after an accepted R, contraction after a rejected R) are evaluated AT ONCE, and minimize.f90:97-114
 * decides which of them count -- inside ONE warp:
 *   - the warp is four sub-groups of eight lanes, sub-group c evaluates candidate c; lane l of a sub-group owns the
 *     coordinates j = l + 8k (k < CJ) and the sorted-list positions e = l + 8k; psum, the worst vertex and the
 *     sorted (y, vertex) list are held redundantly by the four sub-groups, so every lane takes the same decisions and
 *     nothing but the four values is exchanged (four shuffles; no barrier, no shared-memory round trip);
 *   - `term` j is computed by the lane that owns j, the first lane of each sub-group folds its candidate's terms in
 *     index order (the roundings of a single-thread evaluation, tiktak_obj.cuh) -- the four folds run in the same
 *     instructions;
 *   - vertices sorted by (y ascending, vertex index descending): ylo, yhi, ynhi, ihi are lane reads, replacing the
 *     worst vertex is one ballot + shuffle insertion.
 * A warp that finishes takes the next restart from a ticket counter.  world == 1: the warp that finishes the last
 * restart of a round commits the complete rounds, in order, into the device-resident searchResults table
 * (ingest_kernel's work); a committed round that changes the best row stops the launch -- the later rounds in it were
 * started early against the old row -- and the host queues the rest again (NmQueue, nm_common.cuh).  How far tickets
 * may run ahead of the committed rounds is bounded, and the bound doubles with every round committed without a change.
 */
#include <float.h>
#include <stdlib.h>

#include <algorithm>

#include "nm_common.cuh"

using namespace tknm;

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <class O>
__host__ __device__ constexpr bool nm_uses_b() { /* plugins may declare kFoldUsesB = false: `fold` ignores b */
    if constexpr (requires { O::kFoldUsesB; }) return O::kFoldUsesB;
    else return true;
}

struct TkLaneX1 { /* accessor for plugins whose term i reads x[i] only */
    double v;
    __device__ __forceinline__ double operator[](int) const { return v; }
};

/* per-warp shared memory */
struct NmWarpMem {
    double* ta;   /* [4][ts] per-term partials a_i of the four candidates; also sort scratch */
    double* tb;   /* [4][ts] per-term partials b_i */
    double* pn;   /* [4][ts] per-coordinate penalty terms (out-of-bounds points only) */
    double* cand; /* [4][ts] the candidate points, for plugins whose terms read neighbouring coordinates */
    double* yv;   /* [ts] y by vertex (initial simplex, shrink) */
    double* p;    /* [np*nx] the simplex, p[v*nx + j] */
    int ts;
};

/* objFun (objective.f90:125-142) at four points at once: sub-group sg evaluates the point whose coordinate j = l8 + 8k
 * sits in x[k] of its lane l8; every lane of the sub-group returns its point's value.  Lane i of the sub-group evaluates
 * `term` i, its first lane folds them in index order: the roundings of a single-thread evaluation.  A point outside
 * the bounds takes the literal getPenalty sum (:51-74) and dfovec's penalty branch (:92-95). */
template <class Obj, int CJ, int NXC, int LP = 8>
__device__ __forceinline__ double nm_w4_eval(const tk_obj_ctx& ctx, const NmWarpMem& m, const double (&x)[CJ], const double* sbl,
                                             const double* sbu, int sg, int l8, int nx, int nt) {
    double* ta = m.ta + sg * m.ts;
    double* tb = m.tb + sg * m.ts;
    bool ob = false;
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int j = l8 + LP * k;
        if (j < nx) ob = ob || (x[k] < sbl[j] || x[k] > sbu[j]);
    }
    const unsigned obm = __ballot_sync(FULL, ob);
    const bool oob = ((obm >> ((LP * sg) & 31)) & (LP == 32 ? 0xffffffffu : ((1u << (LP & 31)) - 1u))) != 0u;
    if (!Obj::kTermLocal) {
#pragma unroll
        for (int k = 0; k < CJ; k++)
            if (l8 + LP * k < nx) m.cand[sg * m.ts + l8 + LP * k] = x[k];
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int i = l8 + LP * k;
        if (i < nt) {
            double ta_, tb_;
            if (Obj::kTermLocal) Obj::term(ctx, TkLaneX1{x[k]}, i, ta_, tb_);
            else Obj::term(ctx, TkMemX{m.cand + sg * m.ts, 1}, i, ta_, tb_);
            ta[i] = ta_;
            if (nm_uses_b<Obj>()) tb[i] = tb_;
        }
    }
    if (obm != 0u && oob) { /* rare: getPenalty :57-66, coordinate by coordinate */
#pragma unroll
        for (int k = 0; k < CJ; k++) {
            const int j = l8 + LP * k;
            if (j < nx) {
                const double blj = sbl[j], buj = sbu[j];
                const double a = TK_DIV(fmax(0.0, TK_SUB(blj, x[k])), fmax(0.1, fabs(blj)));
                const double b = TK_DIV(fmax(0.0, TK_SUB(x[k], buj)), fmax(0.1, fabs(buj)));
                m.pn[sg * m.ts + j] = TK_ADD(TK_MUL(ctx.penaltyc, TK_MUL(a, a)), TK_MUL(ctx.penaltyc, TK_MUL(b, b)));
            }
        }
    }
    __syncwarp();
    double yc = 0.0;
    if (l8 == 0) {
        double pen = 0.0;
        if (oob) {
            double penalty = 0.0;
            for (int i = 0; i < nx; i++) penalty = TK_ADD(penalty, m.pn[sg * m.ts + i]);
            if (penalty > TK_MYEPS) pen = fmin(ctx.max_penalty, TK_ADD(ctx.penalty0, penalty));
        }
        double f;
        if (pen > TK_MYEPS) {
            f = TK_DIV(pen, ctx.sqrt_nm); /* objective.f90:93-94 */
        } else {
            double s_, p_;
            Obj::fold_init(ctx, s_, p_);
            const double2* qa = reinterpret_cast<const double2*>(ta);
            const double2* qb = reinterpret_cast<const double2*>(tb);
            if (NXC > 0 && NXC <= 20) { /* loads first, then the in-order chain */
                constexpr int H = (NXC > 0 && NXC <= 20) ? (NXC + 1) / 2 : 1;
                double2 va[H + 1], vb[H + 1];
#pragma unroll
                for (int i = 0; i < H; i++) {
                    va[i] = qa[i];
                    vb[i] = nm_uses_b<Obj>() ? qb[i] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int i = 0; i < H; i++) {
                    if (2 * i < nt) Obj::fold(s_, p_, va[i].x, vb[i].x);
                    if (2 * i + 1 < nt) Obj::fold(s_, p_, va[i].y, vb[i].y);
                }
            } else {
                int i = 0;
                for (; i + 8 <= nt; i += 8) {
                    const double2 a0 = qa[(i >> 1)], a1 = qa[(i >> 1) + 1], a2 = qa[(i >> 1) + 2], a3 = qa[(i >> 1) + 3];
                    double2 b0 = make_double2(0.0, 0.0), b1 = b0, b2 = b0, b3 = b0;
                    if (nm_uses_b<Obj>()) { b0 = qb[(i >> 1)]; b1 = qb[(i >> 1) + 1]; b2 = qb[(i >> 1) + 2]; b3 = qb[(i >> 1) + 3]; }
                    Obj::fold(s_, p_, a0.x, b0.x); Obj::fold(s_, p_, a0.y, b0.y);
                    Obj::fold(s_, p_, a1.x, b1.x); Obj::fold(s_, p_, a1.y, b1.y);
                    Obj::fold(s_, p_, a2.x, b2.x); Obj::fold(s_, p_, a2.y, b2.y);
                    Obj::fold(s_, p_, a3.x, b3.x); Obj::fold(s_, p_, a3.y, b3.y);
                }
                for (; i < nt; i++) Obj::fold(s_, p_, ta[i], nm_uses_b<Obj>() ? tb[i] : 0.0);
            }
            f = Obj::finish(ctx, s_, p_);
        }
        yc = tk_objfun_finish(ctx, f, pen);
    }
    return __shfl_sync(FULL, yc, 0, LP); /* the sub-group's value; also orders the scratch reads before the next writes */
}

/* the same objFun at FOUR points at once with all 32 lanes on the coordinates (j = lane + 32k): the four points' terms are
 * independent instruction chains of one lane (the in-order warp overlaps their latencies), lanes 0..3 fold one point each.
 * y[c] = value of point c, on every lane. */
template <class Obj, int CJ, int NXC>
__device__ __forceinline__ void nm_w1_eval4(const tk_obj_ctx& ctx, const NmWarpMem& m, const double (&x)[4][CJ], const double* sbl,
                                            const double* sbu, int lane, int nx, int nt, double (&y)[4]) {
    unsigned ob = 0u;
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int j = lane + 32 * k;
        if (j < nx) {
            const double blj = sbl[j], buj = sbu[j];
#pragma unroll
            for (int c = 0; c < 4; c++) ob |= (x[c][k] < blj || x[c][k] > buj) ? (1u << c) : 0u;
        }
    }
    const unsigned oobm = __reduce_or_sync(FULL, ob);
    if (!Obj::kTermLocal) {
#pragma unroll
        for (int k = 0; k < CJ; k++)
            if (lane + 32 * k < nx) {
#pragma unroll
                for (int c = 0; c < 4; c++) m.cand[c * m.ts + lane + 32 * k] = x[c][k];
            }
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int i = lane + 32 * k;
        if (i < nt) {
            double ta_[4], tb_[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (Obj::kTermLocal) Obj::term(ctx, TkLaneX1{x[c][k]}, i, ta_[c], tb_[c]);
                else Obj::term(ctx, TkMemX{m.cand + c * m.ts, 1}, i, ta_[c], tb_[c]);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                m.ta[c * m.ts + i] = ta_[c];
                if (nm_uses_b<Obj>()) m.tb[c * m.ts + i] = tb_[c];
            }
        }
    }
    if (oobm != 0u) { /* rare: getPenalty :57-66, coordinate by coordinate, for the points outside the box */
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (!((oobm >> c) & 1u)) continue;
#pragma unroll
            for (int k = 0; k < CJ; k++) {
                const int j = lane + 32 * k;
                if (j < nx) {
                    const double blj = sbl[j], buj = sbu[j];
                    const double a = TK_DIV(fmax(0.0, TK_SUB(blj, x[c][k])), fmax(0.1, fabs(blj)));
                    const double b = TK_DIV(fmax(0.0, TK_SUB(x[c][k], buj)), fmax(0.1, fabs(buj)));
                    m.pn[c * m.ts + j] = TK_ADD(TK_MUL(ctx.penaltyc, TK_MUL(a, a)), TK_MUL(ctx.penaltyc, TK_MUL(b, b)));
                }
            }
        }
    }
    __syncwarp();
    double yc = 0.0;
    if (lane < 4) {
        const double* ta = m.ta + lane * m.ts;
        const double* tb = m.tb + lane * m.ts;
        const bool oob = (oobm >> lane) & 1u;
        double pen = 0.0;
        if (oob) {
            double penalty = 0.0;
            for (int i = 0; i < nx; i++) penalty = TK_ADD(penalty, m.pn[lane * m.ts + i]);
            if (penalty > TK_MYEPS) pen = fmin(ctx.max_penalty, TK_ADD(ctx.penalty0, penalty));
        }
        double f;
        if (pen > TK_MYEPS) {
            f = TK_DIV(pen, ctx.sqrt_nm); /* objective.f90:93-94 */
        } else {
            double s_, p_;
            Obj::fold_init(ctx, s_, p_);
            const double2* qa = reinterpret_cast<const double2*>(ta);
            const double2* qb = reinterpret_cast<const double2*>(tb);
            if (NXC > 0 && NXC <= 20) { /* loads first, then the in-order chain */
                constexpr int H = (NXC > 0 && NXC <= 20) ? (NXC + 1) / 2 : 1;
                double2 va[H + 1], vb[H + 1];
#pragma unroll
                for (int i = 0; i < H; i++) {
                    va[i] = qa[i];
                    vb[i] = nm_uses_b<Obj>() ? qb[i] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int i = 0; i < H; i++) {
                    if (2 * i < nt) Obj::fold(s_, p_, va[i].x, vb[i].x);
                    if (2 * i + 1 < nt) Obj::fold(s_, p_, va[i].y, vb[i].y);
                }
            } else {
                int i = 0;
                for (; i + 8 <= nt; i += 8) {
                    const double2 a0 = qa[(i >> 1)], a1 = qa[(i >> 1) + 1], a2 = qa[(i >> 1) + 2], a3 = qa[(i >> 1) + 3];
                    double2 b0 = make_double2(0.0, 0.0), b1 = b0, b2 = b0, b3 = b0;
                    if (nm_uses_b<Obj>()) { b0 = qb[(i >> 1)]; b1 = qb[(i >> 1) + 1]; b2 = qb[(i >> 1) + 2]; b3 = qb[(i >> 1) + 3]; }
                    Obj::fold(s_, p_, a0.x, b0.x); Obj::fold(s_, p_, a0.y, b0.y);
                    Obj::fold(s_, p_, a1.x, b1.x); Obj::fold(s_, p_, a1.y, b1.y);
                    Obj::fold(s_, p_, a2.x, b2.x); Obj::fold(s_, p_, a2.y, b2.y);
                    Obj::fold(s_, p_, a3.x, b3.x); Obj::fold(s_, p_, a3.y, b3.y);
                }
                for (; i < nt; i++) Obj::fold(s_, p_, ta[i], nm_uses_b<Obj>() ? tb[i] : 0.0);
            }
            f = Obj::finish(ctx, s_, p_);
        }
        yc = tk_objfun_finish(ctx, f, pen);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) y[c] = __shfl_sync(FULL, yc, c);
}

/* sort the np (y, vertex) pairs y[v] = m.yv[v] into the position-owned registers: rank by counting, scatter, read back */
template <int CJ, int LP = 8>
__device__ __forceinline__ void nm_w4_sort(const NmWarpMem& m, double (&ys)[CJ], int (&rows)[CJ], int np, int l8) {
    double my[CJ];
    int rk[CJ];
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int e = l8 + LP * k;
        my[k] = (e < np) ? m.yv[e] : 0.0;
        rk[k] = 0;
    }
    for (int v = 0; v < np; v++) {
        const double yk = m.yv[v];
#pragma unroll
        for (int k = 0; k < CJ; k++) rk[k] += (v != l8 + LP * k && nm_before(yk, v, my[k], l8 + LP * k)) ? 1 : 0;
    }
    __syncwarp();
    int* scr_r = reinterpret_cast<int*>(m.pn);
#pragma unroll
    for (int k = 0; k < CJ; k++)
        if (l8 + LP * k < np) { m.ta[rk[k]] = my[k]; scr_r[rk[k]] = l8 + LP * k; } /* the four sub-groups store the same values */
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int e = l8 + LP * k;
        ys[k] = (e < np) ? m.ta[e] : 0.0;
        rows[k] = (e < np) ? scr_r[e] : 0;
    }
    __syncwarp();
}

/* the entry at sorted position `pos` (lane pos % 8, slot pos / 8 of every sub-group) */
template <int CJ, int LP = 8>
__device__ __forceinline__ double nm_w4_y(const double (&ys)[CJ], int pos) {
    double v = ys[0];
#pragma unroll
    for (int k = 1; k < CJ; k++) v = ((pos / LP) == k) ? ys[k] : v;
    return __shfl_sync(FULL, v, pos % LP, LP);
}
template <int CJ, int LP = 8>
__device__ __forceinline__ int nm_w4_r(const int (&rows)[CJ], int pos) {
    int v = rows[0];
#pragma unroll
    for (int k = 1; k < CJ; k++) v = ((pos / LP) == k) ? rows[k] : v;
    return __shfl_sync(FULL, v, pos % LP, LP);
}
/* MINLOC: the lowest vertex index among the entries equal to ylo */
template <int CJ, int LP = 8>
__device__ __forceinline__ int nm_w4_ilo(const double (&ys)[CJ], const int (&rows)[CJ], double ylo, int np, int l8) {
    unsigned c = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < CJ; k++)
        if (l8 + LP * k < np && ys[k] == ylo) c = min(c, (unsigned)rows[k]);
    return (int)__reduce_min_sync(FULL, c);
}
/* vertex ihi takes the value ynew: drop sorted position nx, insert (ynew, ihi) at its place (every sub-group of LP lanes alike) */
template <int CJ, int LP>
__device__ __forceinline__ void nm_w4_insert(double (&ys)[CJ], int (&rows)[CJ], double ynew, int ihi, int nx, int l8) {
    constexpr unsigned LM = (LP == 32) ? 0xffffffffu : ((1u << (LP & 31)) - 1u);
    int pos = 0;
    double yu[CJ];
    int ru[CJ];
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int e = l8 + LP * k;
        pos += __popc(__ballot_sync(FULL, e < nx && nm_before(ys[k], rows[k], ynew, ihi)) & LM);
        yu[k] = __shfl_up_sync(FULL, ys[k], 1, LP);
        ru[k] = __shfl_up_sync(FULL, rows[k], 1, LP);
        if (k > 0) {
            const double yw = __shfl_sync(FULL, ys[k - 1], LP - 1, LP);
            const int rw = __shfl_sync(FULL, rows[k - 1], LP - 1, LP);
            if (l8 == 0) { yu[k] = yw; ru[k] = rw; }
        }
    }
#pragma unroll
    for (int k = 0; k < CJ; k++) {
        const int e = l8 + LP * k;
        if (e > pos && e <= nx) { ys[k] = yu[k]; rows[k] = ru[k]; }
        if (e == pos) { ys[k] = ynew; rows[k] = ihi; }
    }
}

/* One round's outcomes -> the device-resident searchResults table, by one warp (the in-kernel form of ingest_kernel,
 * kernels_nm.cu): rows through the f40.20 record, evaluation counts, the running best row (getBestLine :781-802:
 * lowest fval, first row among equals).  Returns true when the launch must stop here. */
__device__ bool nm_commit_round(const NmArgs& a, const NmQueue& q, int r, int lane) {
    const int nx = q.nx;
    const int s0 = r * q.round_size, s1 = min(a.n_k, s0 + q.round_size);
    double v = DBL_MAX;
    int bk = 0x7fffffff, cnt = 0;
    for (int s = s0 + lane; s < s1; s += 32) {
        const double f = __ldcg(a.out_f + s);
        if (!(f == f)) continue;
        const int k = a.first_k + s;
        const double frt = q.roundtrip ? tk_f4020_roundtrip(f) : f;
        double* dst = q.res + (size_t)k * (nx + 1);
        dst[0] = frt;
        for (int d = 0; d < nx; d++) {
            const double xv = __ldcg(a.out_x + (size_t)d * a.n_k + s);
            dst[1 + d] = q.roundtrip ? tk_f4020_roundtrip(xv) : xv;
        }
        q.valid[k] = 1;
        q.arch_evals[k] = __ldcg(a.out_evals + s);
        cnt++;
        if (bk == 0x7fffffff || frt < v) { v = frt; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(FULL, v, o);
        const int k2 = __shfl_xor_sync(FULL, bk, o);
        cnt += __shfl_xor_sync(FULL, cnt, o);
        if (k2 != 0x7fffffff && (bk == 0x7fffffff || v2 < v || (v2 == v && k2 < bk))) { v = v2; bk = k2; }
    }
    const double b0 = __ldcg(q.best + 0), b1 = __ldcg(q.best + 1), b2 = __ldcg(q.best + 2);
    const bool none = b1 < 0.0;
    const double rows_before = none ? 0.0 : b2;
    const bool take = (bk != 0x7fffffff) && (none || v < b0);
    const bool stop = take || (rows_before <= 1.0 && cnt > 0);
    __syncwarp();
    if (lane == 0) {
        q.best[2] = rows_before + (double)cnt;
        if (take) { q.best[0] = v; q.best[1] = (double)bk; }
    }
    if (take)
        for (int d = lane; d < nx; d += 32) q.best[3 + d] = q.res[(size_t)bk * (nx + 1) + 1 + d];
    __threadfence();
    return stop;
}

/* called by the warp whose restart was the last one of its round to finish: commit every complete round, in order */
__device__ void nm_try_commit(const NmArgs& a, const NmQueue& q, int lane) {
    volatile int* ctl = q.ctl;
    for (;;) {
        int go = 0;
        if (lane == 0) {
            const int c = ctl[NMQ_COMMIT];
            if (c < q.n_rounds && ctl[NMQ_GATE] >= 0) {
                const int need = min(a.n_k, (c + 1) * q.round_size) - c * q.round_size;
                if (ctl[NMQ_DONE0 + c] >= need && atomicCAS(q.ctl + NMQ_LOCK, 0, 1) == 0) go = 1;
            }
        }
        go = __shfl_sync(FULL, go, 0);
        if (!go) return;
        __threadfence();
        for (;;) { /* lock held */
            int c = 0, ready = 0;
            if (lane == 0) {
                c = ctl[NMQ_COMMIT];
                if (c < q.n_rounds && ctl[NMQ_GATE] >= 0) {
                    const int need = min(a.n_k, (c + 1) * q.round_size) - c * q.round_size;
                    ready = ctl[NMQ_DONE0 + c] >= need;
                }
            }
            c = __shfl_sync(FULL, c, 0);
            ready = __shfl_sync(FULL, ready, 0);
            if (!ready) break;
            __threadfence();
            const bool stop = nm_commit_round(a, q, c, lane);
            if (lane == 0) {
                const int committed = min(a.n_k, (c + 1) * q.round_size);
                ctl[NMQ_ACCEPTED] = committed;
                const int streak = stop ? 0 : min(ctl[NMQ_STREAK] + 1, 20);
                ctl[NMQ_STREAK] = streak;
                __threadfence();
                /* tickets below depth0 + gate may start: the committed restarts plus a window that doubles with every
                 * round committed without a change of the best row; a negative gate stops the launch */
                const long long lim = (long long)committed + ((long long)q.depth0 << streak) - q.depth0;
                ctl[NMQ_GATE] = stop ? -(1 << 30) : (int)min(lim, (long long)(1 << 29));
                ctl[NMQ_COMMIT] = c + 1;
                __threadfence();
            }
            __syncwarp();
        }
        if (lane == 0) {
            __threadfence();
            atomicExch(q.ctl + NMQ_LOCK, 0);
        }
        __syncwarp();
        /* look again: a round may have completed between the last check and the unlock */
    }
}

/* Where a warp's simplex lives.  In shared memory a 50-D simplex is 20 KB: 10 warps fit an SM, and a single in-order warp per
 * scheduler exposes every latency (ncu: 14 % warps active, issue slots 27 % busy).  One coordinate pair per lane and one row per
 * evaluation are all a warp touches between shrinks, coalesced -- so for nx >= 32 (two or more coordinates per lane) the simplex
 * goes to a per-warp slab of global memory (L1/L2 resident) and the registers, not the shared memory, bound the occupancy
 * (94 registers: 20 warps per SM).  TK_NMW4_PGLOBAL=0 keeps it in shared memory (experiments). */
#ifndef TK_NMW4_PGLOBAL
#define TK_NMW4_PGLOBAL 1
#endif
template <int CJ, int LP>
__host__ __device__ constexpr bool w4_pglobal() { return TK_NMW4_PGLOBAL && LP == 32 && CJ >= 2; }

#ifndef TK_NMW4_MINB
#define TK_NMW4_MINB 1
#endif

/* LP = lanes per evaluated point: 8 = the four-candidate form described at the top of the file; 32 = one point at a time
 * (reflection, then expansion OR contraction as minimize.f90:97-105 orders them), for objectives and dimensions where
 * most iterations are a plain accepted reflection and the speculative points would be wasted work (C4: 1.18 evaluations
 * counted per iteration).  CJ = coordinates per lane = ceil((nx+1) / LP). */
/* resident blocks the kernel is compiled for: with the simplex in global memory (two coordinates per lane: nx = 32..63, C4) the
 * registers bound the occupancy; 6 blocks of four warps = 80 registers, no spills, 24 warps per SM (measured at C4: 73.7 ms against
 * 85 ms at 124 registers / 16 warps; 64 registers / 32 warps spill: 83 ms).  The same bound for one point at a time with one coordinate
 * per lane (nx <= 31: the instance of the asynchronous run at C3) -- no spills at 80 registers, 24 instead of 20 warps per SM */
template <int CJ, int LP, int NC>
__host__ __device__ constexpr int w4_minb() { return ((w4_pglobal<CJ, LP>() && CJ == 2 && NC == 1) || (LP == 32 && CJ == 1 && NC == 1)) ? 6 : TK_NMW4_MINB; }

template <class Obj, int CJ, int NXC, int LP, int NC>
__global__ void __launch_bounds__(128, w4_minb<CJ, LP, NC>()) nm_w4_kernel(const TkTablesG T, const TkScalars sc, const NmArgs a, const NmQueue q) {
    extern __shared__ double smem[];
    constexpr int NPT = (32 / LP) * NC; /* points per evaluation pass: four sub-groups (LP = 8), or NC points as parallel chains of one lane */
    static_assert(NC == 1 || (NC == 4 && LP == 32), "NC = 4 goes with LP = 32");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sg = lane / LP, l8 = lane % LP;
    const int nx = NXC ? NXC : sc.nx, np = nx + 1;
    const int nt = Obj::nterms(nx);
    const int ts = (nx + 2) & ~1; /* even stride >= np */
    double* sbl = smem;
    double* sbu = sbl + ts;
    double* saux = sbu + ts;
    NmWarpMem m;
    {
        constexpr int NA = 2 + (nm_uses_b<Obj>() ? 1 : 0) + (Obj::kTermLocal ? 0 : 1); /* arrays of NPT*ts: ta, pn, [tb], [cand] */
        constexpr bool PG = w4_pglobal<CJ, LP>();
        double* mine = saux + ts + (size_t)w * ((NPT * NA + 1) * ts + (PG ? (size_t)0 : (size_t)np * nx));
        m.ta = mine; m.pn = mine + NPT * ts;
        double* nxt = mine + 2 * NPT * ts;
        m.tb = m.ta; m.cand = m.ta;
        if (nm_uses_b<Obj>()) { m.tb = nxt; nxt += NPT * ts; }
        if (!Obj::kTermLocal) { m.cand = nxt; nxt += NPT * ts; }
        m.yv = nxt;
        if constexpr (PG) m.p = q.pslab + ((size_t)blockIdx.x * (blockDim.x >> 5) + w) * ((size_t)np * nx);
        else m.p = nxt + ts;
        m.ts = ts;
    }
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    __syncthreads();
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, sbl, sbu, saux);
    volatile int* ctl = q.ctl;
    int jj[CJ]; /* my coordinates, clamped for addressing */
#pragma unroll
    for (int k = 0; k < CJ; k++) jj[k] = (l8 + LP * k < nx) ? l8 + LP * k : 0;

    /* ---- asynchronous instances (nm_common.cuh; the scheme of nm_quad_async_kernel with a warp as the instance and the objective
     * evaluation as the tick of the virtual clock: a restart costs cb + #evaluations) ---- */
    const tknm::NmAsync* aq = q.aq;
    const int AP = q.aq_P;
    const int inst = (int)blockIdx.x * (int)(blockDim.x >> 5) + w;
    const bool fr = aq && aq->free_run; /* free-running instances: no virtual clock, nobody waits (nm_common.cuh) */
    if (fr && blockIdx.x == gridDim.x - 1) return; /* (no time to keep) */
    if (aq && blockIdx.x == gridDim.x - 1) { /* the time keeper: every warp of the block folds its share of the clocks (thousands of them:
                                                one warp's sweep made the global time ~50 us stale) */
        __shared__ int km[4];
        const int nw = (int)(blockDim.x >> 5);
        int* gvt = aq->clk + AP + 32;
        for (;;) {
            int mn = 0x7fffffff;
            for (int i = (int)threadIdx.x; i < AP; i += (int)blockDim.x) mn = min(mn, tknm::ld_vol(aq->clk + i));
            mn = (int)__reduce_min_sync(FULL, (unsigned)mn);
            if (lane == 0) km[w] = mn;
            __syncthreads();
            int g = km[0];
            for (int j = 1; j < nw; j++) g = min(g, km[j]);
            __syncthreads();
            if (w == 0 && lane < TK_GVT_COPIES) tknm::st_vol(gvt + lane * 32, g);
            if (g == 0x7fffffff) return;
            __nanosleep(100);
        }
    }
    if (aq && inst >= AP) return;
    int ak = aq ? aq->first_k + inst : 0, aTv = 0, a_bk = -1, a_vis = 0;

    for (;;) { /* ------------------------------ one restart per trip ------------------------------ */
        int ticket = 0;
        if (aq) { /* which restart, and the best row visible at its (virtual) start */
            int cnt_ev = 0, cnt_vis = 0, bt = 0, bs = 0, bk = -1;
            double bf = 0.0;
            const int hi = min(aq->last_k, __ldcg(aq->clk + AP));
            for (int jr = lane; jr <= hi; jr += 32) {
                const int4 raw = __ldcg(reinterpret_cast<const int4*>(aq->ev + jr));
                const int tj = raw.z, sj = raw.w;
                if (tj < 0 || (!fr && tj > aTv)) continue;
                if (jr >= aq->first_k && (tj < aTv || sj < inst)) cnt_ev++;
                const double fj = __hiloint2double(raw.y, raw.x);
                if (!(fj == fj)) continue;
                cnt_vis++;
                if (bk < 0 || tknm::nm_row_before(fj, tj, sj, jr, bf, bt, bs, bk)) { bf = fj; bt = tj; bs = sj; bk = jr; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double f2 = __shfl_xor_sync(FULL, bf, o);
                const int t2 = __shfl_xor_sync(FULL, bt, o), s2 = __shfl_xor_sync(FULL, bs, o), k2 = __shfl_xor_sync(FULL, bk, o);
                if (k2 >= 0 && (bk < 0 || tknm::nm_row_before(f2, t2, s2, k2, bf, bt, bs, bk))) { bf = f2; bt = t2; bs = s2; bk = k2; }
                cnt_ev += __shfl_xor_sync(FULL, cnt_ev, o);
                cnt_vis += __shfl_xor_sync(FULL, cnt_vis, o);
            }
            if (aTv > 0) {
                if (fr) { /* the next number from the counter (clk[AP]: highest number handed out) */
                    int nk = 0;
                    if (lane == 0) nk = atomicAdd(aq->clk + AP, 1) + 1;
                    ak = __shfl_sync(FULL, nk, 0);
                } else ak = aq->first_k + AP + cnt_ev;
            }
            if (ak > aq->last_k) { if (lane == 0) tknm::st_vol(aq->clk + inst, 0x7fffffff); break; }
            a_bk = bk; a_vis = cnt_vis;
            if (lane == 0) { if (aTv > 0 && !fr) atomicMax(aq->clk + AP, ak); aq->base[ak] = (cnt_vis > 1 && bk >= 0) ? bk : -1; }
            ticket = ak - aq->first_k;
        } else {
        ticket = 0;
        if (lane == 0) ticket = (ctl[NMQ_GATE] < 0) ? 0x7fffffff : atomicAdd(q.ctl + NMQ_HEAD, 1);
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= a.n_run) break;
        }
        if (q.commit && !aq) { /* early starts are bounded: wait until the ticket is inside the window */
            bool stopped = false;
            for (;;) {
                int gate = 0;
                if (lane == 0) gate = ctl[NMQ_GATE];
                gate = __shfl_sync(FULL, gate, 0);
                if (gate < 0) { stopped = true; break; }
                if (ticket < q.depth0 + gate) break;
                __nanosleep(2000);
            }
            if (stopped) break;
        }
        const int slot = a.rank + ticket * a.world;

        /* runAmoeba :622-627 -- initial simplex around the blended start */
        {
            const bool shr = a.shr_have && a.shr_have[0] && ((float)(a.first_k + slot) / (float)a.K > 0.5f);
            const double* wsrc = shr ? a.wshr : a.width;
            for (int e = lane; e < np * nx; e += 32) {
                const int ia = e / nx, i = e - ia * nx;
                double st_i;
                if (aq) { /* getModifiedParam (:683-720), unfused; the searchStart.dat row */
                    const double ev = aq->xs[(size_t)(i + 1) * aq->K + (ak - 1)];
                    st_i = ev;
                    if (a_vis > 1 && a_bk >= 0) {
                        const double wgt = aq->w11[ak];
                        st_i = TK_ADD(TK_MUL(wgt, aq->res[(size_t)a_bk * (nx + 1) + 1 + i]), TK_MUL(TK_SUB(1.0, wgt), ev));
                    }
                    if (ia == 0) aq->arch_starts[(size_t)i * aq->K + (ak - 1)] = st_i;
                } else st_i = a.starts[(size_t)i * a.n_k + slot];
                double tt = TK_ADD(st_i, TK_DIV(TK_MUL(a.simplex[(size_t)ia * nx + i], wsrc[i]), a.scale));
                tt = fmax(fmin(tt, sbu[i]), sbl[i]);
                m.p[e] = tt;
            }
        }
        __syncwarp();
        /* objFun at the vertices, NPT at a time */
        if constexpr (NC == 4) {
            for (int v0 = 0; v0 < np; v0 += 4) {
                double xv[4][CJ], yv4[4];
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int k = 0; k < CJ; k++) xv[c][k] = m.p[(v0 + c < np ? v0 + c : 0) * nx + jj[k]];
                nm_w1_eval4<Obj, CJ, NXC>(ctx, m, xv, sbl, sbu, lane, nx, nt, yv4);
                if (lane < 4 && v0 + lane < np) m.yv[v0 + lane] = (lane == 0) ? yv4[0] : (lane == 1) ? yv4[1] : (lane == 2) ? yv4[2] : yv4[3];
            }
        } else {
            for (int v0 = 0; v0 < np; v0 += NPT) {
                const int v = v0 + sg;
                double xv[CJ];
#pragma unroll
                for (int k = 0; k < CJ; k++) xv[k] = m.p[(v < np ? v : 0) * nx + jj[k]];
                const double yc = nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, xv, sbl, sbu, sg, l8, nx, nt);
                if (l8 == 0 && v < np) m.yv[v] = yc;
            }
        }
        __syncwarp();
        double ys[CJ], psum[CJ];
        int rows[CJ];
        nm_w4_sort<CJ, LP>(m, ys, rows, np, l8);
#pragma unroll
        for (int k = 0; k < CJ; k++) { /* psum, amoeba :49 in index order */
            double s_ = 0.0;
            if (l8 + LP * k < nx)
                for (int i = 0; i < np; i++) s_ = TK_ADD(s_, m.p[i * nx + jj[k]]);
            psum[k] = s_;
        }

        int iter = 0, chk = 0;
        double ylo;
        bool discard = false;
        for (;;) { /* ---------- amoeba's main loop, minimize.f90:77-133 ---------- */
            ylo = __shfl_sync(FULL, ys[0], 0, LP);
            const double yhi = nm_w4_y<CJ, LP>(ys, nx);
            const double ynhi = nm_w4_y<CJ, LP>(ys, nx - 1);
            const int ihi = nm_w4_r<CJ, LP>(rows, nx);
            if (nm_converged(a, ylo, yhi) || iter >= a.itmax) break;
            if (aq && !fr && lane == 0) tknm::st_vol(aq->clk + inst, aTv + aq->cb + np + iter + 2); /* not converged: this pass evaluates at least one more point; then the final one */
            if (q.commit && ((++chk) & 31) == 0) { /* a committed round changed the best row: this restart is void */
                if (ctl[NMQ_GATE] < 0) { discard = true; break; }
            }
            if constexpr (NC == 4) {
                /* the four possible trial points (amotry :124-126, unfused), every lane for its own coordinates */
                double ph[CJ], X[4][CJ], ps1[CJ], y4[4];
#pragma unroll
                for (int k = 0; k < CJ; k++) {
                    ph[k] = m.p[ihi * nx + jj[k]];
                    X[0][k] = TK_SUB(TK_MUL(psum[k], a.fac1r), TK_MUL(ph[k], a.fac2r));     /* R */
                    ps1[k] = TK_ADD(TK_SUB(psum[k], ph[k]), X[0][k]);                       /* psum after R is accepted (:130) */
                    X[1][k] = TK_SUB(TK_MUL(ps1[k], a.fac1e), TK_MUL(X[0][k], a.fac2e));    /* E from R */
                    X[2][k] = TK_SUB(TK_MUL(ps1[k], a.fac1c), TK_MUL(X[0][k], a.fac2c));    /* contraction after an accepted R */
                    X[3][k] = TK_SUB(TK_MUL(psum[k], a.fac1c), TK_MUL(ph[k], a.fac2c));     /* contraction after a rejected R */
                }
                nm_w1_eval4<Obj, CJ, NXC>(ctx, m, X, sbl, sbu, lane, nx, nt, y4);
                const double yR = y4[0];
                const bool accR = yR < yhi; /* :128 */
                double pnew[CJ], ynew = yhi;
                bool moved = accR;
#pragma unroll
                for (int k = 0; k < CJ; k++) { pnew[k] = accR ? X[0][k] : ph[k]; psum[k] = accR ? ps1[k] : psum[k]; }
                if (accR) ynew = yR;
                iter++;
                if (yR <= ylo) { /* :99 expansion */
                    double yE = y4[1];
                    double Ev[CJ];
#pragma unroll
                    for (int k = 0; k < CJ; k++) Ev[k] = X[1][k];
                    if (!accR) { /* only when every y is equal; not speculated -- evaluated from the live state */
                        double X2[4][CJ], y2[4];
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            Ev[k] = TK_SUB(TK_MUL(psum[k], a.fac1e), TK_MUL(ph[k], a.fac2e));
                            X2[0][k] = X2[1][k] = X2[2][k] = X2[3][k] = Ev[k];
                        }
                        nm_w1_eval4<Obj, CJ, NXC>(ctx, m, X2, sbl, sbu, lane, nx, nt, y2);
                        yE = y2[0];
                    }
                    iter++;
                    if (yE < ynew) {
#pragma unroll
                        for (int k = 0; k < CJ; k++) { psum[k] = TK_ADD(TK_SUB(psum[k], pnew[k]), Ev[k]); pnew[k] = Ev[k]; }
                        ynew = yE; moved = true;
                    }
                } else if (yR >= ynhi) { /* :102 contraction */
                    const double yC = accR ? y4[2] : y4[3];
                    iter++;
                    if (yC < ynew) {
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            const double Cv = accR ? X[2][k] : X[3][k];
                            psum[k] = TK_ADD(TK_SUB(psum[k], pnew[k]), Cv); pnew[k] = Cv;
                        }
                        ynew = yC; moved = true;
                    } else { /* :106-113 shrink toward the best vertex, re-evaluate, re-sum */
                        if (moved) {
#pragma unroll
                            for (int k = 0; k < CJ; k++)
                                if (lane + 32 * k < nx) m.p[ihi * nx + lane + 32 * k] = pnew[k];
                        }
                        moved = false;
                        unsigned cand_lo = 0xffffffffu; /* MINLOC: the lowest vertex index among the minima (ynew > ylo here) */
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            const int e = lane + 32 * k;
                            if (e < nx && ys[k] == ylo) cand_lo = min(cand_lo, (unsigned)rows[k]);
                            if (e < np) m.yv[rows[k]] = (e == nx) ? ynew : ys[k];
                        }
                        const int ilo = (int)__reduce_min_sync(FULL, cand_lo);
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            if (lane + 32 * k >= nx) continue;
                            const double pl = m.p[ilo * nx + jj[k]];
                            for (int i = 0; i < np; i++)
                                if (i != ilo) m.p[i * nx + jj[k]] = TK_MUL(0.5, TK_ADD(m.p[i * nx + jj[k]], pl));
                        }
                        __syncwarp();
                        for (int v0 = 0; v0 < np; v0 += 4) {
                            double xv[4][CJ], yv4[4];
#pragma unroll
                            for (int c = 0; c < 4; c++)
#pragma unroll
                                for (int k = 0; k < CJ; k++) xv[c][k] = m.p[(v0 + c < np ? v0 + c : 0) * nx + jj[k]];
                            nm_w1_eval4<Obj, CJ, NXC>(ctx, m, xv, sbl, sbu, lane, nx, nt, yv4);
                            if (lane < 4 && v0 + lane < np && v0 + lane != ilo) m.yv[v0 + lane] = (lane == 0) ? yv4[0] : (lane == 1) ? yv4[1] : (lane == 2) ? yv4[2] : yv4[3];
                        }
                        __syncwarp();
                        iter += nx;
                        nm_w4_sort<CJ, LP>(m, ys, rows, np, l8);
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            double s_ = 0.0;
                            if (lane + 32 * k < nx)
                                for (int i = 0; i < np; i++) s_ = TK_ADD(s_, m.p[i * nx + jj[k]]);
                            psum[k] = s_;
                        }
                    }
                }
                if (moved) { /* vertex ihi takes (pnew, ynew): drop position nx, insert at its sorted position */
#pragma unroll
                    for (int k = 0; k < CJ; k++)
                        if (lane + 32 * k < nx) m.p[ihi * nx + lane + 32 * k] = pnew[k];
                    nm_w4_insert<CJ, LP>(ys, rows, ynew, ihi, nx, l8);
                }
            } else if constexpr (LP == 32) {
                /* amotry(-1) :97, then amotry(2) :100 or amotry(0.5) :104, each from the live psum / p(ihi,:) */
                double ph[CJ], xt[CJ];
#pragma unroll
                for (int k = 0; k < CJ; k++) {
                    ph[k] = m.p[ihi * nx + jj[k]];
                    xt[k] = TK_SUB(TK_MUL(psum[k], a.fac1r), TK_MUL(ph[k], a.fac2r));
                }
                const double yR = nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, xt, sbl, sbu, 0, l8, nx, nt);
                double ynew = yhi;
                bool moved = false;
                if (yR < yhi) { /* :128-131 */
#pragma unroll
                    for (int k = 0; k < CJ; k++) { psum[k] = TK_ADD(TK_SUB(psum[k], ph[k]), xt[k]); ph[k] = xt[k]; }
                    ynew = yR; moved = true;
                }
                iter++;
                const bool expand = yR <= ylo, contract = !expand && yR >= ynhi;
                if (expand || contract) {
                    const double f1 = expand ? a.fac1e : a.fac1c, f2 = expand ? a.fac2e : a.fac2c;
#pragma unroll
                    for (int k = 0; k < CJ; k++) xt[k] = TK_SUB(TK_MUL(psum[k], f1), TK_MUL(ph[k], f2));
                    const double y2 = nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, xt, sbl, sbu, 0, l8, nx, nt);
                    iter++;
                    if (y2 < ynew) {
#pragma unroll
                        for (int k = 0; k < CJ; k++) { psum[k] = TK_ADD(TK_SUB(psum[k], ph[k]), xt[k]); ph[k] = xt[k]; }
                        ynew = y2; moved = true;
                    } else if (contract) { /* :106-113 shrink toward the best vertex, re-evaluate, re-sum */
                        if (moved) {
#pragma unroll
                            for (int k = 0; k < CJ; k++)
                                if (l8 + LP * k < nx) m.p[ihi * nx + l8 + LP * k] = ph[k];
                        }
                        moved = false;
                        unsigned cand_lo = 0xffffffffu; /* MINLOC: the lowest vertex index among the minima (ynew > ylo here) */
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            const int e = l8 + LP * k;
                            if (e < nx && ys[k] == ylo) cand_lo = min(cand_lo, (unsigned)rows[k]);
                            if (e < np) m.yv[rows[k]] = (e == nx) ? ynew : ys[k];
                        }
                        const int ilo = (int)__reduce_min_sync(FULL, cand_lo);
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            if (l8 + LP * k >= nx) continue;
                            const double pl = m.p[ilo * nx + jj[k]];
                            for (int i = 0; i < np; i++)
                                if (i != ilo) m.p[i * nx + jj[k]] = TK_MUL(0.5, TK_ADD(m.p[i * nx + jj[k]], pl));
                        }
                        __syncwarp();
                        for (int v = 0; v < np; v++) {
                            if (v == ilo) continue;
                            double xv[CJ];
#pragma unroll
                            for (int k = 0; k < CJ; k++) xv[k] = m.p[v * nx + jj[k]];
                            const double y3 = nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, xv, sbl, sbu, 0, l8, nx, nt);
                            if (l8 == 0) m.yv[v] = y3;
                        }
                        __syncwarp();
                        iter += nx;
                        nm_w4_sort<CJ, LP>(m, ys, rows, np, l8);
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            double s_ = 0.0;
                            if (l8 + LP * k < nx)
                                for (int i = 0; i < np; i++) s_ = TK_ADD(s_, m.p[i * nx + jj[k]]);
                            psum[k] = s_;
                        }
                    }
                }
                if (moved) { /* vertex ihi takes (ph, ynew): drop position nx, insert at its sorted position */
#pragma unroll
                    for (int k = 0; k < CJ; k++)
                        if (l8 + LP * k < nx) m.p[ihi * nx + l8 + LP * k] = ph[k];
                    nm_w4_insert<CJ, LP>(ys, rows, ynew, ihi, nx, l8);
                }
            } else {
            /* the four possible trial points (amotry :124-126, unfused); sub-group sg evaluates number sg */
            double ph[CJ], R[CJ], ps1[CJ], xc[CJ];
#pragma unroll
            for (int k = 0; k < CJ; k++) {
                ph[k] = m.p[ihi * nx + jj[k]];
                R[k] = TK_SUB(TK_MUL(psum[k], a.fac1r), TK_MUL(ph[k], a.fac2r));
                ps1[k] = TK_ADD(TK_SUB(psum[k], ph[k]), R[k]); /* psum after R is accepted (:130) */
                const double pa = (sg == 3) ? psum[k] : ps1[k];
                const double pb = (sg == 3) ? ph[k] : R[k];
                const double f1 = (sg == 1) ? a.fac1e : a.fac1c;
                const double f2 = (sg == 1) ? a.fac2e : a.fac2c;
                const double o = TK_SUB(TK_MUL(pa, f1), TK_MUL(pb, f2)); /* E, Co or Ci */
                xc[k] = (sg == 0) ? R[k] : o;
            }
            const double yc = nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, xc, sbl, sbu, sg, l8, nx, nt);
            const double yR = __shfl_sync(FULL, yc, 0);
            const double yE0 = __shfl_sync(FULL, yc, 8);
            const double yCo = __shfl_sync(FULL, yc, 16);
            const double yCi = __shfl_sync(FULL, yc, 24);
            const bool accR = yR < yhi; /* :128 */
            double pnew[CJ], ynew = yhi;
            bool moved = accR;
#pragma unroll
            for (int k = 0; k < CJ; k++) { pnew[k] = accR ? R[k] : ph[k]; psum[k] = accR ? ps1[k] : psum[k]; }
            if (accR) ynew = yR;
            iter++;
            if (yR <= ylo) { /* :99 expansion */
                double yE = yE0;
                double Ev[CJ];
                if (!accR) { /* only when every y is equal; not speculated -- evaluated from the live state */
#pragma unroll
                    for (int k = 0; k < CJ; k++) Ev[k] = TK_SUB(TK_MUL(psum[k], a.fac1e), TK_MUL(ph[k], a.fac2e));
                    yE = __shfl_sync(FULL, nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, Ev, sbl, sbu, sg, l8, nx, nt), 0);
                } else {
#pragma unroll
                    for (int k = 0; k < CJ; k++) Ev[k] = TK_SUB(TK_MUL(ps1[k], a.fac1e), TK_MUL(R[k], a.fac2e));
                }
                iter++;
                if (yE < ynew) {
#pragma unroll
                    for (int k = 0; k < CJ; k++) { psum[k] = TK_ADD(TK_SUB(psum[k], pnew[k]), Ev[k]); pnew[k] = Ev[k]; }
                    ynew = yE; moved = true;
                }
            } else if (yR >= ynhi) { /* :102 contraction */
                const double yC = accR ? yCo : yCi;
                iter++;
                if (yC < ynew) {
#pragma unroll
                    for (int k = 0; k < CJ; k++) {
                        const double Cv = accR ? TK_SUB(TK_MUL(ps1[k], a.fac1c), TK_MUL(R[k], a.fac2c))
                                               : TK_SUB(TK_MUL(psum[k], a.fac1c), TK_MUL(ph[k], a.fac2c));
                        psum[k] = TK_ADD(TK_SUB(psum[k], pnew[k]), Cv); pnew[k] = Cv;
                    }
                    ynew = yC; moved = true;
                } else { /* :106-113 shrink toward the best vertex, re-evaluate, re-sum */
                    if (moved && sg == 0) {
#pragma unroll
                        for (int k = 0; k < CJ; k++)
                            if (l8 + 8 * k < nx) m.p[ihi * nx + l8 + 8 * k] = pnew[k];
                    }
                    moved = false;
                    /* y by vertex; MINLOC: the lowest vertex index among the minima (ynew > ylo here, so ihi is not one) */
                    unsigned cand_lo = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < CJ; k++) {
                        const int e = l8 + 8 * k;
                        if (e < nx && ys[k] == ylo) cand_lo = min(cand_lo, (unsigned)rows[k]);
                        if (e < np && sg == 0) m.yv[rows[k]] = (e == nx) ? ynew : ys[k];
                    }
                    const int ilo = (int)__reduce_min_sync(FULL, cand_lo);
                    __syncwarp();
                    if (sg == 0) {
#pragma unroll
                        for (int k = 0; k < CJ; k++) {
                            if (l8 + 8 * k >= nx) continue;
                            const double pl = m.p[ilo * nx + jj[k]];
                            for (int i = 0; i < np; i++)
                                if (i != ilo) m.p[i * nx + jj[k]] = TK_MUL(0.5, TK_ADD(m.p[i * nx + jj[k]], pl));
                        }
                    }
                    __syncwarp();
                    for (int v0 = 0; v0 < np; v0 += 4) {
                        const int v = v0 + sg;
                        double xv[CJ];
#pragma unroll
                        for (int k = 0; k < CJ; k++) xv[k] = m.p[(v < np ? v : 0) * nx + jj[k]];
                        const double y2 = nm_w4_eval<Obj, CJ, NXC, LP>(ctx, m, xv, sbl, sbu, sg, l8, nx, nt);
                        if (l8 == 0 && v < np && v != ilo) m.yv[v] = y2;
                    }
                    __syncwarp();
                    iter += nx;
                    nm_w4_sort<CJ, LP>(m, ys, rows, np, l8);
#pragma unroll
                    for (int k = 0; k < CJ; k++) {
                        double s_ = 0.0;
                        if (l8 + 8 * k < nx)
                            for (int i = 0; i < np; i++) s_ = TK_ADD(s_, m.p[i * nx + jj[k]]);
                        psum[k] = s_;
                    }
                }
            }
            if (moved) { /* vertex ihi takes (pnew, ynew): drop position nx, insert at its sorted position */
                if (sg == 0) {
#pragma unroll
                    for (int k = 0; k < CJ; k++)
                        if (l8 + 8 * k < nx) m.p[ihi * nx + l8 + 8 * k] = pnew[k];
                }
                nm_w4_insert<CJ, 8>(ys, rows, ynew, ihi, nx, l8);
            }
            }
            __syncwarp();
        }
        if (discard) break;
        /* runAmoeba :631-632 returns the first minimal vertex; completeSearch :582 re-evaluates objFun there: same point,
         * same value */
        {
            const int il = nm_w4_ilo<CJ, LP>(ys, rows, ylo, np, l8);
            if (aq) { /* completeSearch :582-597: the row of restart ak, then the finish event, then the wait for the global time */
                int F = aTv + aq->cb + np + iter + 1;
                double* dst = aq->res + (size_t)ak * (nx + 1);
                const double frt = aq->roundtrip ? tk_f4020_roundtrip(ylo) : ylo;
                for (int j = lane; j < nx; j += 32) { const double xv = m.p[il * nx + j]; dst[1 + j] = aq->roundtrip ? tk_f4020_roundtrip(xv) : xv; }
                if (lane == 0) {
                    dst[0] = frt;
                    aq->ev[ak].fv = frt;
                    aq->ev[ak].slt = inst;
                    aq->valid[ak] = (frt == frt) ? 1 : 0;
                    aq->arch_evals[ak] = np + iter + 1;
                }
                __syncwarp();
                if (fr) { /* the order of the commits is the order of the file; nobody waits for anybody */
                    if (lane == 0) {
                        F = 1 + atomicAdd(aq->clk + AP + 1, 1);
                        __threadfence();
                        tknm::st_vol(&aq->ev[ak].fin, F);
                        __threadfence();
                    }
                    __syncwarp();
                    aTv = 1;
                    continue;
                }
                if (lane == 0) {
                    __threadfence();
                    tknm::st_vol(&aq->ev[ak].fin, F);
                    __threadfence();
                    tknm::st_vol(aq->clk + inst, F + aq->cb + np + 1);
                    const int* gvt = aq->clk + AP + 32 + (inst % TK_GVT_COPIES) * 32;
                    unsigned ns = 100;
                    const long long t0w = clock64();
                    while (!(tknm::ld_vol(gvt) > F)) {
                        if (clock64() - t0w > TK_ASYNC_PATIENCE) { atomicExch(aq->err, 1); break; }
                        __nanosleep(ns);
                        if (ns < 800) ns += ns;
                    }
                    __threadfence();
                }
                __syncwarp();
                int bad = 0;
                if (lane == 0) bad = tknm::ld_vol(aq->err);
                bad = __shfl_sync(FULL, bad, 0);
                if (bad) { if (lane == 0) tknm::st_vol(aq->clk + inst, 0x7fffffff); break; }
                aTv = F;
                continue;
            }
            for (int j = lane; j < nx; j += 32) a.out_x[(size_t)j * a.n_k + slot] = m.p[il * nx + j];
            if (lane == 0) {
                a.out_f[slot] = ylo;
                a.out_evals[slot] = np + iter + 1;
            }
        }
        if (q.commit) {
            __threadfence();
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                const int r = ticket / q.round_size;
                const int need = min(a.n_k, (r + 1) * q.round_size) - r * q.round_size;
                last = (atomicAdd(q.ctl + NMQ_DONE0 + r, 1) + 1 == need);
            }
            last = __shfl_sync(FULL, last, 0);
            if (last) nm_try_commit(a, q, lane);
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------------------------- */

/* which form: four candidates per pass (LP = 8) unless most of them would be wasted */
/* mode = LP * 10 + NC: 81 = four sub-groups of eight lanes, 321 = one point at a time, 324 = four points as parallel chains */
static int g_w4_mode_force = 0; /* the asynchronous launcher's choice while it configures and launches */
inline int w4_mode(int nx) {
    if (g_w4_mode_force) return (g_w4_mode_force == 81 && nx > 63) ? 321 : g_w4_mode_force;
    const char* e = getenv("TIKTAK_NM_MODE");
    int mode = (nx > 31) ? 321 : 324;
    if (e && (atoi(e) == 81 || atoi(e) == 321 || atoi(e) == 324)) mode = atoi(e);
    if (mode == 81 && nx > 63) mode = 321;
    return mode;
}
template <class Obj>
inline int w4_lp(int nx) { return w4_mode(nx) / 10; }
template <class Obj>
inline size_t w4_warp_doubles(int nx) {
    const int ts = (nx + 2) & ~1;
    const int na = 2 + (nm_uses_b<Obj>() ? 1 : 0) + (Obj::kTermLocal ? 0 : 1);
    const int npt = (32 / w4_lp<Obj>(nx)) * (w4_mode(nx) % 10);
    const bool pg = TK_NMW4_PGLOBAL && w4_lp<Obj>(nx) == 32 && (nx + 1 + 31) / 32 >= 2; /* = w4_pglobal<CJ, LP>() */
    return (size_t)(npt * na + 1) * ts + (pg ? 0 : (size_t)(nx + 1) * nx);
}
/* warps per block: as many as keep a block's shared memory small enough for many blocks per SM, at least one */
template <class Obj>
inline int w4_wpb(int nx) {
    const size_t per_warp = w4_warp_doubles<Obj>(nx) * sizeof(double);
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 24 * 1024) wpb >>= 1;
    return wpb;
}
template <class Obj>
inline size_t w4_shmem(int nx, int wpb) {
    const int ts = (nx + 2) & ~1;
    return ((size_t)3 * ts + w4_warp_doubles<Obj>(nx) * wpb) * sizeof(double);
}

template <class Obj, int CJ, int NXC, int LP, int NC>
int w4_configure(int nx, int wpb, int* blocks_per_sm) {
    const size_t shmem = w4_shmem<Obj>(nx, wpb);
    TK_CUDA(cudaFuncSetAttribute(nm_w4_kernel<Obj, CJ, NXC, LP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(shmem, 48 * 1024)));
    TK_CUDA(cudaFuncSetAttribute(nm_w4_kernel<Obj, CJ, NXC, LP, NC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    TK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, nm_w4_kernel<Obj, CJ, NXC, LP, NC>, wpb * 32, shmem));
    return TIKTAK_OK;
}

/* f() with the instantiation for (Obj, nx): compiled-in nx for the BASELINE pairs, run-time nx otherwise */
template <class Obj, class F>
int w4_dispatch(int nx, F&& f) {
    const int mode = w4_mode(nx), lp = mode / 10;
    const int cj = (nx + 1 + lp - 1) / lp;
    if (mode == 81) {
        if constexpr (Obj::kId == TIKTAK_OBJ_GRIEWANK) { if (nx == 10) return f.template operator()<2, 10, 8, 1>(); }
        if (cj <= 1) return f.template operator()<1, 0, 8, 1>();
        if (cj <= 2) return f.template operator()<2, 0, 8, 1>();
        if (cj <= 4) return f.template operator()<4, 0, 8, 1>();
        return f.template operator()<8, 0, 8, 1>();
    }
    if (mode == 324) {
        if constexpr (Obj::kId == TIKTAK_OBJ_TEMPLATE) { if (nx == 4) return f.template operator()<1, 4, 32, 4>(); }
        if constexpr (Obj::kId == TIKTAK_OBJ_GRIEWANK) { if (nx == 10) return f.template operator()<1, 10, 32, 4>(); }
        if constexpr (Obj::kId == TIKTAK_OBJ_RASTRIGIN) { if (nx == 20) return f.template operator()<1, 20, 32, 4>(); }
        if constexpr (Obj::kId == TIKTAK_OBJ_ROSENBROCK) { if (nx == 50) return f.template operator()<2, 50, 32, 4>(); }
        if (cj <= 1) return f.template operator()<1, 0, 32, 4>();
        if (cj <= 2) return f.template operator()<2, 0, 32, 4>();
        return f.template operator()<4, 0, 32, 4>();
    }
    if constexpr (Obj::kId == TIKTAK_OBJ_GRIEWANK) { if (nx == 10) return f.template operator()<1, 10, 32, 1>(); }
    if constexpr (Obj::kId == TIKTAK_OBJ_RASTRIGIN) { if (nx == 20) return f.template operator()<1, 20, 32, 1>(); }
    if constexpr (Obj::kId == TIKTAK_OBJ_ROSENBROCK) { if (nx == 50) return f.template operator()<2, 50, 32, 1>(); }
    if (cj <= 1) return f.template operator()<1, 0, 32, 1>();
    if (cj <= 2) return f.template operator()<2, 0, 32, 1>();
    return f.template operator()<4, 0, 32, 1>();
}

} /* namespace */

bool tk_nm_w4_supported(tiktak_handle* h) {
    const int nx = h->cfg.nx;
    if (h->cfg.objective_id == TIKTAK_OBJ_SMM || nx > 127) return false;
    bool ok = false;
    tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int { ok = Obj::nterms(nx) <= nx && w4_shmem<Obj>(nx, 1) <= 227 * 1024; return TIKTAK_OK; });
    return ok;
}

/* restarts resident at once (all SMs) */
int tk_nm_w4_capacity(tiktak_handle* h) {
    const int nx = h->cfg.nx;
    int bps = 0, wpb = 1;
    int rc = tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        wpb = w4_wpb<Obj>(nx);
        return w4_dispatch<Obj>(nx, [&]<int CJ, int NXC, int LP, int NC>() -> int { return w4_configure<Obj, CJ, NXC, LP, NC>(nx, wpb, &bps); });
    });
    if (rc) return 0;
    return bps * wpb * h->sm_count;
}

int tk_launch_nm_w4(tiktak_handle* h, const NmArgs& a, const NmQueue& q) {
    const int nx = h->cfg.nx;
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    return tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        const int wpb = w4_wpb<Obj>(nx);
        return w4_dispatch<Obj>(nx, [&]<int CJ, int NXC, int LP, int NC>() -> int {
            int bps = 0;
            int rc = w4_configure<Obj, CJ, NXC, LP, NC>(nx, wpb, &bps);
            if (rc) return rc;
            if (bps < 1) { tk_set_error("nx=%d: the Nelder-Mead simplex does not fit in shared memory", nx); return TIKTAK_ERR_UNSUPPORTED; }
            /* warps that could only wait for the window to open are not launched */
            const int useful = q.commit ? std::min(a.n_run, std::max(4 * q.depth0, q.round_size)) : a.n_run;
            const int need = (useful + wpb - 1) / wpb;
            const int grid = std::max(1, std::min(need, bps * h->sm_count));
            NmQueue qq = q;
            qq.pslab = nullptr;
            if (w4_pglobal<CJ, LP>()) { /* the warps' simplices in global memory */
                const size_t doubles = (size_t)grid * wpb * (size_t)(nx + 1) * nx;
                if (h->nm_pslab_doubles < doubles) {
                    if (h->d_nm_pslab) cudaFree(h->d_nm_pslab);
                    h->d_nm_pslab = nullptr; h->nm_pslab_doubles = 0;
                    TK_CUDA(cudaMalloc(&h->d_nm_pslab, doubles * sizeof(double)));
                    h->nm_pslab_doubles = doubles;
                }
                qq.pslab = h->d_nm_pslab;
            }
            nm_w4_kernel<Obj, CJ, NXC, LP, NC><<<grid, wpb * 32, w4_shmem<Obj>(nx, wpb), h->stream>>>(T, h->sc, a, qq);
            h->launches++;
            TK_CUDA(cudaGetLastError());
            return TIKTAK_OK;
        });
    });
}

/* Asynchronous instances with a WARP as the instance (nm_common.cuh; tk_launch_nm_async in kernels_nm.cu decides between this and
 * the four-warp blocks of nm_quad_async_kernel): the tick of the virtual clock is one objective evaluation, a restart costs
 * cb + its evaluation count.  One cooperative launch: every instance and the time keeper's block must be resident. */
int tk_launch_nm_w4_async(tiktak_handle* h, const NmArgs& a, const NmAsync& aq, int want, int* used) {
    const int nx = h->cfg.nx;
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    /* one point at a time (mode 321) spends the fewest instructions per counted evaluation: what a throughput-bound run wants */
    int mode = 321;
    if (const char* e = getenv("TIKTAK_ASYNC_MODE")) { if (atoi(e) == 81 || atoi(e) == 321 || atoi(e) == 324) mode = atoi(e); }
    g_w4_mode_force = mode;
    int rc = tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        int wpb = w4_wpb<Obj>(nx);
        if (const char* e = getenv("TIKTAK_ASYNC_WPB")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) wpb = std::min(wpb, v); }
        return w4_dispatch<Obj>(nx, [&]<int CJ, int NXC, int LP, int NC>() -> int {
            int bps = 0;
            int rc = w4_configure<Obj, CJ, NXC, LP, NC>(nx, wpb, &bps);
            if (rc) return rc;
            if (bps < 1) { tk_set_error("nx=%d: the Nelder-Mead simplex does not fit in shared memory", nx); return TIKTAK_ERR_UNSUPPORTED; }
            const int n = aq.last_k - aq.first_k + 1;
            const int P = std::max(1, std::min(std::min(want, n), (bps * h->sm_count - 1) * wpb));
            const int grid = (P + wpb - 1) / wpb + 1; /* + the time keeper */
            NmQueue qq{};
            if (!h->d_nmq) TK_CUDA(cudaMalloc(&h->d_nmq, (size_t)(NMQ_DONE0 + h->cfg.maxpoints_plus1 + 1) * sizeof(int)));
            qq.ctl = h->d_nmq; qq.round_size = n; qq.n_rounds = 1; qq.roundtrip = h->cfg.text_roundtrip; qq.nx = nx;
            qq.res = h->d_res; qq.valid = h->d_res_valid; qq.arch_evals = h->d_arch_evals; qq.best = h->d_best;
            if (w4_pglobal<CJ, LP>()) {
                const size_t doubles = (size_t)grid * wpb * (size_t)(nx + 1) * nx;
                if (h->nm_pslab_doubles < doubles) {
                    if (h->d_nm_pslab) cudaFree(h->d_nm_pslab);
                    h->d_nm_pslab = nullptr; h->nm_pslab_doubles = 0;
                    TK_CUDA(cudaMalloc(&h->d_nm_pslab, doubles * sizeof(double)));
                    h->nm_pslab_doubles = doubles;
                }
                qq.pslab = h->d_nm_pslab;
            }
            rc = tk_async_init(h, aq, P);
            if (rc) return rc;
            if (!h->d_async_desc) TK_CUDA(cudaMalloc(&h->d_async_desc, sizeof(NmAsync)));
            TK_CUDA(cudaMemcpyAsync(h->d_async_desc, &aq, sizeof(NmAsync), cudaMemcpyHostToDevice, h->stream)); /* (pageable source: staged before the call returns) */
            qq.aq = (const NmAsync*)h->d_async_desc;
            qq.aq_P = P;
            void* args[] = {(void*)&T, (void*)&h->sc, (void*)&a, (void*)&qq};
            TK_CUDA(cudaLaunchCooperativeKernel((const void*)nm_w4_kernel<Obj, CJ, NXC, LP, NC>, dim3((unsigned)grid), dim3((unsigned)wpb * 32), args,
                                                w4_shmem<Obj>(nx, wpb), h->stream));
            h->launches++;
            *used = P;
            return TIKTAK_OK;
        });
    });
    g_w4_mode_force = 0;
    return rc;
}
