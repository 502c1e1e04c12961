/* kernels_dfnls.cu -- batched DFNLS (derivative-free nonlinear least squares): one restart per block.
 *
 * Replaces bobyqa_h and its family as called from completeSearch (TiktakGlobalSearch.f90:564-572):
 *   BOBYQA_H minimize.f90:262-397, BOBYQB_H :405-1171, PRELIM_H :1473-1619, TRSBOX_H :2032-2488,
 *   ALTMOV :1186-1464, UPDATE :2492-2563, RESCUE_H :1627-2023.
 *
 * Mapping: everything indexed by the moment m1 = 1..mv (GQV, HQV, PQV, FVAL_V, WV,
 * v_err, v_diff ...; 80*mv doubles per restart, in a per-block slab of global memory laid out [column][m] so
 * the threads of the block read/write it coalesced) is the parallel axis; the small matrices of Powell's
 * method (XPT, BMAT, ZMAT, the packed HQ, W, VLAG: O(npt*(npt+n))) live in shared memory and the sequential
 * scalar logic -- the labels 20/60/90/210/230/360/650/680/720 of BOBYQB_H and all of TRSBOX/ALTMOV/UPDATE --
 * runs on ONE master thread, statement for statement as in the reference.  Whenever the master reaches a
 * loop over the moments it posts an operation code and the whole block executes it between two named
 * barriers (workers spin in worker_loop()).  The only reductions ACROSS moments (F = sum v^2 :1173-1183,
 * GOPT = 2 J^T v :2068-2076, HQ :2095-2125) are summed in index order by a single thread each, so the
 * result is bit-identical to the serial reference order.  All arithmetic unfused (-fmad=false).
 *
 * The objective is evaluated inside the kernel by the block itself (OP_DFOVEC); objectives whose moments are
 * copies of one value (the shipped template) use the scalar plugin contract.
 *
 * rescue: RESCUE_H (:1627-2023, label 190 :643-675) is entered only after a denominator was damaged by rounding (:746-747,
 * :782-783) -- 0 times in the 1001 restarts of C5 and in 3600 oracle runs over the four synthetic objectives.  It is on the
 * device all the same (master routine RESCUE_H + OP_RESC_*; out of line, so the rare path costs the common one no registers);
 * tiktak_cuda_local_nrescue returns the calls per restart.  The test hook TIKTAK_DFNLS_FORCE_RESCUE=<nf>[,<every>] makes the
 * two denominator tests count as failed at chosen evaluation numbers (the oracle has the same hook): tests/test_gpu_parity.py
 * compares whole runs with dozens of rescues, with and without evaluations inside RESCUE_H, bit for bit.
 * nx <= 100 (nmax of :416): above nx ~ 55 XPT, BMAT and ZMAT move from shared memory to the restart's slab (DfArgs.big_in_global).
 */
#include "nm_common.cuh"
#include "smm.cuh"

namespace {

enum DfOp : int {
    OP_EXIT = 0, OP_DFOVEC, OP_PRELIM_ZERO, OP_PRELIM_STORE, OP_PRELIM_G, OP_PRELIM_H, OP_PRELIM_SWAP,
    OP_L20_HQV, OP_L20_PQV, OP_SHIFT, OP_VQUAD, OP_UPDATE_MODEL, OP_GQV_W, OP_OPT_MOVE, OP_LFN, OP_TRS_MODEL, OP_FINAL,
    OP_RESC_SHIFT, OP_RESC_NEWPT, OP_RESC_MODEL, OP_ASYNC_PICK
};

struct DfArgs {
    int n_run, n_k, rank, world;
    const double* starts; /* (n_k, nx) column-major */
    const double* rhobeg; /* per restart slot */
    const double* rhoend;
    int maxfun, npt, mv;
    double* slab;         /* per-block scratch: slab_doubles each */
    size_t slab_doubles;
    double* out_x;        /* (n_k, nx) column-major */
    double* out_f;
    int* out_evals;
    int* out_status;
    int force_first, force_every; /* test hook: treat the denominator test as failed when NF reaches force_first (+ k force_every) */
    int big_in_global;    /* XPT, BMAT, ZMAT live in the slab instead of shared memory (nx > ~55) */
    int async;            /* asynchronous instances (nm_common.cuh): block r is instance r, block n_run keeps the time; rhobeg / rhoend are
                             indexed by k - aq.first_k */
    tknm::NmAsync aq;
};

struct DfCtl { /* master <-> workers */
    int op;
    int i0, i1, i2, i3;
    double d0, d1, d2, d3;
    int model_update;
};

template <class Obj>
struct Df {
    /* sizes */
    int N, NPT, NDIM, mv, NP, NPTM, NH, tid, nthr;
    DfCtl* ctl;
    const tk_obj_ctx* ctx;
    /* shared small arrays (1-based through the macros below) */
    double *XBASE, *XPT, *FVAL, *XOPT, *GOPT, *HQ, *BMAT, *ZMAT, *SL, *SU, *XNEW, *XALT, *D, *VLAG, *W, *HD1, *TK, *XC, *VBUF, *VSH, *YSH;
    const double *XL, *XU;
    /* global per-moment arrays */
    double *GQV, *HQV, *PQV, *FVAL_V, *WV, *v_sumpq, *v_err, *v_beg, *v_temp, *v_diff, *v_vquad, *PQVold, *v_fbase;
    int* v_itest;
    double eps;
    bool model_update, opt_update;
    int nrescue;
    int force_first, force_every, forced_at;
    /* asynchronous instances */
    const tknm::NmAsync* aq = nullptr;
    int a_t0 = 0, a_nev = 0; /* virtual time at which the running restart's evaluations start; evaluations so far */
    int* a_clk = nullptr;    /* this instance's clock (NULL: rounds, and the Nelder-Mead kernel for vector objectives) */
    double* pk_f; int *pk_t, *pk_s, *pk_k, *pk_v, *pk_e; /* per-warp results of OP_ASYNC_PICK */

#define XPT_(k, j) XPT[((j)-1) * NPT + ((k)-1)]
#define BMAT_(i, j) BMAT[((j)-1) * NDIM + ((i)-1)]
#define ZMAT_(k, j) ZMAT[((j)-1) * NPT + ((k)-1)]
#define GQV_(m, j) GQV[(size_t)((j)-1) * mv + ((m)-1)]
#define HQV_(m, ih) HQV[(size_t)((ih)-1) * mv + ((m)-1)]
#define PQV_(m, k) PQV[(size_t)((k)-1) * mv + ((m)-1)]
#define FVALV_(m, k) FVAL_V[(size_t)((k)-1) * mv + ((m)-1)]
#define WV_(m, j) WV[(size_t)((j)-1) * mv + ((m)-1)]
#define V1(a, i) a[(i)-1]
#define FOR_M for (int m1 = tid + 1; m1 <= mv; m1 += nthr)

    __device__ __forceinline__ void bar() { asm volatile("barrier.sync 1, %0;" ::"r"(nthr) : "memory"); }

    /* the moment-parallel parts of RESCUE_H: out of line, so that the rare path does not raise the register pressure of the
     * operations every iteration runs */
    __device__ __noinline__ void run_op_rescue(int op) {
        const DfCtl& c = *ctl;
        switch (op) {
            case OP_RESC_SHIFT: { /* RESCUE_H :1663-1690, :1713-1714: XPT is shifted already, XOPT not yet zeroed */
                const int KOPT = c.i0;
                FOR_M {
                    double sp = 0.0;
                    for (int K = 1; K <= NPT; K++) sp = sp + PQV_(m1, K);
                    V1(v_sumpq, m1) = sp;
                    for (int J = 1; J <= N; J++) {
                        double wv = 0.5 * sp * V1(XOPT, J);
                        for (int K = 1; K <= NPT; K++) wv = wv + PQV_(m1, K) * XPT_(K, J);
                        WV_(m1, J) = wv;
                    }
                    int IH = 0;
                    for (int J = 1; J <= N; J++)
                        for (int I = 1; I <= J; I++) {
                            IH++;
                            HQV_(m1, IH) = HQV_(m1, IH) + WV_(m1, I) * V1(XOPT, J) + WV_(m1, J) * V1(XOPT, I);
                        }
                    V1(v_fbase, m1) = FVALV_(m1, KOPT);
                }
                break;
            }
            case OP_RESC_NEWPT: { /* :1903-1949: the provisional point KPT enters XPT; the old row is in TK[1..N], the
                                   * scalar products XP*XPT(K,IP) + XQ*XPT(K,IQ) in TK[N+K] */
                const int KPT = c.i0, IP = c.i1, IQ = c.i2;
                const double XP = c.d0, XQ = c.d1;
                const int IHP = (IP + IP * IP) / 2, IHQ = (IQ + IQ * IQ) / 2;
                FOR_M {
                    const double pk = PQV_(m1, KPT);
                    int IH = 0;
                    for (int J = 1; J <= N; J++) {
                        const double vt = pk * V1(TK, J);
                        for (int I = 1; I <= J; I++) {
                            IH++;
                            HQV_(m1, IH) = HQV_(m1, IH) + vt * V1(TK, I);
                        }
                    }
                    PQV_(m1, KPT) = 0.0;
                    double vq = V1(v_fbase, m1);
                    if (IP > 0) vq = vq + XP * (GQV_(m1, IP) + 0.5 * XP * HQV_(m1, IHP));
                    if (IQ > 0) {
                        vq = vq + XQ * (GQV_(m1, IQ) + 0.5 * XQ * HQV_(m1, IHQ));
                        if (IP > 0) {
                            const int IW = (IHP > IHQ ? IHP : IHQ) - (IP > IQ ? IP - IQ : IQ - IP);
                            vq = vq + XP * XQ * HQV_(m1, IW);
                        }
                    }
                    for (int K = 1; K <= NPT; K++) vq = vq + 0.5 * PQV_(m1, K) * V1(TK, N + K) * V1(TK, N + K);
                    V1(v_vquad, m1) = vq;
                }
                break;
            }
            case OP_RESC_MODEL: { /* :1979-2019 after the evaluation at the new point: SUM_K = ZMAT(K,.) . ZMAT(KPT,.) in TK[K] */
                const int KPT = c.i0;
                const double* PTSAUX = W;              /* (2, N) column-major, the caller's W(1) */
                const double* PTSID = W + (N + NP) - 2; /* 1-based: PTSID[K] = W(N+NP-1+K) */
                FOR_M {
                    const double e = V1(v_err, m1);
                    FVALV_(m1, KPT) = e;
                    const double df = e - V1(v_vquad, m1);
                    V1(v_diff, m1) = df;
                    for (int I = 1; I <= N; I++) GQV_(m1, I) = GQV_(m1, I) + df * BMAT_(KPT, I);
                    for (int K = 1; K <= NPT; K++) {
                        const double vt = df * V1(TK, K);
                        const double id = PTSID[K];
                        if (fabs(id) <= eps) {
                            PQV_(m1, K) = PQV_(m1, K) * vt; /* sic (:1999): a product where Powell's RESCUE adds */
                        } else {
                            const int IP = (int)id;
                            const int IQ = (int)((double)NP * id - (double)(IP * NP));
                            const int IHQ = (IQ * IQ + IQ) / 2;
                            if (IP == 0) {
                                const double a2 = PTSAUX[(IQ - 1) * 2 + 1];
                                HQV_(m1, IHQ) = HQV_(m1, IHQ) + vt * (a2 * a2);
                            } else if (IQ > 0) { /* (the scalar HQ(IHP) update of :2009 is the master's) */
                                const int IHP = (IP * IP + IP) / 2;
                                const double ap = PTSAUX[(IP - 1) * 2], aq = PTSAUX[(IQ - 1) * 2];
                                const int IW = (IHP > IHQ ? IHP : IHQ) - (IP > IQ ? IP - IQ : IQ - IP);
                                HQV_(m1, IHQ) = HQV_(m1, IHQ) + vt * (aq * aq);
                                HQV_(m1, IW) = HQV_(m1, IW) + vt * ap * aq;
                            }
                        }
                    }
                }
                break;
            }
            default: break;
        }
    }

    /* ---- the moment-parallel operations (all threads) ---- */
    __device__ void run_op(int op) {
        const DfCtl& c = *ctl;
        switch (op) {
            case OP_DFOVEC: { /* dfovec (objective.f90:76-123) at XC */
                if constexpr (Obj::kVector) {
                    if (tid == 0) {
                        const double pen = tk_get_penalty(*ctx, TkMemX{XC, 1});
                        ctl->d1 = pen; /* getPenalty(x), added to F by the caller (:823) */
                        ctl->d0 = (pen > TK_MYEPS) ? TK_DIV(pen, ctx->sqrt_nm) : 0.0;
                    }
                    bar();
                    if (ctl->d1 > TK_MYEPS) {
                        const double f = ctl->d0;
                        FOR_M { VSH[m1 - 1] = f; V1(v_err, m1) = f; }
                    } else {
                        Obj::dfovec_block(ctx->user, XC, VSH, YSH); /* ends with a block barrier */
                        FOR_M V1(v_err, m1) = VSH[m1 - 1];
                    }
                } else {
                    if (tid == 0) {
                        const TkMemX x{XC, 1};
                        const double pen = tk_get_penalty(*ctx, x);
                        double f;
                        if (pen > TK_MYEPS) f = TK_DIV(pen, ctx->sqrt_nm);
                        else f = tk_value_seq<Obj>(*ctx, x);
                        ctl->d0 = f;
                        ctl->d1 = pen; /* getPenalty(x), added to F by the caller (:823) */
                    }
                    bar();
                    const double f = ctl->d0;
                    FOR_M V1(v_err, m1) = f;
                }
                break;
            }
            case OP_PRELIM_ZERO:
                for (int ih = 1; ih <= NH; ih++) FOR_M HQV_(m1, ih) = 0.0;
                for (int k = 1; k <= NPT; k++) FOR_M PQV_(m1, k) = 0.0;
                FOR_M v_itest[m1 - 1] = 0;
                break;
            case OP_PRELIM_STORE: { /* :1565-1571 */
                const int NF = c.i0;
                FOR_M {
                    FVALV_(m1, NF) = V1(v_err, m1);
                    if (NF == 1) V1(v_beg, m1) = V1(v_err, m1);
                }
                break;
            }
            case OP_PRELIM_G: { /* :1581-1582 */
                const int NFM = c.i1;
                const double STEPA = c.d0;
                FOR_M GQV_(m1, NFM) = (V1(v_err, m1) - V1(v_beg, m1)) / STEPA;
                break;
            }
            case OP_PRELIM_H: { /* :1591-1595 */
                const int NFX = c.i1, IH = c.i2;
                const double STEPA = c.d0, STEPB = c.d1, DIFF = c.d2;
                FOR_M {
                    const double TEMP = (V1(v_err, m1) - V1(v_beg, m1)) / STEPB;
                    HQV_(m1, IH) = 2.0 * (TEMP - GQV_(m1, NFX)) / DIFF;
                    GQV_(m1, NFX) = (GQV_(m1, NFX) * STEPB - TEMP * STEPA) / DIFF;
                }
                break;
            }
            case OP_PRELIM_SWAP: { /* :1600-1603 */
                const int NF = c.i0;
                FOR_M {
                    FVALV_(m1, NF) = FVALV_(m1, NF - N);
                    FVALV_(m1, NF - N) = V1(v_err, m1);
                }
                break;
            }
            case OP_L20_HQV: { /* :492-504 and :994-1006: GQV += HQV * vec  (vec = XOPT or D, in TK[1..N]) */
                FOR_M {
                    int IH = 0;
                    for (int J = 1; J <= N; J++)
                        for (int I = 1; I <= J; I++) {
                            IH++;
                            if (I < J) GQV_(m1, J) = GQV_(m1, J) + HQV_(m1, IH) * V1(TK, I);
                            GQV_(m1, I) = GQV_(m1, I) + HQV_(m1, IH) * V1(TK, J);
                        }
                }
                break;
            }
            case OP_L20_PQV: { /* :506-514 and :1007-1016: TEMP_K in TK[N+K] */
                FOR_M {
                    for (int K = 1; K <= NPT; K++) {
                        const double t = PQV_(m1, K) * V1(TK, N + K);
                        for (int i = 1; i <= N; i++) GQV_(m1, i) = GQV_(m1, i) + t * XPT_(K, i);
                    }
                }
                break;
            }
            case OP_SHIFT: { /* :569-573, :611-624 with the OLD XPT (the master shifts XPT afterwards) */
                FOR_M {
                    double sp = 0.0;
                    for (int K = 1; K <= NPT; K++) sp = sp + PQV_(m1, K);
                    V1(v_sumpq, m1) = sp;
                    for (int J = 1; J <= N; J++) {
                        double wv = -0.5 * sp * V1(XOPT, J);
                        for (int K = 1; K <= NPT; K++) wv = wv + PQV_(m1, K) * XPT_(K, J);
                        WV_(m1, J) = wv;
                    }
                    int IH = 0;
                    for (int J = 1; J <= N; J++)
                        for (int I = 1; I <= J; I++) {
                            IH++;
                            HQV_(m1, IH) = HQV_(m1, IH) + WV_(m1, I) * V1(XOPT, J) + V1(XOPT, I) * WV_(m1, J);
                        }
                }
                break;
            }
            case OP_VQUAD: { /* :846-862; W(NPT+K) in TK[K] */
                const int KOPT = c.i0;
                FOR_M {
                    double vq = 0.0;
                    int IH = 0;
                    for (int J = 1; J <= N; J++) {
                        vq = vq + V1(D, J) * GQV_(m1, J);
                        for (int I = 1; I <= J; I++) {
                            IH++;
                            double TEMP = V1(D, I) * V1(D, J);
                            if (I == J) TEMP = 0.5 * TEMP;
                            vq = vq + HQV_(m1, IH) * TEMP;
                        }
                    }
                    for (int K = 1; K <= NPT; K++) vq = vq + 0.5 * PQV_(m1, K) * (V1(TK, K) * V1(TK, K)); /* HALF*PQV*W(NPT+K)**2 */
                    V1(v_vquad, m1) = vq;
                    V1(v_diff, m1) = V1(v_err, m1) - FVALV_(m1, KOPT) - vq;
                }
                break;
            }
            case OP_UPDATE_MODEL: { /* :940-968 (XPT(KNEW,.) still the OLD point, ZMAT already updated) */
                const int KNEW = c.i0;
                FOR_M {
                    const double pold = PQV_(m1, KNEW);
                    PQV_(m1, KNEW) = 0.0;
                    int IH = 0;
                    for (int I = 1; I <= N; I++) {
                        const double vt = pold * XPT_(KNEW, I);
                        for (int J = 1; J <= I; J++) {
                            IH++;
                            HQV_(m1, IH) = HQV_(m1, IH) + vt * XPT_(KNEW, J);
                        }
                    }
                    for (int JJ = 1; JJ <= NPTM; JJ++) {
                        const double vt = V1(v_diff, m1) * ZMAT_(KNEW, JJ);
                        for (int K = 1; K <= NPT; K++) PQV_(m1, K) = PQV_(m1, K) + vt * ZMAT_(K, JJ);
                    }
                    FVALV_(m1, KNEW) = V1(v_err, m1);
                }
                break;
            }
            case OP_GQV_W: /* :982-985, W(I) in TK[I] */
                FOR_M for (int I = 1; I <= N; I++) GQV_(m1, I) = GQV_(m1, I) + V1(v_diff, m1) * V1(TK, I);
                break;
            case OP_LFN: { /* :1030-1075 least-Frobenius-norm test, per moment; v_sum in TK[1..NPT] */
                const int KOPT = c.i0;
                bool upd = false;
                FOR_M {
                    double VLm[201 + 100 + 8], Wm[2 * 201 + 2]; /* NPT <= 201 (n <= 100) */
                    for (int K = 1; K <= NPT; K++) { VLm[K] = FVALV_(m1, K) - FVALV_(m1, KOPT); Wm[K] = 0.0; }
                    for (int J = 1; J <= NPTM; J++) {
                        double SUM = 0.0;
                        for (int K = 1; K <= NPT; K++) SUM = SUM + ZMAT_(K, J) * VLm[K];
                        for (int K = 1; K <= NPT; K++) Wm[K] = Wm[K] + SUM * ZMAT_(K, J);
                    }
                    for (int K = 1; K <= NPT; K++) { Wm[K + NPT] = Wm[K]; Wm[K] = V1(TK, K) * Wm[K]; }
                    double GQSQ = 0.0, GISQ = 0.0;
                    for (int I = 1; I <= N; I++) {
                        double SUM = 0.0;
                        for (int K = 1; K <= NPT; K++) SUM = SUM + BMAT_(K, I) * VLm[K] + XPT_(K, I) * Wm[K];
                        const double g = GQV_(m1, I);
                        if (fabs(V1(XOPT, I) - V1(SL, I)) <= eps) {
                            GQSQ = GQSQ + fmin(0.0, g) * fmin(0.0, g);
                            GISQ = GISQ + fmin(0.0, SUM) * fmin(0.0, SUM);
                        } else if (fabs(V1(XOPT, I) - V1(SU, I)) <= eps) {
                            GQSQ = GQSQ + fmax(0.0, g) * fmax(0.0, g);
                            GISQ = GISQ + fmax(0.0, SUM) * fmax(0.0, SUM);
                        } else {
                            GQSQ = GQSQ + g * g;
                            GISQ = GISQ + SUM * SUM;
                        }
                        VLm[NPT + I] = SUM;
                    }
                    int it = v_itest[m1 - 1] + 1;
                    if (GQSQ < 10.0 * GISQ) it = 0;
                    if (it >= 3) {
                        upd = true;
                        const int lim = NPT > NH ? NPT : NH;
                        for (int i = 1; i <= lim; i++) {
                            if (i < N) GQV_(m1, i) = VLm[NPT + i]; /* sic: i < n (:1069) */
                            if (i <= NPT) PQV_(m1, i) = Wm[NPT + i];
                            if (i <= NH) HQV_(m1, i) = 0.0;
                        }
                        it = 0;
                    }
                    v_itest[m1 - 1] = it;
                }
                if (upd) atomicExch(&ctl->model_update, 1);
                break;
            }
            case OP_TRS_MODEL: { /* TRSBOX_H :2060-2125 */
                const int KOPT = c.i0;
                double vmax = 0.0;
                FOR_M {
                    const double t = FVALV_(m1, KOPT);
                    vmax = fmax(vmax, fabs(t));
                    V1(v_temp, m1) = t;
                }
                /* max is order independent: reduce through VBUF */
                VBUF[tid] = vmax;
                bar();
                if (tid == 0) {
                    double v = 0.0;
                    for (int t = 0; t < nthr; t++) v = fmax(v, VBUF[t]);
                    ctl->d0 = v; /* voptmax */
                }
                /* GOPT(i) = 2 * sum_m v_temp(m) * GQV(m,i), index order, one thread per i */
                for (int i = nthr - tid; i <= N; i += nthr) { /* threads from the top end: tid 0 is busy above */
                    double g = 0.0;
                    for (int m1 = 1; m1 <= mv; m1++) g = g + V1(v_temp, m1) * GQV_(m1, i);
                    V1(GOPT, i) = 2.0 * g;
                }
                if (tid == 1) { /* f_base = f_value(v_temp) */
                    double f = 0.0;
                    for (int m1 = 1; m1 <= mv; m1++) f = f + V1(v_temp, m1) * V1(v_temp, m1);
                    ctl->d1 = f;
                }
                bar();
                if (tid == 0) {
                    double gnorm2 = 0.0;
                    for (int i = 1; i <= N; i++) gnorm2 = gnorm2 + V1(GOPT, i) * V1(GOPT, i);
                    const double f_base = ctl->d1;
                    int cases;
                    if (gnorm2 >= 1.0) cases = 1;
                    else if (f_base <= sqrt(gnorm2)) cases = 2;
                    else cases = 3;
                    ctl->i1 = cases;
                }
                bar();
                const int cases = ctl->i1;
                const double voptmax = ctl->d0;
                for (int IH = tid + 1; IH <= NH; IH += nthr) { /* one packed element per thread, sums in index order */
                    int j = 1;
                    while ((j * (j + 1)) / 2 < IH) j++;
                    const int i = IH - (j * (j - 1)) / 2;
                    double t1 = 0.0;
                    if (cases != 3) {
                        for (int m1 = 1; m1 <= mv; m1++) t1 = t1 + GQV_(m1, i) * GQV_(m1, j);
                        double h = 2.0 * t1;
                        if (cases == 2 && i == j) h = h + 1.0e-3 * voptmax;
                        V1(HQ, IH) = h;
                    } else {
                        for (int m1 = 1; m1 <= mv; m1++) {
                            double t2 = 0.0;
                            for (int k = 1; k <= NPT; k++) t2 = t2 + XPT_(k, i) * PQV_(m1, k) * XPT_(k, j);
                            t2 = t2 + HQV_(m1, IH);
                            t1 = t1 + (GQV_(m1, i) * GQV_(m1, j) + V1(v_temp, m1) * t2);
                        }
                        V1(HQ, IH) = 2.0 * t1;
                    }
                }
                break;
            }
            case OP_RESC_SHIFT: case OP_RESC_NEWPT: case OP_RESC_MODEL: run_op_rescue(op); break;
            case OP_ASYNC_PICK: { /* asynchronous instances: finish events in front of (Tv, s), the best visible row (kernels_nm.cu has the same pass) */
                const tknm::NmAsync& q = *aq;
                const int Tv = c.i0, s_ = c.i1, P = c.i2;
                const int lane = tid & 31, wp = tid >> 5, nw = (nthr + 31) >> 5;
                int cnt_ev = 0, cnt_vis = 0, bt = 0, bs = 0, bk = -1;
                double bf = 0.0;
                const int hi = min(q.last_k, __ldcg(q.clk + P));
                for (int jj = tid; jj <= hi; jj += nthr) {
                    const int4 raw = __ldcg(reinterpret_cast<const int4*>(q.ev + jj));
                    const int tj = raw.z, sj = raw.w;
                    if (tj < 0 || tj > Tv) continue;
                    if (jj >= q.first_k && (tj < Tv || sj < s_)) cnt_ev++;
                    const double fj = __hiloint2double(raw.y, raw.x);
                    if (!(fj == fj)) continue;
                    cnt_vis++;
                    if (bk < 0 || tknm::nm_row_before(fj, tj, sj, jj, bf, bt, bs, bk)) { bf = fj; bt = tj; bs = sj; bk = jj; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double f2 = __shfl_xor_sync(0xffffffffu, bf, o);
                    const int t2 = __shfl_xor_sync(0xffffffffu, bt, o), s2 = __shfl_xor_sync(0xffffffffu, bs, o), k2 = __shfl_xor_sync(0xffffffffu, bk, o);
                    if (k2 >= 0 && (bk < 0 || tknm::nm_row_before(f2, t2, s2, k2, bf, bt, bs, bk))) { bf = f2; bt = t2; bs = s2; bk = k2; }
                    cnt_ev += __shfl_xor_sync(0xffffffffu, cnt_ev, o);
                    cnt_vis += __shfl_xor_sync(0xffffffffu, cnt_vis, o);
                }
                if (lane == 0) { pk_f[wp] = bf; pk_t[wp] = bt; pk_s[wp] = bs; pk_k[wp] = bk; pk_v[wp] = cnt_vis; pk_e[wp] = cnt_ev; }
                bar();
                if (tid == 0) {
                    bf = pk_f[0]; bt = pk_t[0]; bs = pk_s[0]; bk = pk_k[0]; cnt_vis = pk_v[0]; cnt_ev = pk_e[0];
                    for (int o = 1; o < nw; o++) {
                        if (pk_k[o] >= 0 && (bk < 0 || tknm::nm_row_before(pk_f[o], pk_t[o], pk_s[o], pk_k[o], bf, bt, bs, bk))) { bf = pk_f[o]; bt = pk_t[o]; bs = pk_s[o]; bk = pk_k[o]; }
                        cnt_vis += pk_v[o]; cnt_ev += pk_e[o];
                    }
                    ctl->i3 = cnt_ev; ctl->i0 = cnt_vis; ctl->i1 = bk;
                }
                break;
            }
            case OP_FINAL:
            default: break;
        }
    }

    __device__ void par(int op) { /* master side */
        ctl->op = op;
        bar();
        run_op(op);
        bar();
    }
    __device__ void worker_loop() {
        for (;;) {
            bar();
            const int op = ctl->op;
            if (op == OP_EXIT) return;
            run_op(op);
            bar();
        }
    }

    /* F = sum v_err^2 in index order (f_value :1173-1183) + getPenalty; master only, after OP_DFOVEC */
    __device__ double eval_at_XC() {
        a_nev++;
        if (a_clk) tknm::st_vol(a_clk, a_t0 + a_nev); /* the restart cannot finish before this evaluation has */
        par(OP_DFOVEC);
        double F = 0.0;
        if constexpr (Obj::kVector) {
            for (int m1 = 0; m1 < mv; m1++) F = F + VSH[m1] * VSH[m1];
        } else {
            const double f = ctl->d0;
            for (int m1 = 1; m1 <= mv; m1++) F = F + f * f; /* all moments equal for scalar plugins: same sum */
        }
        return F + ctl->d1;
    }

    /* ---------------------------------------------------------------- UPDATE :2492-2563 (master) */
    __device__ void UPDATE(double* VLp, double BETA, double DENOM, int KNEW, double* Wp) {
        double* VL = VLp - 1; double* Wk = Wp - 1;
        const int nptm = NPT - N - 1;
        double ZTEST = 0.0;
        for (int K = 1; K <= NPT; K++)
            for (int J = 1; J <= nptm; J++) ZTEST = fmax(ZTEST, fabs(ZMAT_(K, J)));
        ZTEST = 1.0e-20 * ZTEST;
        for (int J = 2; J <= nptm; J++) {
            if (fabs(ZMAT_(KNEW, J)) > ZTEST) {
                double TEMP = sqrt(ZMAT_(KNEW, 1) * ZMAT_(KNEW, 1) + ZMAT_(KNEW, J) * ZMAT_(KNEW, J));
                const double TEMPA = ZMAT_(KNEW, 1) / TEMP;
                const double TEMPB = ZMAT_(KNEW, J) / TEMP;
                for (int I = 1; I <= NPT; I++) {
                    TEMP = TEMPA * ZMAT_(I, 1) + TEMPB * ZMAT_(I, J);
                    ZMAT_(I, J) = TEMPA * ZMAT_(I, J) - TEMPB * ZMAT_(I, 1);
                    ZMAT_(I, 1) = TEMP;
                }
            }
            ZMAT_(KNEW, J) = 0.0;
        }
        for (int I = 1; I <= NPT; I++) Wk[I] = ZMAT_(KNEW, 1) * ZMAT_(I, 1);
        const double ALPHA = Wk[KNEW];
        const double TAU = VL[KNEW];
        VL[KNEW] = VL[KNEW] - 1.0;
        double TEMP = sqrt(DENOM);
        double TEMPB = ZMAT_(KNEW, 1) / TEMP;
        double TEMPA = TAU / TEMP;
        for (int I = 1; I <= NPT; I++) ZMAT_(I, 1) = TEMPA * ZMAT_(I, 1) - TEMPB * VL[I];
        for (int J = 1; J <= N; J++) {
            const int JP = NPT + J;
            Wk[JP] = BMAT_(KNEW, J);
            TEMPA = (ALPHA * VL[JP] - TAU * Wk[JP]) / DENOM;
            TEMPB = (-BETA * Wk[JP] - TAU * VL[JP]) / DENOM;
            for (int I = 1; I <= JP; I++) {
                BMAT_(I, J) = BMAT_(I, J) + TEMPA * VL[I] + TEMPB * Wk[I];
                if (I > NPT) BMAT_(JP, I - NPT) = BMAT_(I, J);
            }
        }
    }

    /* ---------------------------------------------------------------- PRELIM_H :1473-1619 (master) */
    __device__ void PRELIM_H(double RHOBEG, int MAXFUN, int& NF, int& KOPT, double& FBEG) {
        const double HALF = 0.5, ONE = 1.0, TWO = 2.0, ZERO = 0.0;
        const double RHOSQ = RHOBEG * RHOBEG;
        double STEPA = 0, STEPB = 0, F = 0;
        for (int J = 1; J <= N; J++) {
            V1(XBASE, J) = V1(XC, J);
            for (int K = 1; K <= NPT; K++) XPT_(K, J) = ZERO;
            for (int I = 1; I <= NDIM; I++) BMAT_(I, J) = ZERO;
        }
        for (int IH = 1; IH <= NH; IH++) V1(HQ, IH) = ZERO;
        for (int K = 1; K <= NPT; K++)
            for (int J = 1; J <= NPT - NP; J++) ZMAT_(K, J) = ZERO;
        par(OP_PRELIM_ZERO);
        NF = 0;
        for (;;) {
            const int NFM = NF, NFX = NF - N;
            NF = NF + 1;
            if (NFM <= 2 * N) {
                if (NFM >= 1 && NFM <= N) {
                    STEPA = RHOBEG;
                    if (fabs(V1(SU, NFM)) <= eps) STEPA = -STEPA;
                    XPT_(NF, NFM) = STEPA;
                } else if (NFM > N) {
                    STEPA = XPT_(NF - N, NFX);
                    STEPB = -RHOBEG;
                    if (fabs(V1(SL, NFX)) <= eps) STEPB = fmin(TWO * RHOBEG, V1(SU, NFX));
                    if (fabs(V1(SU, NFX)) <= eps) STEPB = fmax(-TWO * RHOBEG, V1(SL, NFX));
                    XPT_(NF, NFX) = STEPB;
                }
            }
            for (int J = 1; J <= N; J++) {
                V1(XC, J) = fmin(fmax(V1(XL, J), V1(XBASE, J) + XPT_(NF, J)), V1(XU, J));
                if (fabs(XPT_(NF, J) - V1(SL, J)) <= eps) V1(XC, J) = V1(XL, J);
                if (fabs(XPT_(NF, J) - V1(SU, J)) <= eps) V1(XC, J) = V1(XU, J);
            }
            F = eval_at_XC();
            V1(FVAL, NF) = F;
            ctl->i0 = NF;
            par(OP_PRELIM_STORE);
            if (NF == 1) { FBEG = F; KOPT = 1; }
            else if (F < V1(FVAL, KOPT)) KOPT = NF;
            if (NF <= 2 * N + 1) {
                if (NF >= 2 && NF <= N + 1) {
                    ctl->i1 = NFM; ctl->d0 = STEPA;
                    par(OP_PRELIM_G);
                    if (NPT < NF + N) {
                        BMAT_(1, NFM) = -ONE / STEPA;
                        BMAT_(NF, NFM) = ONE / STEPA;
                        BMAT_(NPT + NFM, NFM) = -HALF * RHOSQ;
                    }
                } else if (NF >= N + 2) {
                    const int IH = (NFX * (NFX + 1)) / 2;
                    const double DIFF = STEPB - STEPA;
                    ctl->i1 = NFX; ctl->i2 = IH; ctl->d0 = STEPA; ctl->d1 = STEPB; ctl->d2 = DIFF;
                    par(OP_PRELIM_H);
                    if (STEPA * STEPB < ZERO) {
                        if (F < V1(FVAL, NF - N)) {
                            V1(FVAL, NF) = V1(FVAL, NF - N);
                            V1(FVAL, NF - N) = F;
                            ctl->i0 = NF;
                            par(OP_PRELIM_SWAP);
                            if (KOPT == NF) KOPT = NF - N;
                            XPT_(NF - N, NFX) = STEPB;
                            XPT_(NF, NFX) = STEPA;
                        }
                    }
                    BMAT_(1, NFX) = -(STEPA + STEPB) / (STEPA * STEPB);
                    BMAT_(NF, NFX) = -HALF / XPT_(NF - N, NFX);
                    BMAT_(NF - N, NFX) = -BMAT_(1, NFX) - BMAT_(NF, NFX);
                    ZMAT_(1, NFX) = sqrt(TWO) / (STEPA * STEPB);
                    ZMAT_(NF, NFX) = sqrt(HALF) / RHOSQ;
                    ZMAT_(NF - N, NFX) = -ZMAT_(1, NFX) - ZMAT_(NF, NFX);
                }
            }
            if (NF < NPT && NF < MAXFUN) continue;
            break;
        }
    }

    /* ---------------------------------------------------------------- ALTMOV :1186-1464 (master) */
    __device__ void ALTMOV(int KOPT, int KNEW, double ADELT, double& ALPHA, double& CAUCHY, double* GLAGp, double* HCOLp, double* Wp) {
        double* GLAG = GLAGp - 1; double* HCOL = HCOLp - 1; double* Wk = Wp - 1;
        const double HALF = 0.5, ONE = 1.0, ZERO = 0.0;
        const double CONST = ONE + sqrt(2.0);
        double TEMP, HA, PRESAV, DDERIV, DISTSQ, SUBD, SLBD, SUMIN, DIFF, STEP = 0, VLAGs, TEMPD, TEMPA, TEMPB, PREDSQ;
        double STPSAV = 0, BIGSTP, WFIXSQ, GGFREE, WSQSAV, GW, CURV, SCALE, CSAVE = 0;
        int ILBD, IUBD, ISBD, KSAV = 0, IBDSAV = 0, IFLAG;
        for (int K = 1; K <= NPT; K++) HCOL[K] = ZERO;
        for (int J = 1; J <= NPT - N - 1; J++) {
            TEMP = ZMAT_(KNEW, J);
            for (int K = 1; K <= NPT; K++) HCOL[K] = HCOL[K] + TEMP * ZMAT_(K, J);
        }
        ALPHA = HCOL[KNEW];
        HA = HALF * ALPHA;
        for (int I = 1; I <= N; I++) GLAG[I] = BMAT_(KNEW, I);
        for (int K = 1; K <= NPT; K++) {
            TEMP = ZERO;
            for (int J = 1; J <= N; J++) TEMP = TEMP + XPT_(K, J) * V1(XOPT, J);
            TEMP = HCOL[K] * TEMP;
            for (int I = 1; I <= N; I++) GLAG[I] = GLAG[I] + TEMP * XPT_(K, I);
        }
        PRESAV = ZERO;
        for (int K = 1; K <= NPT; K++) {
            if (K == KOPT) continue;
            DDERIV = ZERO;
            DISTSQ = ZERO;
            for (int I = 1; I <= N; I++) {
                TEMP = XPT_(K, I) - V1(XOPT, I);
                DDERIV = DDERIV + GLAG[I] * TEMP;
                DISTSQ = DISTSQ + TEMP * TEMP;
            }
            SUBD = ADELT / sqrt(DISTSQ);
            SLBD = -SUBD;
            ILBD = 0;
            IUBD = 0;
            SUMIN = fmin(ONE, SUBD);
            for (int I = 1; I <= N; I++) {
                TEMP = XPT_(K, I) - V1(XOPT, I);
                if (TEMP > ZERO) {
                    if (SLBD * TEMP < V1(SL, I) - V1(XOPT, I)) { SLBD = (V1(SL, I) - V1(XOPT, I)) / TEMP; ILBD = -I; }
                    if (SUBD * TEMP > V1(SU, I) - V1(XOPT, I)) { SUBD = fmax(SUMIN, (V1(SU, I) - V1(XOPT, I)) / TEMP); IUBD = I; }
                } else if (TEMP < ZERO) {
                    if (SLBD * TEMP > V1(SU, I) - V1(XOPT, I)) { SLBD = (V1(SU, I) - V1(XOPT, I)) / TEMP; ILBD = I; }
                    if (SUBD * TEMP < V1(SL, I) - V1(XOPT, I)) { SUBD = fmax(SUMIN, (V1(SL, I) - V1(XOPT, I)) / TEMP); IUBD = -I; }
                }
            }
            if (K == KNEW) {
                DIFF = DDERIV - ONE;
                STEP = SLBD;
                VLAGs = SLBD * (DDERIV - SLBD * DIFF);
                ISBD = ILBD;
                TEMP = SUBD * (DDERIV - SUBD * DIFF);
                if (fabs(TEMP) > fabs(VLAGs)) { STEP = SUBD; VLAGs = TEMP; ISBD = IUBD; }
                TEMPD = HALF * DDERIV;
                TEMPA = TEMPD - DIFF * SLBD;
                TEMPB = TEMPD - DIFF * SUBD;
                if (TEMPA * TEMPB < ZERO) {
                    TEMP = TEMPD * TEMPD / DIFF;
                    if (fabs(TEMP) > fabs(VLAGs)) { STEP = TEMPD / DIFF; VLAGs = TEMP; ISBD = 0; }
                }
            } else {
                STEP = SLBD;
                VLAGs = SLBD * (ONE - SLBD);
                ISBD = ILBD;
                TEMP = SUBD * (ONE - SUBD);
                if (fabs(TEMP) > fabs(VLAGs)) { STEP = SUBD; VLAGs = TEMP; ISBD = IUBD; }
                if (SUBD > HALF) {
                    if (fabs(VLAGs) < 0.25) { STEP = HALF; VLAGs = 0.25; ISBD = 0; }
                }
                VLAGs = VLAGs * DDERIV;
            }
            TEMP = STEP * (ONE - STEP) * DISTSQ;
            PREDSQ = VLAGs * VLAGs * (VLAGs * VLAGs + HA * TEMP * TEMP);
            if (PREDSQ > PRESAV) { PRESAV = PREDSQ; KSAV = K; STPSAV = STEP; IBDSAV = ISBD; }
        }
        for (int I = 1; I <= N; I++) {
            TEMP = V1(XOPT, I) + STPSAV * (XPT_(KSAV, I) - V1(XOPT, I));
            V1(XNEW, I) = fmax(V1(SL, I), fmin(V1(SU, I), TEMP));
        }
        if (IBDSAV < 0) V1(XNEW, -IBDSAV) = V1(SL, -IBDSAV);
        if (IBDSAV > 0) V1(XNEW, IBDSAV) = V1(SU, IBDSAV);
        BIGSTP = ADELT + ADELT;
        IFLAG = 0;
    L100:
        WFIXSQ = ZERO;
        GGFREE = ZERO;
        for (int I = 1; I <= N; I++) {
            Wk[I] = ZERO;
            TEMPA = fmin(V1(XOPT, I) - V1(SL, I), GLAG[I]);
            TEMPB = fmax(V1(XOPT, I) - V1(SU, I), GLAG[I]);
            if (TEMPA > ZERO || TEMPB < ZERO) { Wk[I] = BIGSTP; GGFREE = GGFREE + GLAG[I] * GLAG[I]; }
        }
        if (fabs(GGFREE) <= eps) { CAUCHY = ZERO; return; }
    L120:
        TEMP = ADELT * ADELT - WFIXSQ;
        if (TEMP > ZERO) {
            WSQSAV = WFIXSQ;
            STEP = sqrt(TEMP / GGFREE);
            GGFREE = ZERO;
            for (int I = 1; I <= N; I++) {
                if (fabs(Wk[I] - BIGSTP) <= eps) {
                    TEMP = V1(XOPT, I) - STEP * GLAG[I];
                    if (TEMP <= V1(SL, I)) { Wk[I] = V1(SL, I) - V1(XOPT, I); WFIXSQ = WFIXSQ + Wk[I] * Wk[I]; }
                    else if (TEMP >= V1(SU, I)) { Wk[I] = V1(SU, I) - V1(XOPT, I); WFIXSQ = WFIXSQ + Wk[I] * Wk[I]; }
                    else GGFREE = GGFREE + GLAG[I] * GLAG[I];
                }
            }
            if (WFIXSQ > WSQSAV && GGFREE > ZERO) goto L120;
        }
        GW = ZERO;
        for (int I = 1; I <= N; I++) {
            if (fabs(Wk[I] - BIGSTP) <= eps) {
                Wk[I] = -STEP * GLAG[I];
                V1(XALT, I) = fmax(V1(SL, I), fmin(V1(SU, I), V1(XOPT, I) + Wk[I]));
            } else if (fabs(Wk[I]) <= eps) {
                V1(XALT, I) = V1(XOPT, I);
            } else if (GLAG[I] > ZERO) {
                V1(XALT, I) = V1(SL, I);
            } else {
                V1(XALT, I) = V1(SU, I);
            }
            GW = GW + GLAG[I] * Wk[I];
        }
        CURV = ZERO;
        for (int K = 1; K <= NPT; K++) {
            TEMP = ZERO;
            for (int J = 1; J <= N; J++) TEMP = TEMP + XPT_(K, J) * Wk[J];
            CURV = CURV + HCOL[K] * TEMP * TEMP;
        }
        if (IFLAG == 1) CURV = -CURV;
        if (CURV > -GW && CURV < -CONST * GW) {
            SCALE = -GW / CURV;
            for (int I = 1; I <= N; I++) {
                TEMP = V1(XOPT, I) + SCALE * Wk[I];
                V1(XALT, I) = fmax(V1(SL, I), fmin(V1(SU, I), TEMP));
            }
            CAUCHY = (HALF * GW * SCALE) * (HALF * GW * SCALE);
        } else {
            CAUCHY = (GW + HALF * CURV) * (GW + HALF * CURV);
        }
        if (IFLAG == 0) {
            for (int I = 1; I <= N; I++) { GLAG[I] = -GLAG[I]; Wk[N + I] = V1(XALT, I); }
            CSAVE = CAUCHY;
            IFLAG = 1;
            goto L100;
        }
        if (CSAVE > CAUCHY) {
            for (int I = 1; I <= N; I++) V1(XALT, I) = Wk[N + I];
            CAUCHY = CSAVE;
        }
    }

    /* ---------------------------------------------------------------- TRSBOX_H :2032-2488 (master) */
    __device__ void TRSBOX_H(double DELTA, double* GNEWp, double* XBDIp, double* Sp, double* HSp, double* HREDp, double& DSQ,
                             double& CRVMIN, int KOPT, double& vquad) {
        double* GNEW = GNEWp - 1; double* XBDI = XBDIp - 1; double* S = Sp - 1; double* HS = HSp - 1; double* HRED = HREDp - 1;
        int IH;
        double HALF, ONE, ONEMIN, ZERO, DELSQ, QRED, BETA, STEPSQ, GREDSQ = 0, gredsq0 = 0, RESID, DS, SHS = 0, TEMP, BLEN = 0, STPLEN,
            XSUM, SDEC, GGSAV = 0, DREDSQ = 0, DREDG = 0, SREDG = 0, ANGBD = 0, TEMPA, TEMPB, SSQ, XSAV = 0, DHS, DHD, REDMAX, REDSAV,
            ANGT = 0, STH, REDNEW, RDPREV = 0, RDNEXT = 0, CTH;
        int ITERC, NACT, ITERMAX = 0, IACT = 0, ITCSAV = 0, ISAV, IU;
        if (model_update || opt_update) {
            model_update = false;
            opt_update = false;
            ctl->i0 = KOPT;
            par(OP_TRS_MODEL);
        }
        HALF = 0.5; ONE = 1.0; ONEMIN = -1.0; ZERO = 0.0;
        ITERC = 0;
        NACT = 0;
        for (int I = 1; I <= N; I++) {
            XBDI[I] = ZERO;
            if (V1(XOPT, I) <= V1(SL, I)) { if (V1(GOPT, I) >= ZERO) XBDI[I] = ONEMIN; }
            else if (V1(XOPT, I) >= V1(SU, I)) { if (V1(GOPT, I) <= ZERO) XBDI[I] = ONE; }
            if (fabs(XBDI[I]) > eps) NACT = NACT + 1;
            V1(D, I) = ZERO;
            GNEW[I] = V1(GOPT, I);
        }
        DELSQ = DELTA * DELTA;
        QRED = ZERO;
        CRVMIN = ONEMIN;
    L20:
        BETA = ZERO;
    L30:
        STEPSQ = ZERO;
        for (int I = 1; I <= N; I++) {
            if (fabs(XBDI[I]) > eps) S[I] = ZERO;
            else if (fabs(BETA) <= eps) S[I] = -GNEW[I];
            else S[I] = BETA * S[I] - GNEW[I];
            STEPSQ = STEPSQ + S[I] * S[I];
        }
        if (fabs(STEPSQ) <= eps) goto L190;
        if (fabs(BETA) <= eps) { GREDSQ = STEPSQ; ITERMAX = ITERC + N - NACT; }
        if (ITERC == 0) gredsq0 = GREDSQ;
        if (GREDSQ <= fmin(1.0e-6 * gredsq0, 1.0e-18)) goto L190;
        if (GREDSQ <= 1.0e-14 * gredsq0) goto L190;
        if (GREDSQ * DELSQ <= fmin(1.0e-6 * QRED * QRED, 1.0e-18)) goto L190;
        if (GREDSQ * DELSQ <= 1.0e-14 * QRED * QRED) goto L190;
        goto L210;
    L50:
        RESID = DELSQ;
        DS = ZERO;
        SHS = ZERO;
        for (int I = 1; I <= N; I++) {
            if (fabs(XBDI[I]) <= eps) {
                RESID = RESID - V1(D, I) * V1(D, I);
                DS = DS + S[I] * V1(D, I);
                SHS = SHS + S[I] * HS[I];
            }
        }
        if (RESID <= ZERO) goto L90;
        TEMP = sqrt(STEPSQ * RESID + DS * DS);
        if (DS < ZERO) BLEN = (TEMP - DS) / STEPSQ;
        else BLEN = RESID / (TEMP + DS);
        STPLEN = BLEN;
        if (SHS > ZERO) STPLEN = fmin(BLEN, GREDSQ / SHS);
        if (STPLEN <= (double)1.e-30f) goto L190;
        IACT = 0;
        for (int I = 1; I <= N; I++) {
            if (fabs(S[I]) > eps) {
                XSUM = V1(XOPT, I) + V1(D, I);
                if (S[I] > ZERO) TEMP = (V1(SU, I) - XSUM) / S[I];
                else TEMP = (V1(SL, I) - XSUM) / S[I];
                if (TEMP < STPLEN) { STPLEN = TEMP; IACT = I; }
            }
        }
        SDEC = ZERO;
        if (STPLEN > ZERO) {
            ITERC = ITERC + 1;
            TEMP = SHS / STEPSQ;
            if (IACT == 0 && TEMP > ZERO) {
                CRVMIN = fmin(CRVMIN, TEMP);
                if (fabs(CRVMIN - ONEMIN) <= eps) CRVMIN = TEMP;
            }
            GGSAV = GREDSQ;
            GREDSQ = ZERO;
            for (int I = 1; I <= N; I++) {
                GNEW[I] = GNEW[I] + STPLEN * HS[I];
                if (fabs(XBDI[I]) <= eps) GREDSQ = GREDSQ + GNEW[I] * GNEW[I];
                V1(D, I) = V1(D, I) + STPLEN * S[I];
            }
            SDEC = fmax(STPLEN * (GGSAV - HALF * STPLEN * SHS), ZERO);
            QRED = QRED + SDEC;
        }
        if (IACT > 0) {
            NACT = NACT + 1;
            XBDI[IACT] = ONE;
            if (S[IACT] < ZERO) XBDI[IACT] = ONEMIN;
            DELSQ = DELSQ - V1(D, IACT) * V1(D, IACT);
            if (DELSQ <= ZERO) goto L90;
            goto L20;
        }
        if (STPLEN < BLEN) {
            if (ITERC == ITERMAX) goto L190;
            if (SDEC <= 1.0e-6 * QRED) goto L190;
            BETA = GREDSQ / GGSAV;
            goto L30;
        }
    L90:
        CRVMIN = ZERO;
    L100:
        if (NACT >= N - 1) goto L190;
        DREDSQ = ZERO;
        DREDG = ZERO;
        GREDSQ = ZERO;
        for (int I = 1; I <= N; I++) {
            if (fabs(XBDI[I]) <= eps) {
                DREDSQ = DREDSQ + V1(D, I) * V1(D, I);
                DREDG = DREDG + V1(D, I) * GNEW[I];
                GREDSQ = GREDSQ + GNEW[I] * GNEW[I];
                S[I] = V1(D, I);
            } else S[I] = ZERO;
        }
        ITCSAV = ITERC;
        goto L210;
    L120:
        ITERC = ITERC + 1;
        TEMP = GREDSQ * DREDSQ - DREDG * DREDG;
        if (TEMP <= 1.0e-4 * QRED * QRED) goto L190;
        TEMP = sqrt(TEMP);
        for (int I = 1; I <= N; I++) {
            if (fabs(XBDI[I]) <= eps) S[I] = (DREDG * V1(D, I) - DREDSQ * GNEW[I]) / TEMP;
            else S[I] = ZERO;
        }
        SREDG = -TEMP;
        ANGBD = ONE;
        IACT = 0;
        for (int I = 1; I <= N; I++) {
            if (fabs(XBDI[I]) <= eps) {
                TEMPA = V1(XOPT, I) + V1(D, I) - V1(SL, I);
                TEMPB = V1(SU, I) - V1(XOPT, I) - V1(D, I);
                if (TEMPA <= ZERO) { NACT = NACT + 1; XBDI[I] = ONEMIN; goto L100; }
                else if (TEMPB <= ZERO) { NACT = NACT + 1; XBDI[I] = ONE; goto L100; }
                SSQ = V1(D, I) * V1(D, I) + S[I] * S[I];
                TEMP = SSQ - (V1(XOPT, I) - V1(SL, I)) * (V1(XOPT, I) - V1(SL, I));
                if (TEMP > ZERO) {
                    TEMP = sqrt(TEMP) - S[I];
                    if (ANGBD * TEMP > TEMPA) { ANGBD = TEMPA / TEMP; IACT = I; XSAV = ONEMIN; }
                }
                TEMP = SSQ - (V1(SU, I) - V1(XOPT, I)) * (V1(SU, I) - V1(XOPT, I));
                if (TEMP > ZERO) {
                    TEMP = sqrt(TEMP) + S[I];
                    if (ANGBD * TEMP > TEMPB) { ANGBD = TEMPB / TEMP; IACT = I; XSAV = ONE; }
                }
            }
        }
        goto L210;
    L150:
        SHS = ZERO;
        DHS = ZERO;
        DHD = ZERO;
        for (int I = 1; I <= N; I++) {
            if (fabs(XBDI[I]) <= eps) {
                SHS = SHS + S[I] * HS[I];
                DHS = DHS + V1(D, I) * HS[I];
                DHD = DHD + V1(D, I) * HRED[I];
            }
        }
        REDMAX = ZERO;
        ISAV = 0;
        REDSAV = ZERO;
        IU = (int)(17.0 * ANGBD + 3.1);
        for (int I = 1; I <= IU; I++) {
            ANGT = ANGBD * (double)I / (double)IU;
            STH = (ANGT + ANGT) / (ONE + ANGT * ANGT);
            TEMP = SHS + ANGT * (ANGT * DHD - DHS - DHS);
            REDNEW = STH * (ANGT * DREDG - SREDG - HALF * STH * TEMP);
            if (REDNEW > REDMAX) { REDMAX = REDNEW; ISAV = I; RDPREV = REDSAV; }
            else if (I == ISAV + 1) RDNEXT = REDNEW;
            REDSAV = REDNEW;
        }
        if (ISAV == 0) goto L190;
        if (ISAV < IU) {
            TEMP = (RDNEXT - RDPREV) / (REDMAX + REDMAX - RDPREV - RDNEXT);
            ANGT = ANGBD * ((double)ISAV + HALF * TEMP) / (double)IU;
        }
        CTH = (ONE - ANGT * ANGT) / (ONE + ANGT * ANGT);
        STH = (ANGT + ANGT) / (ONE + ANGT * ANGT);
        TEMP = SHS + ANGT * (ANGT * DHD - DHS - DHS);
        SDEC = STH * (ANGT * DREDG - SREDG - HALF * STH * TEMP);
        if (SDEC <= ZERO) goto L190;
        DREDG = ZERO;
        GREDSQ = ZERO;
        for (int I = 1; I <= N; I++) {
            GNEW[I] = GNEW[I] + (CTH - ONE) * HRED[I] + STH * HS[I];
            if (fabs(XBDI[I]) <= eps) {
                V1(D, I) = CTH * V1(D, I) + STH * S[I];
                DREDG = DREDG + V1(D, I) * GNEW[I];
                GREDSQ = GREDSQ + GNEW[I] * GNEW[I];
            }
            HRED[I] = CTH * HRED[I] + STH * HS[I];
        }
        QRED = QRED + SDEC;
        if (IACT > 0 && ISAV == IU) { NACT = NACT + 1; XBDI[IACT] = XSAV; goto L100; }
        if (SDEC > 0.01 * QRED) goto L120;
    L190:
        DSQ = ZERO;
        for (int I = 1; I <= N; I++) {
            V1(XNEW, I) = fmax(fmin(V1(XOPT, I) + V1(D, I), V1(SU, I)), V1(SL, I));
            if (fabs(XBDI[I] - ONEMIN) <= eps) V1(XNEW, I) = V1(SL, I);
            if (fabs(XBDI[I] - ONE) <= eps) V1(XNEW, I) = V1(SU, I);
            V1(D, I) = V1(XNEW, I) - V1(XOPT, I);
            DSQ = DSQ + V1(D, I) * V1(D, I);
        }
        for (int i = 1; i <= N; i++) V1(HD1, i) = 0.0;
        IH = 0;
        for (int J = 1; J <= N; J++)
            for (int I = 1; I <= J; I++) {
                IH = IH + 1;
                if (I < J) V1(HD1, J) = V1(HD1, J) + V1(HQ, IH) * V1(D, I);
                V1(HD1, I) = V1(HD1, I) + V1(HQ, IH) * V1(D, J);
            }
        vquad = 0.0;
        for (int i = 1; i <= N; i++) vquad = vquad + V1(D, i) * (V1(GOPT, i) + 0.5 * V1(HD1, i));
        return;
    L210:
        IH = 0;
        for (int J = 1; J <= N; J++) {
            HS[J] = ZERO;
            for (int I = 1; I <= J; I++) {
                IH = IH + 1;
                if (I < J) HS[J] = HS[J] + V1(HQ, IH) * S[I];
                HS[I] = HS[I] + V1(HQ, IH) * S[J];
            }
        }
        if (fabs(CRVMIN) > eps) goto L50;
        if (ITERC > ITCSAV) goto L150;
        for (int I = 1; I <= N; I++) HRED[I] = HS[I];
        goto L120;
    }

    /* test hook (DfArgs.force_first > 0): the denominator tests :746 / :782 count as failed when NF = force_first + k force_every */
    __device__ bool forced(int NF) { /* once per NF: the tests are met again, with the same NF, right after the rescue */
        if (!(force_first > 0 && NF >= force_first && (NF - force_first) % force_every == 0) || NF == forced_at) return false;
        forced_at = NF;
        return true;
    }

    /* ---------------------------------------------------------------- RESCUE_H :1627-2023 (master)
     * Rebuilds BMAT and ZMAT from scratch around XOPT: a provisional set of points on the coordinate axes (identified by
     * PTSID: p + q/(n+1) + 1/(2n+2) encodes the two axes p and q) is exchanged, one UPDATE at a time, for as many of the
     * original points as keep the denominators healthy; the provisional points that survive replace the original ones
     * and cost one objective evaluation each.  PTSAUX(2,n), PTSID(npt) and the scratch vector are slices of the caller's
     * W, laid out as in the reference's call (:647-650) so that OP_RESC_MODEL finds them at fixed offsets.  The two
     * oddities of the reference's moment loop are kept: PQV is MULTIPLIED by the correction (:1999), and the scalar
     * HQ(IHP) takes a stale TEMP (:2009; HQ is rebuilt from the moments before its next use). */
    __device__ __noinline__ void RESCUE_H(int MAXFUN, int& NF, double DELTA, int& KOPT) {
        double* PA = W;                    /* PTSAUX(i,j) = PA[(j-1)*2 + (i-1)] */
        double* ID = W + (N + NP) - 2;     /* PTSID[k], 1-based */
        double* Wr = W + (NDIM + NP) - 2;  /* the routine's own W, 1-based */
        double* VL = VLAG - 1;
        const double sfrac = 0.5 / (double)NP;
        nrescue++;
        /* shift the points; squared distances from XOPT at the end of Wr */
        double winc = 0.0;
        for (int K = 1; K <= NPT; K++) {
            double dist = 0.0;
            for (int J = 1; J <= N; J++) {
                XPT_(K, J) = XPT_(K, J) - V1(XOPT, J);
                dist = dist + XPT_(K, J) * XPT_(K, J);
            }
            Wr[NDIM + K] = dist;
            winc = fmax(winc, dist);
            for (int J = 1; J <= NPTM; J++) ZMAT_(K, J) = 0.0;
        }
        ctl->i0 = KOPT;
        par(OP_RESC_SHIFT);
        for (int J = 1; J <= N; J++) {
            V1(XBASE, J) = V1(XBASE, J) + V1(XOPT, J);
            V1(SL, J) = V1(SL, J) - V1(XOPT, J);
            V1(SU, J) = V1(SU, J) - V1(XOPT, J);
            V1(XOPT, J) = 0.0;
            double a1 = fmin(DELTA, V1(SU, J)), a2 = fmax(-DELTA, V1(SL, J));
            if (a1 + a2 < 0.0) { const double sw = a1; a1 = a2; a2 = sw; }
            if (fabs(a2) < 0.5 * fabs(a1)) a2 = 0.5 * a1;
            PA[(J - 1) * 2] = a1; PA[(J - 1) * 2 + 1] = a2;
            for (int I = 1; I <= NDIM; I++) BMAT_(I, J) = 0.0;
        }
        /* identifiers and nonzero elements of BMAT / ZMAT for the provisional points */
        ID[1] = sfrac;
        for (int J = 1; J <= N; J++) {
            const int JP = J + 1, JPN = JP + N;
            const double a1 = PA[(J - 1) * 2], a2 = PA[(J - 1) * 2 + 1];
            ID[JP] = (double)J + sfrac;
            if (JPN <= NPT) {
                ID[JPN] = (double)J / (double)NP + sfrac;
                const double tinv = 1.0 / (a1 - a2);
                BMAT_(JP, J) = -tinv + 1.0 / a1;
                BMAT_(JPN, J) = tinv + 1.0 / a2;
                BMAT_(1, J) = -BMAT_(JP, J) - BMAT_(JPN, J);
                ZMAT_(1, J) = sqrt(2.0) / fabs(a1 * a2);
                ZMAT_(JP, J) = ZMAT_(1, J) * a2 * tinv;
                ZMAT_(JPN, J) = -ZMAT_(1, J) * a1 * tinv;
            } else {
                BMAT_(1, J) = -1.0 / a1;
                BMAT_(JP, J) = 1.0 / a1;
                BMAT_(J + NPT, J) = -0.5 * (a1 * a1);
            }
        }
        if (NPT >= N + NP) {
            for (int K = 2 * NP; K <= NPT; K++) {
                const int iw = (int)(((double)(K - NP) - 0.5) / (double)N);
                const int ip = K - NP - iw * N;
                int iq = ip + iw;
                if (iq > N) iq = iq - N;
                ID[K] = (double)ip + (double)iq / (double)NP + sfrac;
                const double tinv = 1.0 / (PA[(ip - 1) * 2] * PA[(iq - 1) * 2]);
                ZMAT_(1, K - NP) = tinv;
                ZMAT_(ip + 1, K - NP) = -tinv;
                ZMAT_(iq + 1, K - NP) = -tinv;
                ZMAT_(K, K - NP) = tinv;
            }
        }
        int nrem = NPT, kold = 1, knew = KOPT;
        double beta = 0.0, denom = 0.0, stale = 0.0; /* `stale`: the reference's TEMP as :2009 finds it = loop 280's last value */
        for (;;) {
            /* label 80: exchange PTSID(kold) with PTSID(knew) */
            for (int J = 1; J <= N; J++) { const double sw = BMAT_(kold, J); BMAT_(kold, J) = BMAT_(knew, J); BMAT_(knew, J) = sw; }
            for (int J = 1; J <= NPTM; J++) { const double sw = ZMAT_(kold, J); ZMAT_(kold, J) = ZMAT_(knew, J); ZMAT_(knew, J) = sw; }
            ID[kold] = ID[knew];
            ID[knew] = 0.0;
            Wr[NDIM + knew] = 0.0;
            nrem--;
            if (knew != KOPT) {
                const double sw = VL[kold]; VL[kold] = VL[knew]; VL[knew] = sw;
                UPDATE(VLAG, beta, denom, knew, Wr + 1);
                if (nrem == 0) return;
                for (int K = 1; K <= NPT; K++) Wr[NDIM + K] = fabs(Wr[NDIM + K]);
            }
            /* label 120: the nearest original point not yet reinstated (negative marks: tried and refused) */
            bool placed = false;
            while (!placed) {
                double dsqmin = 0.0;
                for (int K = 1; K <= NPT; K++)
                    if (Wr[NDIM + K] > 0.0 && (fabs(dsqmin) <= eps || Wr[NDIM + K] < dsqmin)) { knew = K; dsqmin = Wr[NDIM + K]; }
                if (fabs(dsqmin) <= eps) goto L260;
                for (int J = 1; J <= N; J++) Wr[NPT + J] = XPT_(knew, J);
                for (int K = 1; K <= NPT; K++) {
                    double sum = 0.0;
                    if (K != KOPT) {
                        if (fabs(ID[K]) <= eps) {
                            for (int J = 1; J <= N; J++) sum = sum + Wr[NPT + J] * XPT_(K, J);
                        } else {
                            const int ip = (int)ID[K];
                            if (ip > 0) sum = Wr[NPT + ip] * PA[(ip - 1) * 2];
                            const int iq = (int)((double)NP * ID[K] - (double)(ip * NP));
                            if (iq > 0) sum = sum + Wr[NPT + iq] * PA[(iq - 1) * 2 + (ip == 0 ? 1 : 0)];
                        }
                    }
                    Wr[K] = 0.5 * sum * sum;
                }
                /* VLAG and BETA of reinstating XPT(knew,.) */
                for (int K = 1; K <= NPT; K++) {
                    double sum = 0.0;
                    for (int J = 1; J <= N; J++) sum = sum + BMAT_(K, J) * Wr[NPT + J];
                    VL[K] = sum;
                }
                beta = 0.0;
                for (int J = 1; J <= NPTM; J++) {
                    double sum = 0.0;
                    for (int K = 1; K <= NPT; K++) sum = sum + ZMAT_(K, J) * Wr[K];
                    beta = beta - sum * sum;
                    for (int K = 1; K <= NPT; K++) VL[K] = VL[K] + sum * ZMAT_(K, J);
                }
                double bsum = 0.0, dist = 0.0;
                for (int J = 1; J <= N; J++) {
                    double sum = 0.0;
                    for (int K = 1; K <= NPT; K++) sum = sum + BMAT_(K, J) * Wr[K];
                    const int JP = J + NPT;
                    bsum = bsum + sum * Wr[JP];
                    for (int ip = NPT + 1; ip <= NDIM; ip++) sum = sum + BMAT_(ip, J) * Wr[ip];
                    bsum = bsum + sum * Wr[JP];
                    VL[JP] = sum;
                    dist = dist + XPT_(knew, J) * XPT_(knew, J);
                }
                beta = 0.5 * dist * dist + beta - bsum;
                VL[KOPT] = VL[KOPT] + 1.0;
                /* kold: the provisional point whose removal keeps the denominator of UPDATE largest */
                denom = 0.0;
                double vlmxsq = 0.0;
                for (int K = 1; K <= NPT; K++) {
                    if (fabs(ID[K]) > eps) {
                        double hdiag = 0.0;
                        for (int J = 1; J <= NPTM; J++) hdiag = hdiag + ZMAT_(K, J) * ZMAT_(K, J);
                        const double den = beta * hdiag + VL[K] * VL[K];
                        if (den > denom) { kold = K; denom = den; }
                    }
                    vlmxsq = fmax(vlmxsq, VL[K] * VL[K]);
                }
                if (denom <= 1.0e-2 * vlmxsq) Wr[NDIM + knew] = -Wr[NDIM + knew] - winc; /* refused: try the next one */
                else placed = true;
            }
        }
    L260: /* the surviving provisional points enter XPT; one evaluation each */
        for (int KPT = 1; KPT <= NPT; KPT++) {
            if (fabs(ID[KPT]) <= eps) continue;
            if (NF >= MAXFUN) { NF = -1; return; }
            for (int J = 1; J <= N; J++) { V1(TK, J) = XPT_(KPT, J); XPT_(KPT, J) = 0.0; }
            const int ip = (int)ID[KPT];
            const int iq = (int)((double)NP * ID[KPT] - (double)(ip * NP));
            double xp = 0.0, xq = 0.0;
            if (ip > 0) { xp = PA[(ip - 1) * 2]; XPT_(KPT, ip) = xp; }
            if (iq > 0) { xq = PA[(iq - 1) * 2 + (ip == 0 ? 1 : 0)]; XPT_(KPT, iq) = xq; }
            for (int K = 1; K <= NPT; K++) {
                double tk = 0.0;
                if (ip > 0) tk = tk + xp * XPT_(K, ip);
                if (iq > 0) tk = tk + xq * XPT_(K, iq);
                V1(TK, N + K) = tk;
                stale = tk;
            }
            ctl->i0 = KPT; ctl->i1 = ip; ctl->i2 = iq; ctl->d0 = xp; ctl->d1 = xq;
            par(OP_RESC_NEWPT);
            for (int I = 1; I <= N; I++) {
                V1(XC, I) = fmin(fmax(V1(XL, I), V1(XBASE, I) + XPT_(KPT, I)), V1(XU, I));
                if (fabs(XPT_(KPT, I) - V1(SL, I)) <= eps) V1(XC, I) = V1(XL, I);
                if (fabs(XPT_(KPT, I) - V1(SU, I)) <= eps) V1(XC, I) = V1(XU, I);
            }
            NF = NF + 1;
            const double F = eval_at_XC();
            V1(FVAL, KPT) = F;
            if (F < V1(FVAL, KOPT)) KOPT = KPT;
            for (int K = 1; K <= NPT; K++) {
                double sum = 0.0;
                for (int J = 1; J <= NPTM; J++) sum = sum + ZMAT_(K, J) * ZMAT_(KPT, J);
                V1(TK, K) = sum;
                if (fabs(ID[K]) > eps) { /* :2009 */
                    const int kp = (int)ID[K];
                    if (kp != 0) { const double a1 = PA[(kp - 1) * 2]; const int ihp = (kp * kp + kp) / 2; V1(HQ, ihp) = V1(HQ, ihp) + stale * (a1 * a1); }
                }
            }
            ctl->i0 = KPT;
            par(OP_RESC_MODEL);
            ID[KPT] = 0.0;
        }
    }

    /* ---------------------------------------------------------------- BOBYQB_H :405-1171 (master) */
    __device__ void BOBYQB_H(double RHOBEG, double RHOEND, int MAXFUN, int& NF_out, int& status) {
        const double HALF = 0.5, ONE = 1.0, TEN = 10.0, TENTH = 0.1, TWO = 2.0, ZERO = 0.0;
        double* Wk = W - 1; double* VL = VLAG - 1;
        int NF = 0, KOPT = 0, KBASE, NRESC, NTRITS, NFSAV, KNEW = 0, IH, IP, KSAV;
        double FBEG = 0, XOPTSQ, FSAVE, RHO, DELTA, DIFFA, DIFFB, DIFFC = 0, DSQ = 0, CRVMIN = 0, vquad1 = 0, DNORM = 0, DISTSQ = 0,
               ERRBIG, FRHOSQ, BDTOL, BDTEST, CURV, FRACSQ, SUM, TEMP, SUMZ, SUMW, ADELT = 0, ALPHA = 0, CAUCHY = 0, SUMA, SUMB, BETA = 0,
               BSUM, DX, DENOM = 0, DELSQ, SCADEN, BIGLSQ, HDIAG, DEN, F = 0, FOPT, DIFF, RATIO = 0, DENSAV, t, DIST;
        status = 0;
        model_update = true;
        opt_update = true;
        PRELIM_H(RHOBEG, MAXFUN, NF, KOPT, FBEG);
        XOPTSQ = ZERO;
        for (int I = 1; I <= N; I++) { V1(XOPT, I) = XPT_(KOPT, I); XOPTSQ = XOPTSQ + V1(XOPT, I) * V1(XOPT, I); }
        FSAVE = V1(FVAL, 1);
        if (NF < NPT) goto L720;
        KBASE = 1;
        RHO = RHOBEG;
        DELTA = RHO;
        NRESC = NF;
        NTRITS = 0;
        DIFFA = ZERO;
        DIFFB = ZERO;
        NFSAV = NF;
    L20:
        if (KOPT != KBASE) { /* label 20, :490-516 */
            model_update = true;
            for (int I = 1; I <= N; I++) V1(TK, I) = V1(XOPT, I);
            par(OP_L20_HQV);
            if (NF > NPT) {
                for (int K = 1; K <= NPT; K++) {
                    TEMP = ZERO;
                    for (int J = 1; J <= N; J++) TEMP = TEMP + XPT_(K, J) * V1(XOPT, J);
                    V1(TK, N + K) = TEMP;
                }
                par(OP_L20_PQV);
            }
        }
    L60:
        TRSBOX_H(DELTA, &Wk[1], &Wk[NP], &Wk[NP + N], &Wk[NP + 2 * N], &Wk[NP + 3 * N], DSQ, CRVMIN, KOPT, vquad1);
        DNORM = fmin(DELTA, sqrt(DSQ));
        if (DNORM < HALF * RHO) {
            NTRITS = -1;
            DISTSQ = (TEN * RHO) * (TEN * RHO);
            if (NF <= NFSAV + 2) goto L650;
            ERRBIG = fmax(fmax(DIFFA, DIFFB), DIFFC);
            FRHOSQ = 0.125 * RHO * RHO;
            if (CRVMIN > ZERO && ERRBIG > FRHOSQ * CRVMIN) goto L650;
            BDTOL = ERRBIG / RHO;
            for (int J = 1; J <= N; J++) {
                BDTEST = BDTOL;
                if (fabs(V1(XNEW, J) - V1(SL, J)) <= eps) BDTEST = Wk[J];
                if (fabs(V1(XNEW, J) - V1(SU, J)) <= eps) BDTEST = -Wk[J];
                if (BDTEST < BDTOL) {
                    CURV = V1(HQ, (J + J * J) / 2);
                    BDTEST = BDTEST + HALF * CURV * RHO;
                    if (BDTEST < BDTOL) goto L650;
                }
            }
            goto L680;
        }
        NTRITS = NTRITS + 1;
    L90:
        if (DSQ <= 1.0e-1 * XOPTSQ) {
            model_update = true;
            FRACSQ = 0.25 * XOPTSQ;
            par(OP_SHIFT); /* moment part first: it needs the XPT of before the shift */
            for (int K = 1; K <= NPT; K++) {
                SUM = -HALF * XOPTSQ;
                for (int I = 1; I <= N; I++) SUM = SUM + XPT_(K, I) * V1(XOPT, I);
                Wk[NPT + K] = SUM;
                TEMP = FRACSQ - HALF * SUM;
                for (int I = 1; I <= N; I++) {
                    Wk[I] = BMAT_(K, I);
                    VL[I] = SUM * XPT_(K, I) + TEMP * V1(XOPT, I);
                    IP = NPT + I;
                    for (int J = 1; J <= I; J++) BMAT_(IP, J) = BMAT_(IP, J) + Wk[I] * VL[J] + VL[I] * Wk[J];
                }
            }
            for (int JJ = 1; JJ <= NPTM; JJ++) {
                SUMZ = ZERO;
                SUMW = ZERO;
                for (int K = 1; K <= NPT; K++) {
                    SUMZ = SUMZ + ZMAT_(K, JJ);
                    VL[K] = Wk[NPT + K] * ZMAT_(K, JJ);
                    SUMW = SUMW + VL[K];
                }
                for (int J = 1; J <= N; J++) {
                    SUM = (FRACSQ * SUMZ - HALF * SUMW) * V1(XOPT, J);
                    for (int K = 1; K <= NPT; K++) SUM = SUM + VL[K] * XPT_(K, J);
                    Wk[J] = SUM;
                    for (int K = 1; K <= NPT; K++) BMAT_(K, J) = BMAT_(K, J) + SUM * ZMAT_(K, JJ);
                }
                for (int I = 1; I <= N; I++) {
                    IP = I + NPT;
                    TEMP = Wk[I];
                    for (int J = 1; J <= I; J++) BMAT_(IP, J) = BMAT_(IP, J) + TEMP * Wk[J];
                }
            }
            for (int J = 1; J <= N; J++) {
                for (int K = 1; K <= NPT; K++) XPT_(K, J) = XPT_(K, J) - V1(XOPT, J);
                for (int I = 1; I <= J; I++) BMAT_(NPT + I, J) = BMAT_(NPT + J, I);
            }
            for (int I = 1; I <= N; I++) {
                V1(XBASE, I) = V1(XBASE, I) + V1(XOPT, I);
                V1(XNEW, I) = V1(XNEW, I) - V1(XOPT, I);
                V1(SL, I) = V1(SL, I) - V1(XOPT, I);
                V1(SU, I) = V1(SU, I) - V1(XOPT, I);
                V1(XOPT, I) = ZERO;
            }
            XOPTSQ = ZERO;
        }
        if (NTRITS == 0) goto L210;
        goto L230;
    L190: /* :643-675: BMAT and ZMAT are rebuilt around XOPT */
        NFSAV = NF;
        KBASE = KOPT;
        RESCUE_H(MAXFUN, NF, DELTA, KOPT);
        model_update = true;
        XOPTSQ = ZERO;
        if (KOPT != KBASE) {
            for (int I = 1; I <= N; I++) { V1(XOPT, I) = XPT_(KOPT, I); XOPTSQ = XOPTSQ + V1(XOPT, I) * V1(XOPT, I); }
        }
        if (NF < 0) { NF = MAXFUN; goto L720; }
        NRESC = NF;
        if (NFSAV < NF) { NFSAV = NF; goto L20; }
        if (NTRITS > 0) goto L60;
    L210:
        ALTMOV(KOPT, KNEW, ADELT, ALPHA, CAUCHY, &Wk[1], &Wk[NP], &Wk[NDIM + 1]);
        for (int I = 1; I <= N; I++) V1(D, I) = V1(XNEW, I) - V1(XOPT, I);
    L230:
        for (int K = 1; K <= NPT; K++) {
            SUMA = ZERO;
            SUMB = ZERO;
            SUM = ZERO;
            for (int J = 1; J <= N; J++) {
                SUMA = SUMA + XPT_(K, J) * V1(D, J);
                SUMB = SUMB + XPT_(K, J) * V1(XOPT, J);
                SUM = SUM + BMAT_(K, J) * V1(D, J);
            }
            Wk[K] = SUMA * (HALF * SUMA + SUMB);
            VL[K] = SUM;
            Wk[NPT + K] = SUMA;
        }
        BETA = ZERO;
        for (int JJ = 1; JJ <= NPTM; JJ++) {
            SUM = ZERO;
            for (int K = 1; K <= NPT; K++) SUM = SUM + ZMAT_(K, JJ) * Wk[K];
            BETA = BETA - SUM * SUM;
            for (int K = 1; K <= NPT; K++) VL[K] = VL[K] + SUM * ZMAT_(K, JJ);
        }
        DSQ = ZERO;
        BSUM = ZERO;
        DX = ZERO;
        for (int J = 1; J <= N; J++) {
            DSQ = DSQ + V1(D, J) * V1(D, J);
            SUM = ZERO;
            for (int K = 1; K <= NPT; K++) SUM = SUM + Wk[K] * BMAT_(K, J);
            BSUM = BSUM + SUM * V1(D, J);
            const int JP = NPT + J;
            for (int I = 1; I <= N; I++) SUM = SUM + BMAT_(JP, I) * V1(D, I);
            VL[JP] = SUM;
            BSUM = BSUM + SUM * V1(D, J);
            DX = DX + V1(D, J) * V1(XOPT, J);
        }
        BETA = DX * DX + DSQ * (XOPTSQ + DX + DX + HALF * DSQ) + BETA - BSUM;
        VL[KOPT] = VL[KOPT] + ONE;
        if (NTRITS == 0) {
            DENOM = VL[KNEW] * VL[KNEW] + ALPHA * BETA;
            if (DENOM < CAUCHY && CAUCHY > ZERO) {
                for (int I = 1; I <= N; I++) { V1(XNEW, I) = V1(XALT, I); V1(D, I) = V1(XNEW, I) - V1(XOPT, I); }
                CAUCHY = ZERO;
                goto L230;
            }
            if (DENOM <= HALF * VL[KNEW] * VL[KNEW] || forced(NF)) {
                if (NF > NRESC) goto L190; /* :747 */
                goto L720;
            }
        } else {
            DELSQ = DELTA * DELTA;
            SCADEN = ZERO;
            BIGLSQ = ZERO;
            KNEW = 0;
            for (int K = 1; K <= NPT; K++) {
                if (K == KOPT) continue;
                HDIAG = ZERO;
                for (int JJ = 1; JJ <= NPTM; JJ++) HDIAG = HDIAG + ZMAT_(K, JJ) * ZMAT_(K, JJ);
                DEN = BETA * HDIAG + VL[K] * VL[K];
                DISTSQ = ZERO;
                for (int J = 1; J <= N; J++) DISTSQ = DISTSQ + (XPT_(K, J) - V1(XOPT, J)) * (XPT_(K, J) - V1(XOPT, J));
                TEMP = fmax(ONE, (DISTSQ / DELSQ) * (DISTSQ / DELSQ));
                if (TEMP * DEN > SCADEN) { SCADEN = TEMP * DEN; KNEW = K; DENOM = DEN; }
                BIGLSQ = fmax(BIGLSQ, TEMP * (VL[K] * VL[K])); /* TEMP*VLAG(K)**2 */
            }
            if (SCADEN <= HALF * BIGLSQ || forced(NF)) {
                if (NF > NRESC) goto L190; /* :783 */
                goto L720;
            }
        }
    L360:
        for (int I = 1; I <= N; I++) {
            V1(XC, I) = fmin(fmax(V1(XL, I), V1(XBASE, I) + V1(XNEW, I)), V1(XU, I));
            if (fabs(V1(XNEW, I) - V1(SL, I)) <= eps) V1(XC, I) = V1(XL, I);
            if (fabs(V1(XNEW, I) - V1(SU, I)) <= eps) V1(XC, I) = V1(XU, I);
        }
        if (NF >= MAXFUN) goto L720;
        NF = NF + 1;
        F = eval_at_XC();
        if (F <= fmax(1.0e-12, 1.0e-20 * FBEG)) goto L720;
        if (NTRITS == -1) { FSAVE = F; goto L720; }
        for (int K = 1; K <= NPT; K++) V1(TK, K) = Wk[NPT + K];
        ctl->i0 = KOPT;
        par(OP_VQUAD);
        if (NTRITS <= 0) {
            for (int I = 1; I <= N; I++) V1(HD1, I) = ZERO;
            IH = 0;
            for (int J = 1; J <= N; J++)
                for (int I = 1; I <= J; I++) {
                    IH = IH + 1;
                    if (I < J) V1(HD1, J) = V1(HD1, J) + V1(HQ, IH) * V1(D, I);
                    V1(HD1, I) = V1(HD1, I) + V1(HQ, IH) * V1(D, J);
                }
            vquad1 = ZERO;
            for (int i = 1; i <= N; i++) vquad1 = vquad1 + V1(D, i) * (V1(GOPT, i) + HALF * V1(HD1, i));
        }
        FOPT = V1(FVAL, KOPT);
        DIFF = F - FOPT - vquad1;
        DIFFC = DIFFB;
        DIFFB = DIFFA;
        DIFFA = fabs(DIFF);
        if (DNORM > RHO) NFSAV = NF;
        if (NTRITS > 0) {
            if (vquad1 >= ZERO) goto L720;
            RATIO = (F - FOPT) / vquad1;
            if (RATIO <= TENTH) DELTA = fmin(HALF * DELTA, DNORM);
            else if (RATIO <= 0.7) DELTA = fmax(HALF * DELTA, DNORM);
            else DELTA = fmin(fmax(2.0 * DELTA, 4.0 * DNORM), 1.0e10);
            if (DELTA <= 1.5 * RHO) DELTA = RHO;
            if (F < FOPT) {
                KSAV = KNEW;
                DENSAV = DENOM;
                DELSQ = DELTA * DELTA;
                SCADEN = ZERO;
                BIGLSQ = ZERO;
                KNEW = 0;
                for (int K = 1; K <= NPT; K++) {
                    HDIAG = ZERO;
                    for (int JJ = 1; JJ <= NPTM; JJ++) HDIAG = HDIAG + ZMAT_(K, JJ) * ZMAT_(K, JJ);
                    DEN = BETA * HDIAG + VL[K] * VL[K];
                    DISTSQ = ZERO;
                    for (int J = 1; J <= N; J++) DISTSQ = DISTSQ + (XPT_(K, J) - V1(XNEW, J)) * (XPT_(K, J) - V1(XNEW, J));
                    TEMP = fmax(ONE, (DISTSQ / DELSQ) * (DISTSQ / DELSQ));
                    if (TEMP * DEN > SCADEN) { SCADEN = TEMP * DEN; KNEW = K; DENOM = DEN; }
                    BIGLSQ = fmax(BIGLSQ, TEMP * (VL[K] * VL[K])); /* TEMP*VLAG(K)**2 */
                }
                if (SCADEN <= HALF * BIGLSQ) { KNEW = KSAV; DENOM = DENSAV; }
            }
        }
        UPDATE(VLAG, BETA, DENOM, KNEW, W);
        model_update = true;
        ctl->i0 = KNEW;
        par(OP_UPDATE_MODEL);
        V1(FVAL, KNEW) = F;
        for (int I = 1; I <= N; I++) { XPT_(KNEW, I) = V1(XNEW, I); Wk[I] = BMAT_(KNEW, I); }
        for (int K = 1; K <= NPT; K++) {
            SUMA = ZERO;
            for (int JJ = 1; JJ <= NPTM; JJ++) SUMA = SUMA + ZMAT_(KNEW, JJ) * ZMAT_(K, JJ);
            SUMB = ZERO;
            for (int J = 1; J <= N; J++) SUMB = SUMB + XPT_(K, J) * V1(XOPT, J);
            TEMP = SUMA * SUMB;
            for (int I = 1; I <= N; I++) Wk[I] = Wk[I] + TEMP * XPT_(K, I);
        }
        for (int I = 1; I <= N; I++) V1(TK, I) = Wk[I];
        par(OP_GQV_W);
        if (F < FOPT) {
            opt_update = true;
            model_update = true;
            KOPT = KNEW;
            XOPTSQ = ZERO;
            for (int J = 1; J <= N; J++) {
                V1(XOPT, J) = V1(XNEW, J);
                XOPTSQ = XOPTSQ + V1(XOPT, J) * V1(XOPT, J);
            }
            for (int I = 1; I <= N; I++) V1(TK, I) = V1(D, I);
            par(OP_L20_HQV);
            for (int K = 1; K <= NPT; K++) {
                TEMP = ZERO;
                for (int J = 1; J <= N; J++) TEMP = TEMP + XPT_(K, J) * V1(D, J);
                V1(TK, N + K) = TEMP;
            }
            par(OP_L20_PQV);
        }
        if (NTRITS > 0) {
            for (int k = 1; k <= NPT; k++) {
                t = ZERO;
                for (int j = 1; j <= N; j++) t = t + XPT_(k, j) * V1(XOPT, j);
                V1(TK, k) = t;
            }
            ctl->i0 = KOPT;
            ctl->model_update = 0;
            par(OP_LFN);
            if (ctl->model_update) model_update = true;
        }
        if (NTRITS == 0) goto L60;
        if (F <= FOPT + TENTH * vquad1) goto L60;
        DISTSQ = fmax((TWO * DELTA) * (TWO * DELTA), (TEN * RHO) * (TEN * RHO));
    L650:
        KNEW = 0;
        for (int K = 1; K <= NPT; K++) {
            SUM = ZERO;
            for (int J = 1; J <= N; J++) SUM = SUM + (XPT_(K, J) - V1(XOPT, J)) * (XPT_(K, J) - V1(XOPT, J));
            if (SUM > DISTSQ) { KNEW = K; DISTSQ = SUM; }
        }
        if (KNEW > 0) {
            DIST = sqrt(DISTSQ);
            if (NTRITS == -1) {
                DELTA = fmin(TENTH * DELTA, HALF * DIST);
                if (DELTA <= 1.5 * RHO) DELTA = RHO;
            }
            NTRITS = 0;
            ADELT = fmax(fmin(TENTH * DIST, DELTA), RHO);
            DSQ = ADELT * ADELT;
            goto L90;
        }
        if (NTRITS == -1) goto L680;
        if (RATIO > ZERO) goto L60;
        if (fmax(DELTA, DNORM) > RHO) goto L60;
    L680:
        if (RHO > RHOEND) {
            DELTA = HALF * RHO;
            RATIO = RHO / RHOEND;
            if (RATIO <= 16.0) RHO = RHOEND;
            else if (RATIO <= 250.0) RHO = sqrt(RATIO) * RHOEND;
            else RHO = TENTH * RHO;
            DELTA = fmax(DELTA, RHO);
            NTRITS = 0;
            NFSAV = NF;
            goto L60;
        }
        if (NTRITS == -1) goto L360;
    L720:
        if (V1(FVAL, KOPT) <= FSAVE) {
            for (int I = 1; I <= N; I++) {
                V1(XC, I) = fmin(fmax(V1(XL, I), V1(XBASE, I) + V1(XOPT, I)), V1(XU, I));
                if (fabs(V1(XOPT, I) - V1(SL, I)) <= eps) V1(XC, I) = V1(XL, I);
                if (fabs(V1(XOPT, I) - V1(SU, I)) <= eps) V1(XC, I) = V1(XU, I);
            }
        }
        NF_out = NF;
    }
};

/* BIG: XPT, BMAT and ZMAT live in the restart's slab of global memory (nx > ~55).  A template parameter, not a run-time choice of
 * pointer: with the three matrices carved from `sm` the compiler knows they are shared memory and reads them with LDS; a pointer
 * that may be either makes every access a generic load (measured: the C5 stage 5.8 -> 7.3 s). */
template <class Obj, bool BIG>
__global__ void __launch_bounds__(Obj::kVector ? 512 : 256) dfnls_kernel(const TkTablesG T, const TkScalars sc, const DfArgs a) {
    extern __shared__ double sm[];
    __shared__ DfCtl ctl;
    __shared__ tk_obj_ctx sctx;
    __shared__ double pk_f[16];
    __shared__ int pk_i[5][16];
    const int r = blockIdx.x;
    if (a.async && r == a.n_run) { /* the time keeper of an asynchronous run (see nm_quad_async_kernel) */
        if (threadIdx.x >= 32) return;
        const int P = a.n_run, lane = threadIdx.x;
        int* gvt = a.aq.clk + P + 32;
        for (;;) {
            int m = 0x7fffffff;
            for (int i = lane; i < P; i += 32) m = min(m, tknm::ld_vol(a.aq.clk + i));
            m = (int)__reduce_min_sync(0xffffffffu, (unsigned)m);
            if (lane < TK_GVT_COPIES) tknm::st_vol(gvt + lane * 32, m);
            if (m == 0x7fffffff) return;
            __nanosleep(100);
        }
    }
    if (r >= a.n_run) return;
    const int slot = a.rank + r * a.world;
    Df<Obj> d;
    d.aq = &a.aq; d.a_t0 = 0; d.a_nev = 0; d.a_clk = nullptr;
    d.pk_f = pk_f; d.pk_t = pk_i[0]; d.pk_s = pk_i[1]; d.pk_k = pk_i[2]; d.pk_v = pk_i[3]; d.pk_e = pk_i[4];
    d.tid = threadIdx.x; d.nthr = blockDim.x;
    d.N = sc.nx; d.NPT = a.npt; d.mv = a.mv; d.NDIM = d.NPT + d.N; d.NP = d.N + 1; d.NPTM = d.NPT - d.NP; d.NH = (d.N * d.NP) / 2;
    d.eps = 2.220446049250313e-16;
    d.ctl = &ctl;
    d.ctx = &sctx;
    const int N = d.N, NPT = d.NPT, NDIM = d.NDIM, NH = d.NH;
    /* carve the shared small arrays */
    double* p = sm;
    auto take = [&](int n) { double* q = p; p += n; return q; };
    double* sbl = take(N); double* sbu = take(N); double* saux = take(N);
    d.XL = sbl; d.XU = sbu;
    const int nzm = NPT * (d.NPTM > 0 ? d.NPTM : 1);
    d.XBASE = take(N);
    if constexpr (!BIG) d.XPT = take(NPT * N);
    d.FVAL = take(NPT); d.XOPT = take(N); d.GOPT = take(N); d.HQ = take(NH);
    if constexpr (!BIG) { d.BMAT = take(NDIM * N); d.ZMAT = take(nzm); }
    d.SL = take(N); d.SU = take(N); d.XNEW = take(N);
    d.XALT = take(N); d.D = take(N); d.VLAG = take(NDIM); d.W = take(2 * NDIM + 6 * N + 2 * NPT + 16); d.HD1 = take(N);
    d.TK = take(NDIM + NPT + 8); d.XC = take(N); d.VBUF = take(blockDim.x);
    d.VSH = nullptr; d.YSH = nullptr;
    if constexpr (Obj::kVector) { d.VSH = take(d.mv); d.YSH = take(Obj::smem_doubles(sc.user, 0)); }
    /* per-moment slab */
    double* g = a.slab + (size_t)r * a.slab_doubles;
    const size_t mv = (size_t)d.mv;
    d.GQV = g; g += mv * N; d.HQV = g; g += mv * NH; d.PQV = g; g += mv * NPT; d.FVAL_V = g; g += mv * NPT; d.WV = g; g += mv * N;
    d.v_sumpq = g; g += mv; d.v_err = g; g += mv; d.v_beg = g; g += mv; d.v_temp = g; g += mv; d.v_diff = g; g += mv;
    d.v_vquad = g; g += mv; d.PQVold = g; g += mv; d.v_fbase = g; g += mv; d.v_itest = (int*)g; g += mv;
    if constexpr (BIG) { d.XPT = g; g += (size_t)NPT * N; d.BMAT = g; g += (size_t)NDIM * N; d.ZMAT = g; g += nzm; }
    d.nrescue = 0; d.force_first = a.force_first; d.force_every = a.force_every > 0 ? a.force_every : 1 << 30; d.forced_at = -1;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    if (threadIdx.x == 0) { sctx = tk_make_ctx(sc, N, sbl, sbu, saux); ctl.op = -1; ctl.model_update = 0; }
    __syncthreads();
    if (threadIdx.x != 0) { d.worker_loop(); return; }
    /* ---- master: BOBYQA_H :354-396 for one start; returns objFun at the result (completeSearch :582) ---- */
    auto search = [&](auto start, double RHOBEG, double RHOEND, int& NF, int& status) -> double {
        NF = 0; status = 0;
        d.nrescue = 0; d.forced_at = -1;
        ctl.model_update = 0;
        bool ok = !(NPT < N + 2 || NPT > 2 * N + 1);
        for (int J = 1; J <= N && ok; J++) {
            double xj = start(J - 1);
            const double TEMP = sbu[J - 1] - sbl[J - 1];
            if (TEMP < RHOBEG + RHOBEG) { ok = false; d.XC[J - 1] = xj; break; }
            double sl = sbl[J - 1] - xj, su = sbu[J - 1] - xj;
            if (sl >= -RHOBEG) {
                if (sl >= 0.0) { xj = sbl[J - 1]; sl = 0.0; su = TEMP; }
                else { xj = sbl[J - 1] + RHOBEG; sl = -RHOBEG; su = fmax(sbu[J - 1] - xj, RHOBEG); }
            } else if (su <= RHOBEG) {
                if (su <= 0.0) { xj = sbu[J - 1]; sl = -TEMP; su = 0.0; }
                else { xj = sbu[J - 1] - RHOBEG; sl = fmin(sbl[J - 1] - xj, -RHOBEG); su = RHOBEG; }
            }
            d.XC[J - 1] = xj; d.SL[J - 1] = sl; d.SU[J - 1] = su;
        }
        if (ok) d.BOBYQB_H(RHOBEG, RHOEND, a.maxfun, NF, status);
        else { status = 2; for (int J = 1; J <= N; J++) d.XC[J - 1] = start(J - 1); }
        return d.eval_at_XC();
    };
    /* status (2 bits) and the number of RESCUE_H calls (5 bits, saturating) travel with the evaluation count (pack rows, ingest,
     * local_fetch / local_status / local_nrescue) */
    auto packed_evals = [&](int NF, int status) { return (NF + 1) | ((status | ((d.nrescue < 31 ? d.nrescue : 31) << 2)) << 24); };
    if (!a.async) {
        int NF, status;
        const double fn = search([&](int j) { return a.starts[(size_t)j * a.n_k + slot]; }, a.rhobeg[slot], a.rhoend[slot], NF, status);
        for (int J = 0; J < N; J++) a.out_x[(size_t)J * a.n_k + slot] = d.XC[J];
        a.out_f[slot] = fn;
        a.out_evals[slot] = packed_evals(NF, status);
        a.out_status[slot] = status;
    } else {
        /* asynchronous instances: the scheme of nm_quad_async_kernel with the evaluation (one pass over the panel: the unit
         * every instance spends the same real time on) as the tick of the virtual clock: a restart costs cb + #evaluations */
        const tknm::NmAsync& q = a.aq;
        const int P = a.n_run, cb = q.cb;
        int k = q.first_k + r, Tv = 0;
        d.a_clk = q.clk + r;
        for (;;) {
            ctl.i0 = Tv; ctl.i1 = r; ctl.i2 = P;
            d.par(OP_ASYNC_PICK);
            const int cnt_vis = ctl.i0, bk = ctl.i1;
            double stv[100]; /* the blended start (nx <= 100) */
            if (Tv > 0) k = q.first_k + P + ctl.i3;
            if (k > q.last_k) { tknm::st_vol(q.clk + r, 0x7fffffff); break; }
            if (Tv > 0) atomicMax(q.clk + P, k);
            q.base[k] = (cnt_vis > 1 && bk >= 0) ? bk : -1;
            for (int dd = 0; dd < N; dd++) { /* getModifiedParam :683-720 */
                const double ev = q.xs[(size_t)(dd + 1) * q.K + (k - 1)];
                double out = ev;
                if (cnt_vis > 1 && bk >= 0) {
                    const double wgt = q.w11[k];
                    out = TK_ADD(TK_MUL(wgt, q.res[(size_t)bk * (N + 1) + 1 + dd]), TK_MUL(TK_SUB(1.0, wgt), ev));
                }
                stv[dd] = out;
                q.arch_starts[(size_t)dd * q.K + (k - 1)] = out;
            }
            d.a_t0 = Tv + cb; d.a_nev = 0;
            int NF, status;
            const double fn = search([&](int j) { return stv[j]; }, a.rhobeg[k - q.first_k], a.rhoend[k - q.first_k], NF, status);
            const int F = Tv + cb + d.a_nev;
            double* dst = q.res + (size_t)k * (N + 1);
            const double frt = q.roundtrip ? tk_f4020_roundtrip(fn) : fn;
            dst[0] = frt;
            for (int dd = 0; dd < N; dd++) dst[1 + dd] = q.roundtrip ? tk_f4020_roundtrip(d.XC[dd]) : d.XC[dd];
            q.ev[k].fv = frt;
            q.ev[k].slt = r;
            q.valid[k] = (frt == frt) ? 1 : 0;
            q.arch_evals[k] = packed_evals(NF, status);
            __threadfence();
            tknm::st_vol(&q.ev[k].fin, F);
            __threadfence();
            tknm::st_vol(q.clk + r, F + cb + 1);
            const int* gvt = q.clk + P + 32 + (r % TK_GVT_COPIES) * 32;
            unsigned ns = 100;
            const long long t_wait0 = clock64();
            while (!(tknm::ld_vol(gvt) > F)) {
                if (clock64() - t_wait0 > TK_ASYNC_PATIENCE) { atomicExch(q.err, 1); break; } /* never in a healthy run */
                __nanosleep(ns);
                if (ns < 1600) ns += ns;
            }
            __threadfence();
            if (tknm::ld_vol(q.err)) { tknm::st_vol(q.clk + r, 0x7fffffff); break; }
            Tv = F;
        }
    }
    ctl.op = OP_EXIT;
    d.bar();
}


/* ---------------------------------------------------------------------------------------------
 * Nelder-Mead for objectives whose dfovec fills a VECTOR of moments (objective.f90:76-123; here the plugin form
 * `dfovec_block`: the whole block computes Fnout(1:nm) at one point -- TIKTAK_OBJ_SMM simulates its panel).  objFun =
 * DOT_PRODUCT(F, F) in index order + getPenalty (:125-142), through the same eval_at_XC() as DFNLS.  One restart per block:
 * thread 0 runs runAmoeba (TiktakGlobalSearch.f90:602-633) and amoeba/amotry (minimize.f90:7-135) as written -- first-occurrence
 * MINLOC/MAXLOC, incremental psum (:130), re-sum after a shrink (:112) -- and every objective evaluation is the block's.
 * An evaluation costs a panel pass, so the serial simplex bookkeeping is noise next to it.
 * --------------------------------------------------------------------------------------------- */
template <class Obj>
__global__ void __launch_bounds__(512) nm_vec_kernel(const TkTablesG T, const TkScalars sc, const tknm::NmArgs a) {
    extern __shared__ double sm[];
    __shared__ DfCtl ctl;
    __shared__ tk_obj_ctx sctx;
    const int r = blockIdx.x;
    if (r >= a.n_run) return;
    const int slot = a.rank + r * a.world;
    const int N = sc.nx, NPV = N + 1;
    Df<Obj> d;
    d.tid = threadIdx.x; d.nthr = blockDim.x;
    d.N = N; d.NPT = 0; d.mv = sc.nm; d.NDIM = 0; d.NP = N + 1; d.NPTM = 0; d.NH = 0;
    d.eps = 2.220446049250313e-16;
    d.ctl = &ctl;
    d.ctx = &sctx;
    double* q = sm;
    auto take = [&](int n) { double* t = q; q += n; return t; };
    double* sbl = take(N); double* sbu = take(N); double* saux = take(N);
    d.XL = sbl; d.XU = sbu;
    d.XC = take(N);
    double* P = take(NPV * N); double* Y = take(NPV); double* PSUM = take(N); double* PTRY = take(N);
    d.VSH = take(d.mv); d.YSH = take(Obj::smem_doubles(sc.user, 0));
    for (int i = threadIdx.x; i < N; i += blockDim.x) { sbl[i] = T.bl[i]; sbu[i] = T.bu[i]; saux[i] = T.aux[i]; }
    if (threadIdx.x == 0) { sctx = tk_make_ctx(sc, N, sbl, sbu, saux); ctl.op = -1; ctl.model_update = 0; }
    __syncthreads();
    d.v_err = d.VSH; /* OP_DFOVEC also stores the residuals to v_err: let it store them where they already are */
    if (threadIdx.x != 0) { d.worker_loop(); return; }
    /* ---- master ---- */
    auto objfun = [&](const double* x) -> double {
        for (int j = 0; j < N; j++) d.XC[j] = x[j];
        return d.eval_at_XC();
    };
    const bool shr = a.shr_have && a.shr_have[0] && ((float)(a.first_k + slot) / (float)a.K > 0.5f);
    const double* wsrc = shr ? a.wshr : a.width;
    for (int ia = 0; ia < NPV; ia++) { /* runAmoeba :622-627 */
        for (int i = 0; i < N; i++) {
            double t = TK_ADD(a.starts[(size_t)i * a.n_k + slot], TK_DIV(TK_MUL(a.simplex[(size_t)ia * N + i], wsrc[i]), a.scale));
            P[ia * N + i] = fmax(fmin(t, sbu[i]), sbl[i]);
        }
        Y[ia] = objfun(P + ia * N);
    }
    for (int j = 0; j < N; j++) { double s_ = 0.0; for (int i = 0; i < NPV; i++) s_ = TK_ADD(s_, P[i * N + j]); PSUM[j] = s_; } /* :49 */
    int iter = 0, ilo = 0, ihi = 0, inhi = 0;
    auto amotry = [&](double fac) -> double { /* :116-134 */
        const double fac1 = TK_DIV(TK_SUB(1.0, fac), (double)N), fac2 = TK_SUB(fac1, fac);
        for (int j = 0; j < N; j++) PTRY[j] = TK_SUB(TK_MUL(PSUM[j], fac1), TK_MUL(P[ihi * N + j], fac2));
        const double ytry = objfun(PTRY);
        if (ytry < Y[ihi]) {
            Y[ihi] = ytry;
            for (int j = 0; j < N; j++) { PSUM[j] = TK_ADD(TK_SUB(PSUM[j], P[ihi * N + j]), PTRY[j]); P[ihi * N + j] = PTRY[j]; }
        }
        return ytry;
    };
    for (;;) {
        ilo = 0; for (int i = 1; i < NPV; i++) if (Y[i] < Y[ilo]) ilo = i;        /* iminloc: first occurrence */
        ihi = 0; for (int i = 1; i < NPV; i++) if (Y[i] > Y[ihi]) ihi = i;        /* imaxloc */
        { const double ytmp = Y[ihi]; Y[ihi] = Y[ilo]; inhi = 0; for (int i = 1; i < NPV; i++) if (Y[i] > Y[inhi]) inhi = i; Y[ihi] = ytmp; }
        if (tknm::nm_converged(a, Y[ilo], Y[ihi]) || iter >= a.itmax) break;     /* :84-96 */
        double ytry = amotry(-1.0);
        iter++;
        if (ytry <= Y[ilo]) { ytry = amotry(2.0); iter++; }
        else if (ytry >= Y[inhi]) {
            const double ysave = Y[ihi];
            ytry = amotry(0.5);
            iter++;
            if (ytry >= ysave) { /* :106-113 */
                for (int i = 0; i < NPV; i++)
                    if (i != ilo) { for (int j = 0; j < N; j++) P[i * N + j] = TK_MUL(0.5, TK_ADD(P[i * N + j], P[ilo * N + j])); Y[i] = objfun(P + i * N); }
                iter += N;
                for (int j = 0; j < N; j++) { double s_ = 0.0; for (int i = 0; i < NPV; i++) s_ = TK_ADD(s_, P[i * N + j]); PSUM[j] = s_; }
            }
        }
    }
    /* runAmoeba :631-632 + completeSearch :582: the first minimal vertex and objFun there (same point, same value) */
    for (int j = 0; j < N; j++) a.out_x[(size_t)j * a.n_k + slot] = P[ilo * N + j];
    a.out_f[slot] = Y[ilo];
    a.out_evals[slot] = NPV + iter + 1;
    ctl.op = OP_EXIT;
    d.bar();
}

} /* namespace */

/* Nelder-Mead for vector-valued objectives (one restart per block; see nm_vec_kernel) */
int tk_launch_nm_vec(tiktak_handle* h, const tknm::NmArgs& a) {
    if (h->cfg.objective_id != TIKTAK_OBJ_SMM) { tk_set_error("nm_vec: objective %d has no block-wide dfovec", h->cfg.objective_id); return TIKTAK_ERR_UNSUPPORTED; }
    const int nx = h->cfg.nx, mv = h->cfg.nm;
    const size_t sh = (size_t)(4 * nx + (nx + 1) * nx + (nx + 1) + 2 * nx + mv + 2 * TK_SMM_SMEM_DOUBLES(h->smm_T) + mv) * sizeof(double);
    if (sh > 226 * 1024) { tk_set_error("nm_vec: shared memory (nx=%d, nm=%d)", nx, mv); return TIKTAK_ERR_UNSUPPORTED; }
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    TK_CUDA(cudaFuncSetAttribute(nm_vec_kernel<ObjSmm>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    nm_vec_kernel<ObjSmm><<<a.n_run, 512, sh, h->stream>>>(T, h->sc, a);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

/* instances > 0: asynchronous instances (one cooperative launch for all n_k restarts, rows committed by the kernel; *used = instances run) */
int tk_launch_dfnls(tiktak_handle* h, int first_k, int n_k, int instances, int* used) {
    const int nx = h->cfg.nx, mv = h->cfg.nm, npt = 2 * nx + 1; /* p_ninterppt, genericParams.f90:311 */
    const int ndim = npt + nx, nh = nx * (nx + 1) / 2, nptm = npt - nx - 1;
    DfArgs a;
    a.n_k = n_k; a.rank = h->cfg.rank; a.world = h->cfg.world;
    a.n_run = (n_k > a.rank) ? (n_k - a.rank + a.world - 1) / a.world : 0;
    a.async = 0;
    if (instances > 0) {
        if (h->cfg.world != 1 || h->cfg.search_type != 0) { tk_set_error("local_async (DFNLS): one GPU, search_type 0"); return TIKTAK_ERR_UNSUPPORTED; }
        a.async = 1;
        a.n_run = std::min(instances, n_k); /* clamped to the resident blocks below */
    }
    if (a.n_run <= 0) return TIKTAK_OK;
    if (nx > 100) { tk_set_error("DFNLS supports nx <= 100 (nmax of minimize.f90:416)"); return TIKTAK_ERR_UNSUPPORTED; }
    /* rhobeg / rhoend per restart k (completeSearch :564-571): itratio is a single-precision quotient.
     * Computed once for all K restarts and kept on the device (d_rho[k-1], d_rho[K + k-1]). */
    const int K = h->cfg.maxpoints_plus1;
    if (!h->rho_ready) {
        std::vector<double> rho((size_t)2 * K);
        double mn = h->bound[nx] - h->bound[0];
        for (int i = 1; i < nx; i++) mn = std::min(mn, h->bound[nx + i] - h->bound[i]);
        for (int k = 1; k <= K; k++) {
            const double itratio = (double)((float)k / (float)K);
            if (itratio > 0.5) { rho[k - 1] = (mn / 2.5) / (4 * itratio); rho[K + k - 1] = 1.0e-3 / (4 * itratio); }
            else { rho[k - 1] = mn / 2.50; rho[K + k - 1] = 1.0e-3; }
        }
        if (!h->d_rho) TK_CUDA(cudaMalloc(&h->d_rho, (size_t)2 * K * sizeof(double)));
        TK_CUDA(cudaMemcpyAsync(h->d_rho, rho.data(), rho.size() * 8, cudaMemcpyHostToDevice, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream)); /* rho is stack-local */
        h->rho_ready = true;
    }
    /* shared memory: every small array; the three big ones (XPT, BMAT, ZMAT) move to the slab when they do not fit */
    const int threads = h->cfg.objective_id == TIKTAK_OBJ_SMM ? 512 : (mv >= 256 ? 256 : (mv >= 128 ? 128 : 64));
    const size_t big = (size_t)npt * nx + (size_t)ndim * nx + (size_t)npt * (nptm > 0 ? nptm : 1);
    const size_t small = (size_t)(3 * nx + nx + npt + nx + nx + nh + 5 * nx + ndim + (2 * ndim + 6 * nx + 2 * npt + 16) + nx + (ndim + npt + 8) + nx + threads);
    const size_t extra = h->cfg.objective_id == TIKTAK_OBJ_SMM ? (size_t)(mv + 2 * TK_SMM_SMEM_DOUBLES(h->smm_T)) : 0;
    a.big_in_global = ((small + big + extra) * sizeof(double) > 226 * 1024) ? 1 : 0;
    const size_t slab = (size_t)mv * (nx + nh + npt + npt + nx + 10) + 8 + (a.big_in_global ? big : 0);
    const size_t need = slab * (size_t)a.n_run;
    if (h->dfnls_slab_doubles < need) {
        if (h->d_dfnls_slab) cudaFree(h->d_dfnls_slab);
        h->d_dfnls_slab = nullptr;
        TK_CUDA(cudaMalloc(&h->d_dfnls_slab, need * sizeof(double)));
        h->dfnls_slab_doubles = need;
    }
    if (!h->d_out_status) TK_CUDA(cudaMalloc(&h->d_out_status, (size_t)K * sizeof(int)));
    TK_CUDA(cudaMemsetAsync(h->d_out_status, 0, (size_t)n_k * sizeof(int), h->stream));
    a.starts = h->d_starts;
    a.rhobeg = h->d_rho + (first_k - 1);
    a.rhoend = h->d_rho + K + (first_k - 1);
    a.maxfun = h->cfg.maxeval;
    a.npt = npt; a.mv = mv;
    a.slab = h->d_dfnls_slab; a.slab_doubles = slab;
    a.out_x = h->d_out_x; a.out_f = h->d_out_f; a.out_evals = h->d_out_evals; a.out_status = h->d_out_status;
    { /* test hook: TIKTAK_DFNLS_FORCE_RESCUE=<first NF>[,<every>] */
        a.force_first = 0; a.force_every = 0;
        const char* e = getenv("TIKTAK_DFNLS_FORCE_RESCUE");
        if (e) sscanf(e, "%d,%d", &a.force_first, &a.force_every);
    }
    /* (the SMM objective runs its panel simulation warp-specialised: 256 producer + 256 consumer threads, smm.cuh) */
    const size_t shd = (small + (a.big_in_global ? 0 : big)) * sizeof(double);
    if (shd > 226 * 1024) { tk_set_error("DFNLS small matrices do not fit in shared memory (nx=%d)", nx); return TIKTAK_ERR_UNSUPPORTED; }
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    /* rounds: one block per restart of the launch.  Asynchronous instances: min(instances, resident blocks - 1) blocks + the time
     * keeper, launched cooperatively (the instances wait for one another) */
    auto go = [&](auto kern, size_t sh) -> int {
        TK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        if (!a.async) {
            kern<<<a.n_run, threads, sh, h->stream>>>(T, h->sc, a);
        } else {
            int per_sm = 0;
            TK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, sh));
            a.n_run = std::max(1, std::min(a.n_run, per_sm * h->sm_count - 1));
            int rc = tk_async_setup(h, first_k, n_k, 1, &a.aq);
            if (rc) return rc;
            if ((rc = tk_async_init(h, a.aq, a.n_run))) return rc;
            void* args[] = {(void*)&T, (void*)&h->sc, (void*)&a};
            TK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)a.n_run + 1), dim3((unsigned)threads), args, sh, h->stream));
            if (used) *used = a.n_run;
        }
        h->launches++;
        TK_CUDA(cudaGetLastError());
        return a.async ? tk_async_best(h) : TIKTAK_OK;
    };
    if (h->cfg.objective_id == TIKTAK_OBJ_SMM) {
        const size_t shv = shd + (size_t)(mv + 2 * TK_SMM_SMEM_DOUBLES(h->smm_T)) * sizeof(double);
        if (shv > 226 * 1024) { tk_set_error("DFNLS+SMM shared memory"); return TIKTAK_ERR_UNSUPPORTED; }
        if (a.big_in_global) { tk_set_error("DFNLS+SMM: the small matrices must fit in shared memory"); return TIKTAK_ERR_UNSUPPORTED; }
        return go(dfnls_kernel<ObjSmm, false>, shv);
    }
    return tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        return a.big_in_global ? go(dfnls_kernel<Obj, true>, shd) : go(dfnls_kernel<Obj, false>, shd);
    });
}
