/* nm_common.cuh -- what the Nelder-Mead kernels of kernels_nm.cu (latency form: four warps per restart) and
 * kernels_nm_w4.cu (one warp per restart, device-side restart queue) share. */
#pragma once
#include "tk_common.cuh"

namespace tknm {

struct NmArgs {
    int n_run;            /* restarts this launch runs */
    int n_k;              /* restarts of the launch (column length of starts/out_x) */
    int rank, world;      /* restart slot = rank + r*world */
    int first_k, K;       /* restart number of slot 0; p_maxpoints */
    const double* starts; /* (n_k, nx) column-major */
    const double* width;  /* nx: p_range(:,2)-p_range(:,1) */
    const double* wshr;   /* nx: simplex width from the best minima so far (nm_shrink_mode = 1), else NULL */
    const int* shr_have;  /* device flag: wshr is defined (>= 2 finished restarts), else NULL */
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

/* Device-side restart queue of one launch of nm_w4_kernel, and (world == 1) the in-kernel form of
 * completeSearch's bookkeeping (TiktakGlobalSearch.f90:582-597): the group that finishes the last restart of a
 * round commits the rounds that are complete, in order, into the device-resident searchResults table.  The
 * launch may hold rounds that were started early against the best row known when it was queued; the first
 * committed round that changes that row (or switches the blend on, :701) sets `abort`: groups stop taking
 * tickets, running restarts of later rounds stop at their next check, `accepted` tells the host where to
 * queue the rest again -- the accepted sequence is that of strictly sequential rounds. */
struct NmQueue {
    int* ctl;         /* [0] head ticket, [1] next round to commit, [2] lock, [3] gate: tickets below depth0 + gate may start, negative =
                         the launch stops, [4] accepted restarts, [5] rounds committed in a row without a change of the best row,
                         [8 + r] restarts of round r done */
    int commit;       /* 1: commit in the kernel (world == 1) */
    int round_size;   /* restarts per round */
    int n_rounds;
    int depth0;       /* rounds a ticket may run ahead of the commit pointer at the start of the launch */
    int roundtrip;    /* text_roundtrip */
    int nx;
    double* res;      /* (K+1) x (nx+1) result rows */
    int* valid;
    int* arch_evals;
    double* best;     /* [0] fval [1] k [2] #rows [3..] x */
    double* pslab;    /* simplices of the resident warps when they live in global memory (nm_w4_kernel, nx >= 32): (nx+1)*nx doubles per warp */
    const struct NmAsync* aq = nullptr; /* not NULL: asynchronous instances, one warp = one instance, ticks = objective evaluations (device copy of the
                                 descriptor; the last block of the grid keeps the time) */
    int aq_P = 0;     /* instances (warps) of the asynchronous run */
};
enum { NMQ_HEAD = 0, NMQ_COMMIT = 1, NMQ_LOCK = 2, NMQ_GATE = 3, NMQ_ACCEPTED = 4, NMQ_STREAK = 5, NMQ_DONE0 = 8 };

/* (ya, ra) sorts before (yb, rb): y ascending, vertex index descending among equal y */
__device__ __forceinline__ bool nm_before(double ya, int ra, double yb, int rb) { return ya < yb || (ya == yb && ra > rb); }

/* amoeba's stopping test rtol < ftol (minimize.f90:84-85); the division is only formed when the product test is within 2^-40 */
__device__ __forceinline__ bool nm_converged(const NmArgs& a, double ylo, double yhi) {
    const double num = TK_MUL(2.0, fabs(TK_SUB(yhi, ylo)));
    const double den = TK_ADD(TK_ADD(fabs(yhi), fabs(ylo)), a.tiny);
    const double thr = TK_MUL(a.ftol, den);
    bool conv = num < TK_MUL(thr, 0.99999999999909051);
    if (!conv && !(num > TK_MUL(thr, 1.00000000000090949))) conv = TK_DIV(num, den) < a.ftol;
    return conv;
}

/* ---- asynchronous instances (kernels_nm.cu: nm_quad_async_kernel; kernels_dfnls.cu: dfnls_kernel with DfArgs.aq) ---- */
struct NmAsync {
    int first_k, last_k;   /* restarts of the launch */
    int K, nx, roundtrip;
    const double* xs;      /* x_starts (K, nx+1) column-major */
    const double* w11;     /* blend weight per restart number */
    double* res;           /* (K+1) x (nx+1) result rows [fval, x] */
    int* valid;
    int* arch_evals;
    double* arch_starts;   /* (K, nx) column-major */
    TkAsyncRow* ev;        /* [K+1] value, virtual finish time (-1: not finished; 0: rows known before the launch) and instance (-1: known
                              before) of row k, one 16-byte record per row for the scans */
    int* clk;              /* [P] lower bound of the instance's next finish time; clk[P]: highest restart number handed out */
    int* err;              /* set when an instance gave up waiting (TK_ASYNC_PATIENCE cycles without the global time passing its own):
                              the launch ends instead of hanging the GPU, tiktak_cuda_local_fetch reports it */
    long long* dbg;        /* experiment: per instance [wait cycles, scan cycles, run cycles, restarts, polls] or NULL */
    int free_run;          /* 1: no virtual clock -- an instance takes the next restart number from a counter and sees whatever rows are
                              committed at that moment, like the reference's processes (not reproducible run to run); fin = commit order */
    int* base;             /* [K+1] the row restart k was blended toward (-1: none): lets a free-running run be verified restart by restart */
    int c0, cb;            /* ticks of the initial simplex / of a shrink step; ticks an instance needs between two restarts */
};

#define TK_GVT_COPIES 16
#define TK_ASYNC_PATIENCE (1ll << 34) /* SM cycles (~9 s) an instance waits for the others before it gives up */
__device__ __forceinline__ int ld_vol(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_vol(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
/* (f, fin, slt, k) of row A before row B: lower value first, then the order of the (virtual) file */
__device__ __forceinline__ bool nm_row_before(double fa, int ta, int sa, int ka, double fb, int tb, int sb, int kb) {
    if (fa != fb) return fa < fb;
    if (ta != tb) return ta < tb;
    if (sa != sb) return sa < sb;
    return ka < kb;
}


} /* namespace tknm */

int tk_async_setup(tiktak_handle* h, int first_k, int n_k, int c0, tknm::NmAsync* q);
int tk_async_init(tiktak_handle* h, const tknm::NmAsync& q, int P);
int tk_async_best(tiktak_handle* h);
