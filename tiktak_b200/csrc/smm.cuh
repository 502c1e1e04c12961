/* smm.cuh -- device side of TIKTAK_OBJ_SMM: the moment vector of one parameter vector, computed by one
 * CUDA block.  This is the dfovec of an objective.f90 that simulates a panel per evaluation (the use case
 * the reference's README.md:41-43 benchmarks): np persons x T ages are streamed from HBM once per
 * evaluation (3 shock arrays, SoA [t][person], + alpha0 = 8 (3 np T + np) bytes), simulated 256 persons
 * at a time into shared memory, and accumulated into the 30 T moment cells by (t, series) owner threads
 * in increasing person order -- the same order a serial CPU loop uses, so sums are bit-identical.
 */
#pragma once
#include "../../include/tiktak_smm_model.h"
#include "tk_common.cuh"

struct SmmDev {
    int np, T;
    const double *eta0, *eps0, *u, *alpha0; /* [t][person] x3, [person] */
    const double *mdata, *scale;            /* [30 T] */
};

#define TK_SMM_CHUNK 256
#define TK_SMM_ROW (TK_SMM_CHUNK + 1) /* padded row: owners of different ages hit different banks */
#define TK_SMM_SMEM_DOUBLES(T) (TK_SMM_ROW * (T))

/* non-aligned named barrier: callable from the master/worker structure of the DFNLS kernel */
__device__ __forceinline__ void tk_smm_bar() { asm volatile("barrier.sync 2, %0;" ::"r"((int)blockDim.x) : "memory"); }

/* v_err[30 T] (any address space) for theta[7]; all threads of the block must call; blockDim >= 256.
 * ysh: shared memory of TK_SMM_CHUNK * T doubles. */
__device__ inline void tk_smm_dfovec_block(const SmmDev& S, const double* theta, double* v_err, double* ysh) {
    const int tid = threadIdx.x, T = S.T;
    const int npair = T * TK_SMM_NSERIES; /* <= 240 <= blockDim.x: thread q owns pair q = t*6 + h */
    /* Each cell is summed as 4 interleaved partial sums over the persons (person index mod 4), combined
     * ((a0+a1)+(a2+a3)) at the end: part of the objective's definition (the oracle does the same); it gives
     * every owner thread 20 independent FP64 chains instead of 5. */
    double acc[4][TK_SMM_NSTAT];
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int s = 0; s < TK_SMM_NSTAT; s++) acc[g][s] = 0.0;
    double th[7];
#pragma unroll
    for (int i = 0; i < 7; i++) th[i] = theta[i];
    const bool owner = tid < npair;
    const int t = owner ? tid / TK_SMM_NSERIES : 0, h = owner ? tid - t * TK_SMM_NSERIES : 0, k = tk_smm_horizon(h);
    const bool live = owner && (t + k < T);
    const double* ya = ysh + t * TK_SMM_ROW;
    const double* yb = ysh + (live ? t + k : t) * TK_SMM_ROW;
    for (int c0 = 0; c0 < S.np; c0 += TK_SMM_CHUNK) {
        tk_smm_bar();
        if (tid < TK_SMM_CHUNK) {
            const long i = c0 + tid;
            tk_smm_simulate(th, T, S.alpha0[i], S.eta0 + i, S.eps0 + i, S.u + i, (long)S.np, ysh + tid, TK_SMM_ROW); /* [t][p] */
        }
        tk_smm_bar();
        if (live) { /* one loop for levels (h == 0: s = y_t) and differences (s = y_{t+k} - y_t): no divergence in the warp */
            const bool lvl = (h == 0);
#pragma unroll 2
            for (int p = 0; p < TK_SMM_CHUNK; p += 4) {
                const double a0 = ya[p], a1 = ya[p + 1], a2 = ya[p + 2], a3 = ya[p + 3];
                const double b0 = yb[p], b1 = yb[p + 1], b2 = yb[p + 2], b3 = yb[p + 3];
                tk_smm_accumulate(lvl ? a0 : TK_SUB(b0, a0), acc[0]); tk_smm_accumulate(lvl ? a1 : TK_SUB(b1, a1), acc[1]);
                tk_smm_accumulate(lvl ? a2 : TK_SUB(b2, a2), acc[2]); tk_smm_accumulate(lvl ? a3 : TK_SUB(b3, a3), acc[3]);
            }
        }
    }
    const double rnp = (double)S.np;
    if (owner)
#pragma unroll
        for (int s = 0; s < TK_SMM_NSTAT; s++) {
            const int cell = tid * TK_SMM_NSTAT + s;
            const double tot = TK_ADD(TK_ADD(acc[0][s], acc[1][s]), TK_ADD(acc[2][s], acc[3][s]));
            v_err[cell] = TK_DIV(TK_SUB(TK_DIV(tot, rnp), S.mdata[cell]), S.scale[cell]);
        }
    tk_smm_bar();
}

/* The same moment vector by a block of 512 threads, warp-specialised: threads 256..511 (producers) stream the
 * shocks from HBM and simulate chunk c+1 into one half of a double-buffered shared panel while threads 0..239
 * (the (t, series) owners) accumulate chunk c from the other half.  Hand-over through four named barriers
 * (FULL/EMPTY per buffer; bar.arrive on the releasing side, bar.sync on the waiting side), so the HBM latency of
 * the simulate phase hides behind the FP64 work of the accumulate phase.  Same arithmetic, same order: results are
 * bit-identical to tk_smm_dfovec_block.  ysh: 2 * TK_SMM_SMEM_DOUBLES(T) doubles. */
#define TK_SMM_BAR_FULL 8  /* + buffer index */
#define TK_SMM_BAR_EMPTY 10 /* + buffer index */
__device__ __forceinline__ void tk_bar_sync(int id, int n) { asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tk_bar_arrive(int id, int n) { asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ inline void tk_smm_dfovec_block_ws(const SmmDev& S, const double* theta, double* v_err, double* ysh) {
    const int tid = threadIdx.x, T = S.T;
    const int npair = T * TK_SMM_NSERIES;
    const int nchunks = S.np / TK_SMM_CHUNK;
    const int bufsz = TK_SMM_SMEM_DOUBLES(T);
    tk_smm_bar(); /* the previous evaluation's readers are done with both buffers */
    if (tid >= 256 && tid < 512) { /* ---- producers ---- */
        double th[7];
#pragma unroll
        for (int i = 0; i < 7; i++) th[i] = theta[i];
        const int ptid = tid - 256;
        for (int c = 0; c < nchunks; c++) {
            const int b = c & 1;
            if (c >= 2) tk_bar_sync(TK_SMM_BAR_EMPTY + b, 512);
            const long i = (long)c * TK_SMM_CHUNK + ptid;
            tk_smm_simulate(th, T, S.alpha0[i], S.eta0 + i, S.eps0 + i, S.u + i, (long)S.np, ysh + b * bufsz + ptid, TK_SMM_ROW);
            tk_bar_arrive(TK_SMM_BAR_FULL + b, 512);
        }
    } else if (tid < 256) { /* ---- consumers: thread q < npair owns pair q = t*6 + h ---- */
        double acc[4][TK_SMM_NSTAT];
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int s = 0; s < TK_SMM_NSTAT; s++) acc[g][s] = 0.0;
        const bool owner = tid < npair;
        const int t = owner ? tid / TK_SMM_NSERIES : 0, h = owner ? tid - t * TK_SMM_NSERIES : 0, k = tk_smm_horizon(h);
        const bool live = owner && (t + k < T);
        for (int c = 0; c < nchunks; c++) {
            const int b = c & 1;
            const double* ya = ysh + b * bufsz + t * TK_SMM_ROW;
            const double* yb = ysh + b * bufsz + (live ? t + k : t) * TK_SMM_ROW;
            tk_bar_sync(TK_SMM_BAR_FULL + b, 512);
            if (live) { /* one loop for levels and differences: no divergence inside the warp */
                const bool lvl = (h == 0);
#pragma unroll 2
                for (int p = 0; p < TK_SMM_CHUNK; p += 4) {
                    const double a0 = ya[p], a1 = ya[p + 1], a2 = ya[p + 2], a3 = ya[p + 3];
                    const double b0 = yb[p], b1 = yb[p + 1], b2 = yb[p + 2], b3 = yb[p + 3];
                    tk_smm_accumulate(lvl ? a0 : TK_SUB(b0, a0), acc[0]); tk_smm_accumulate(lvl ? a1 : TK_SUB(b1, a1), acc[1]);
                    tk_smm_accumulate(lvl ? a2 : TK_SUB(b2, a2), acc[2]); tk_smm_accumulate(lvl ? a3 : TK_SUB(b3, a3), acc[3]);
                }
            }
            if (c + 2 < nchunks) tk_bar_arrive(TK_SMM_BAR_EMPTY + b, 512);
        }
        const double rnp = (double)S.np;
        if (owner)
#pragma unroll
            for (int s = 0; s < TK_SMM_NSTAT; s++) {
                const int cell = tid * TK_SMM_NSTAT + s;
                const double tot = TK_ADD(TK_ADD(acc[0][s], acc[1][s]), TK_ADD(acc[2][s], acc[3][s]));
                v_err[cell] = TK_DIV(TK_SUB(TK_DIV(tot, rnp), S.mdata[cell]), S.scale[cell]);
            }
    }
    tk_smm_bar();
}

/* plugin wrapper for the vector-valued contract used by the DFNLS kernel: blocks of >= 512 threads run the
 * warp-specialised form */
struct ObjSmm {
    static constexpr int kId = TIKTAK_OBJ_SMM;
    static constexpr bool kVector = true;
    static void host_init(int nx, double* aux) { for (int i = 0; i < nx; i++) aux[i] = 0.0; }
    static __device__ __forceinline__ int smem_doubles(const void* user, int nm) {
        return TK_SMM_SMEM_DOUBLES(((const SmmDev*)user)->T) * (blockDim.x >= 512 ? 2 : 1) + nm;
    }
    static __device__ __forceinline__ void dfovec_block(const void* user, const double* theta, double* v, double* ysh) {
        if (blockDim.x >= 512) tk_smm_dfovec_block_ws(*(const SmmDev*)user, theta, v, ysh);
        else tk_smm_dfovec_block(*(const SmmDev*)user, theta, v, ysh);
    }
};

int tk_smm_init(tiktak_handle* h);
void tk_smm_free(tiktak_handle* h);
int tk_launch_smm_wave(tiktak_handle* h, const TkWave& wv);
int tk_launch_smm_objfun(tiktak_handle* h, const double* d_x, int64_t n, double* d_f);
