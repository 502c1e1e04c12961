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
        if (live) {
            if (h == 0) {
#pragma unroll 2
                for (int p = 0; p < TK_SMM_CHUNK; p += 4) {
                    tk_smm_accumulate(ya[p], acc[0]); tk_smm_accumulate(ya[p + 1], acc[1]);
                    tk_smm_accumulate(ya[p + 2], acc[2]); tk_smm_accumulate(ya[p + 3], acc[3]);
                }
            } else {
#pragma unroll 2
                for (int p = 0; p < TK_SMM_CHUNK; p += 4) {
                    tk_smm_accumulate(TK_SUB(yb[p], ya[p]), acc[0]); tk_smm_accumulate(TK_SUB(yb[p + 1], ya[p + 1]), acc[1]);
                    tk_smm_accumulate(TK_SUB(yb[p + 2], ya[p + 2]), acc[2]); tk_smm_accumulate(TK_SUB(yb[p + 3], ya[p + 3]), acc[3]);
                }
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

/* plugin wrapper for the vector-valued contract used by the DFNLS kernel */
struct ObjSmm {
    static constexpr int kId = TIKTAK_OBJ_SMM;
    static constexpr bool kVector = true;
    static void host_init(int nx, double* aux) { for (int i = 0; i < nx; i++) aux[i] = 0.0; }
    static __device__ __forceinline__ int smem_doubles(const void* user, int nm) { return TK_SMM_SMEM_DOUBLES(((const SmmDev*)user)->T) + nm; }
    static __device__ __forceinline__ void dfovec_block(const void* user, const double* theta, double* v, double* ysh) {
        tk_smm_dfovec_block(*(const SmmDev*)user, theta, v, ysh);
    }
};

int tk_smm_init(tiktak_handle* h);
void tk_smm_free(tiktak_handle* h);
int tk_launch_smm_wave(tiktak_handle* h, const TkWave& wv);
int tk_launch_smm_objfun(tiktak_handle* h, const double* d_x, int64_t n, double* d_f);
