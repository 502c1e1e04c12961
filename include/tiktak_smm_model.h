/* tiktak_smm_model.h -- the arithmetic of the synthetic SMM objective (TIKTAK_OBJ_SMM), written once
 * for host and device so that the CUDA kernels and any host-side user of the plugin agree bit for bit.
 * (The CPU oracle restates the same model independently in oracle/smm_oracle.inc.)
 * All operations are plain IEEE +,-,*,/ , comparisons and fabs -- no transcendental functions on the
 * evaluation path; the common random numbers are drawn once on the host (Philox4x32-10 + Box-Muller).
 */
#ifndef TIKTAK_SMM_MODEL_H
#define TIKTAK_SMM_MODEL_H
#include "tiktak_math.h"

/* read-only global loads on the device (the shock panel is never written after obj_initialize) */
#if defined(__CUDA_ARCH__)
#define TK_LDG(p) __ldg(p)
#else
#define TK_LDG(p) (*(p))
#endif

#define TK_SMM_NSERIES 6
#define TK_SMM_NSTAT 5
#define TK_SMM_TMAX 40

/* horizon of series h: 0 = level y_t, else y_{t+k} - y_t */
TK_HD int tk_smm_horizon(int h) { return h == 0 ? 0 : (h == 1 ? 1 : (h == 2 ? 2 : (h == 3 ? 3 : (h == 4 ? 5 : 10)))); }

/* one person's path: y[t * ys], t = 0..T-1.  Shocks are read with stride `st` (SoA [t][person]). */
TK_HD void tk_smm_simulate(const double* theta, int T, double alpha0, const double* eta0, const double* eps0, const double* u,
                           long st, double* y, int ys) {
    const double rho = theta[0], sa = theta[1], se = theta[2], s1 = theta[3], s2 = theta[4], p = theta[5], mu1 = theta[6];
    const double mu2 = TK_DIV(-TK_MUL(p, mu1), TK_SUB(1.0, p)); /* zero-mean mixture */
    const double a = TK_MUL(sa, alpha0);
    double z = 0.0;
    /* ages in groups of 8: the 24 shock loads of a group are independent of the z recursion and are issued
     * together (memory-level parallelism on the device); the arithmetic order is unchanged */
    for (int t0 = 0; t0 < T; t0 += 8) {
        double e0[8], uu[8], ep[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; j++) {
            const int t = (t0 + j < T) ? t0 + j : T - 1;
            e0[j] = TK_LDG(eta0 + (long)t * st);
            uu[j] = TK_LDG(u + (long)t * st);
            ep[j] = TK_LDG(eps0 + (long)t * st);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; j++) {
            if (t0 + j < T) {
                /* select the mixture component first, then one multiply-add pair: same operands, same roundings */
                const bool first = uu[j] < p;
                const double eta = TK_ADD(first ? mu1 : mu2, TK_MUL(first ? s1 : s2, e0[j]));
                z = TK_ADD(TK_MUL(rho, z), eta);
                y[(long)(t0 + j) * ys] = TK_ADD(TK_ADD(a, z), TK_MUL(se, ep[j]));
            }
        }
    }
}

/* accumulate the 5 statistics of one observation s.  The power sums use explicit fused multiply-adds
 * (one rounding each) -- part of the DEFINITION of this synthetic objective; the oracle uses fma() too. */
TK_HD void tk_smm_accumulate(double s, double* acc) {
    const double s2 = TK_MUL(s, s);
    acc[0] = TK_ADD(acc[0], s);
    acc[1] = TK_FMA(s, s, acc[1]);
    acc[2] = TK_FMA(s2, s, acc[2]);
    acc[3] = TK_FMA(s2, s2, acc[3]);
    acc[4] = TK_ADD(acc[4], fabs(s));
}
#endif
