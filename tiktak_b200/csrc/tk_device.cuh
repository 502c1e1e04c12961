/* tk_device.cuh -- the device-side declarations the kernels share, free of host-only headers: also pasted into the source
 * of the run-time compiled Sobol kernel (rtc.cu). */
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>

#include "../../include/tiktak_cuda.h"
#include "../../include/tiktak_obj.cuh"
#endif

#define TK_MAXBIT 30          /* utilities.f90:54: atmost < 2^30 */
#define TK_NX_GENERIC_MAX 128 /* largest nx of the runtime-nx fallback kernels */
#define TK_SOBOL_S 7          /* log2 of the Sobol kernel's block size (128 threads) */


/* tables that ride along as __grid_constant__ kernel parameters (constant bank, broadcast reads) */
template <int NX>
struct TkTables {
    double lo[NX], w[NX], bl[NX], bu[NX], aux[NX];
    const uint32_t* tab; /* device memory: Vt[TK_MAXBIT][NXP] bit-major direction numbers, Dt[TK_MAXBIT][NXP] step table, Lt[2^S][NXP] thread part of the seed */
    int nbits;           /* Gray bits in use (rows of the tables a block copies) */
};
#define TK_SOBOL_NXP(nx) (((nx) + 3) & ~3) /* row length of the tables, padded to 16 bytes */
struct TkScalars {
    int nx, nm;
    double fvalmax, penalty0, penaltyc, max_penalty, sqrt_nm;
    const void* user;
};

/* pointers to the same tables in global memory, for the runtime-nx kernels */
struct TkTablesG {
    const uint32_t* V; /* [nx][32] */
    const double *lo, *w, *bl, *bu, *aux;
};

/* one evaluation wave of the pre-test stage: the points n (1-based Sobol index; n = 0 is the
 * skipped all-zero vector) of count blocks [cb0, cb0+ncb) that belong to this rank */
struct TkWave {
    int64_t N;            /* qr_ndraw */
    int64_t cb0, ncb;     /* global count-block range of the wave */
    int rank, world;
    double* fval;         /* local store, indexed by local n (see tk_local_index) */
    unsigned long long* counts; /* per GLOBAL count block */
    int64_t n_lo = 1, n_hi = -1; /* optional sub-range of indices (expensive objectives: finer early-stop grain); n_hi < 0: all */
    int64_t cbf = 0, cbf_q = 0;  /* filled by the launcher: first count block of the wave owned by this rank, and cbf / world */
};


#ifdef __CUDACC__
/* accessors handed to Obj::value */
template <int NX>
struct TkRegX {
    double v[NX];
    __device__ __forceinline__ double operator[](int i) const { return v[i]; }
};
struct TkMemX { /* strided memory (shared or global) */
    const double* p;
    int64_t stride;
    __device__ __forceinline__ double operator[](int i) const { return p[(int64_t)i * stride]; }
};
__device__ __forceinline__ tk_obj_ctx tk_make_ctx(const TkScalars& s, int nx, const double* bl, const double* bu,
                                                  const double* aux) {
    tk_obj_ctx c;
    c.nx = nx; c.nm = s.nm; c.bl = bl; c.bu = bu; c.aux = aux;
    c.fvalmax = s.fvalmax; c.penalty0 = s.penalty0; c.penaltyc = s.penaltyc; c.max_penalty = s.max_penalty;
    c.sqrt_nm = s.sqrt_nm; c.user = s.user;
    c.trig = tk_trig_plain(); /* kernels with a hot loop over tk_cos swap in register-resident copies (kernels_sobol.cu) */
    return c;
}
/* 32-bit Sobol numerator -> double in [0,1), exact: build 1.q in [1,2) and subtract 1 */
__device__ __forceinline__ double tk_q2unit(uint32_t q) {
    return __dsub_rn(__hiloint2double((int)(0x3ff00000u | (q >> 12)), (int)(q << 20)), 1.0);
}
#endif
