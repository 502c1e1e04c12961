/* sobol_kernel.cuh -- the fixed-nx Sobol evaluation kernel, as a header: compiled into the library for the (objective, nx) pairs
 * BASELINE.json names (kernels_sobol.cu) and at run time, through NVRTC, for any other pair (rtc.cu).  Device code only. */
#pragma once
#ifndef __CUDACC_RTC__
#include "tk_device.cuh"
#endif

#ifndef TK_SOBOL_MB
#define TK_SOBOL_MB 6 /* resident blocks per SM the grouped kernels are compiled for (register budget 65536 / (MB * 128)) */
#endif
#ifndef TK_SOBOL_G
#define TK_SOBOL_G 5 /* terms per group of the breadth-first evaluation */
#endif

#ifdef TK_SOBOL_TRACE /* experiment build: cycle stamps of block 0 / thread 0 (tools/sobol_trace.py) */
__device__ unsigned long long tk_trace_buf[32];
__device__ __forceinline__ unsigned long long tk_gtime() { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); return g; }
#define TK_TRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { tk_trace_buf[2 * (i)] = clock64(); tk_trace_buf[2 * (i) + 1] = tk_gtime(); } } while (0)
__device__ unsigned long long tk_blk_trace[3 * 8192]; /* per block: globaltimer at entry, at exit, SM id */
#define TK_BLK_IN() do { if (threadIdx.x == 0 && blockIdx.x < 8192) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); tk_blk_trace[3 * blockIdx.x] = tk_gtime(); tk_blk_trace[3 * blockIdx.x + 2] = sm; } } while (0)
#define TK_BLK_OUT() do { if (threadIdx.x == 0 && blockIdx.x < 8192) tk_blk_trace[3 * blockIdx.x + 1] = tk_gtime(); } while (0)
#else
#define TK_TRACE(i) do { } while (0)
#define TK_BLK_IN() do { } while (0)
#define TK_BLK_OUT() do { } while (0)
#endif


/* INB: the host has verified that the whole Sobol box lies inside the bounds (lo >= bl and fl(lo + w) <= bu; every
 * point is <= fl(lo + w) because q < 1 and rounding is monotonic), so getPenalty is identically 0 and the per-point
 * bounds test (2 DSETP per coordinate on the FP64 pipe) is compiled out -- same values. */
template <class Obj, int NX, int BT, int P, bool INB, int G>
__global__ void __launch_bounds__(BT, G > 0 ? TK_SOBOL_MB : 0)
sobol_eval_kernel(const __grid_constant__ TkTables<NX> T, const TkScalars sc, const TkWave wv) {
    constexpr int S = TK_SOBOL_S;
    static_assert((1 << S) == BT, "the step table is built for blocks of 2^TK_SOBOL_S threads");
    constexpr int CH = BT * P;
    constexpr int CPB = TIKTAK_SOBOL_BLOCK / CH;
    static_assert(CPB * CH == TIKTAK_SOBOL_BLOCK, "chunk must divide the count block");
    constexpr int NXP = TK_SOBOL_NXP(NX);
    /* Direction numbers, bit-major: sV[j][d] = V[d][j]; sD[c][d] = what the step n -> n + BT does to the state of
     * coordinate d when it flips Gray bits S-1 and S+c.  They are copied from global memory by the whole block
     * (independent coalesced loads: one memory latency) instead of being read as constant-bank operands: a
     * constant operand that misses stalls the warp, the misses of a fresh SM come one after the other, and the
     * table spans 2 NX lines -- 5 us of a 34 us launch at C2, 14 us at C4 (tools/sobol_fixed_cost.py). */
    __shared__ __align__(16) uint32_t sV[TK_MAXBIT * NXP];
    __shared__ __align__(16) uint32_t sD[TK_MAXBIT * NXP];
    __shared__ uint32_t qhi[NX];
    __shared__ unsigned int blk_cnt;
    __shared__ double strig[5][32]; /* one copy per lane of the register constants: tk_cos's four and 2^-32 (see below) */
    __shared__ __align__(16) double ssel[12]; /* tk_sel_tab: per-lane coefficient reads become LDS.128 */

    TK_TRACE(0);
    TK_BLK_IN();
    /* count block of this CUDA block: the wave's first owned count block (wv.cbf, found by the launcher: no
     * 64-bit division in the kernel) + world * (blocks of CPB chunks in front of it) */
    const int lcb = (int)(blockIdx.x / CPB);
    const int64_t cb = wv.cbf + (int64_t)lcb * wv.world;
    if (cb >= wv.cb0 + wv.ncb) return;
    const int coff = (int)(blockIdx.x % CPB) * CH; /* offset of the chunk inside its count block */
    const int64_t base = cb * TIKTAK_SOBOL_BLOCK + coff;
    if (base > wv.N) return;
    const int t = threadIdx.x;
    const uint32_t h0 = (uint32_t)(base >> S);
    const int nrow = T.nbits; /* Gray bits in use: rows 0 .. nbits-1 */
    TK_TRACE(1);
    static_assert(P >= 2, "the thread part of the seed assumes chunks of at least 2 BT points");
    uint4 low[NXP / 4]; /* row t of the thread-part table: issued first, needed last */
    {
        const uint4* gL = reinterpret_cast<const uint4*>(T.tab + 2 * TK_MAXBIT * NXP) + t * (NXP / 4);
#pragma unroll
        for (int g = 0; g < NXP / 4; g++) low[g] = __ldg(gL + g);
    }
    {
        const uint4* gV = reinterpret_cast<const uint4*>(T.tab);
        const uint4* gD = reinterpret_cast<const uint4*>(T.tab + TK_MAXBIT * NXP);
        uint4* s4V = reinterpret_cast<uint4*>(sV);
        uint4* s4D = reinterpret_cast<uint4*>(sD);
        for (int i = t; i < nrow * (NXP / 4); i += BT) { s4V[i] = __ldg(gV + i); s4D[i] = __ldg(gD + i); }
    }
    if (t == 0) blk_cnt = 0;
    if (t >= 32 && t < 44) ssel[t - 32] = tk_sel_tab[t - 32];
    if (t < 32) { const tk_trig_regs k = tk_trig_plain(); strig[0][t] = k.twoopi; strig[1][t] = k.s6; strig[2][t] = k.c6; strig[3][t] = k.mhalf; strig[4][t] = 0x1p-32; }
    __syncthreads();
    TK_TRACE(2);

    /* block-uniform part of the seed: Gray bits >= S come from h0 only */
    const uint32_t ghi = h0 ^ (h0 >> 1);
    for (int d = t; d < NX; d += BT) {
        uint32_t q = 0;
        for (int j = 0; j + S < nrow; j++)
            if ((ghi >> j) & 1u) q ^= sV[(j + S) * NXP + d];
        qhi[d] = q;
    }
    __syncthreads();
    TK_TRACE(3);
    /* thread part: Gray bits < S of n0 = base + t.  base is a multiple of 2 BT (P >= 2), so they are the Gray code of
     * t alone, the same in every block: row t of a table the host built once, read at the top of the kernel */
    uint32_t q[NXP]; /* state of coordinate d: the 32-bit numerator of the point (padding entries stay unused) */
#pragma unroll
    for (int g = 0; g < NXP / 4; g++) {
        q[4 * g] = low[g].x; q[4 * g + 1] = low[g].y; q[4 * g + 2] = low[g].z; q[4 * g + 3] = low[g].w;
    }
#pragma unroll
    for (int d = 0; d < NX; d++) q[d] ^= qhi[d];

    TK_TRACE(4);
    tk_obj_ctx ctx = tk_make_ctx(sc, NX, T.bl, T.bu, T.aux);
    { /* the register-resident constants of tk_cos come from a per-lane shared-memory copy: a value loaded from a
       * lane-dependent address is neither re-materialised inside the loop nor moved to the uniform registers
       * (which a DFMA that also carries an immediate cannot read) */
        const int l = t & 31;
        ctx.trig.twoopi = strig[0][l]; ctx.trig.s6 = strig[1][l]; ctx.trig.c6 = strig[2][l]; ctx.trig.mhalf = strig[3][l];
        ctx.trig.sel = reinterpret_cast<const double2*>(ssel);
    }
    const double k2m32 = strig[4][t & 31]; /* 2^-32 in a register: the FMA below also reads lo[d] from the constant bank */
    /* Loop-carried bookkeeping: the index of the current point (32 bits: qr_ndraw < 2^30), where its value goes and
     * the block-uniform step counter.  The empty asm makes them opaque, so ptxas keeps them in three registers instead
     * of re-deriving them from blockIdx and the kernel arguments in every iteration (45 instructions per point). */
    uint32_t n32 = (uint32_t)base + (uint32_t)t;
    double* out = wv.fval + ((wv.cbf_q + lcb) * TIKTAK_SOBOL_BLOCK + coff) + t; /* fval[tk_local_index(base, world) + t] */
    uint32_t hh = h0;
    const uint32_t N32 = (uint32_t)wv.N;
    asm volatile("" : "+r"(n32), "+l"(out), "+r"(hh));
    unsigned int cnt = 0;
#pragma unroll 1
    for (int j = 0; j < P; j++) {
        TkRegX<NX> x;
#pragma unroll
        for (int d = 0; d < NX; d++) {
            /* lo + u*w with u = q * 2^-32 (:834, unfused): the power of two commutes with the rounding of the
             * product, fl(u*w) = fl(q*w) * 2^-32 exactly, so an integer conversion, a product and an FMA whose
             * product is exact give the reference's two roundings with one FP64 instruction less */
            x.v[d] = __fma_rn(__dmul_rn((double)q[d], T.w[d]), k2m32, T.lo[d]);
        }
        double f;
        if constexpr (INB && G > 0) f = tk_objfun_finish(ctx, tk_value_grouped<Obj, Obj::nterms(NX), G, true>(ctx, x), 0.0);
        else f = INB ? tk_objfun_finish(ctx, tk_value_seq<Obj>(ctx, x), 0.0) : tk_objfun<Obj>(ctx, x);
        if (n32 - 1u < N32) { /* 1 <= n <= qr_ndraw */
            asm volatile("st.global.f64 [%0], %1;" ::"l"(out), "d"(f) : "memory"); /* `out` is opaque to the compiler: name the state space */
            cnt += (f < sc.fvalmax) ? 1u : 0u;
        }
        n32 += BT;
        out += BT;
        TK_TRACE(5 + (j < 3 ? j : 3));
        /* advance to n + BT: Gray bits S-1 and S+c flip (one row of sD, broadcast reads) */
        const int c = __ffs((int)~hh) - 1;
        hh++;
        const uint4* row = reinterpret_cast<const uint4*>(sD + c * NXP);
#pragma unroll
        for (int g = 0; g < NXP / 4; g++) {
            const uint4 w = row[g];
            q[4 * g] ^= w.x; q[4 * g + 1] ^= w.y; q[4 * g + 2] ^= w.z; q[4 * g + 3] ^= w.w;
        }
    }
    /* valid count of this chunk -> its count block */
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((t & 31) == 0 && cnt) atomicAdd(&blk_cnt, cnt);
    __syncthreads();
    if (t == 0 && blk_cnt) atomicAdd(&wv.counts[cb], (unsigned long long)blk_cnt);
    TK_TRACE(9);
    TK_BLK_OUT();
}

