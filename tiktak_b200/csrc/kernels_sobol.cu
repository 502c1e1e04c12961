/* kernels_sobol.cu -- the Sobol pre-test stage on the device.
 *
 * Replaces setupSobol (TiktakGlobalSearch.f90:825-838: insobl + N sequential I4_SOBOL calls + a text
 * dump of all points) and the evaluation loop of solveAtSobolPoints (:429-464: one objFun per point
 * bracketed by ~6 lock-file cycles).  Here no point is ever stored: a thread derives its first point
 * from the Gray code of its index (closed form of the Antonov-Saleev recurrence utilities.f90:1086-1106:
 * point n = XOR of the direction numbers selected by the bits of n ^ (n>>1)), then strides through the
 * sequence with ONE XOR per coordinate per point, maps it to the search box with the two roundings of
 * the reference's unfused lo + q*(hi-lo) (:834) and evaluates the objective plugin in registers.
 * Output: 8 bytes per point (fval) + one valid-count per TIKTAK_SOBOL_BLOCK-point count block.
 *
 * Work decomposition: a CUDA block of BT = 2^S threads owns BT*P consecutive Sobol indices; thread t
 * handles n = base + j*BT + t, j = 0..P-1, so the fval stores of a warp are contiguous and the step
 * n -> n+BT flips exactly two Gray-code bits, S-1 and S+ctz(~(n>>S)) -- the same two for the whole
 * block: one row of a precomputed step table, which every block copies to shared memory together with
 * the bit-major direction numbers (fixed-nx kernel; DESIGN.md section 3 tells why the tables do not
 * travel as kernel arguments any more).  The runtime-nx kernel further down is the plain fallback.
 */
#include <stdio.h>
#include <stdlib.h>

#include "tk_common.cuh"

#include "sobol_kernel.cuh"

#ifdef TK_SOBOL_TRACE
extern "C" int tiktak_cuda_debug_trace(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, tk_trace_buf, sizeof(tk_trace_buf)); }
extern "C" int tiktak_cuda_debug_blocks(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, tk_blk_trace, sizeof(tk_blk_trace)); }
#endif

namespace {

template <int NX>
__device__ __forceinline__ uint32_t tk_gray_point_coord(const uint32_t* Vd, uint32_t g) {
    uint32_t q = 0;
    for (int j = 0; j < TK_MAXBIT; j++)
        if ((g >> j) & 1u) q ^= Vd[j];
    return q;
}

/* runtime-nx fallback of the same kernel (nx <= TK_NX_GENERIC_MAX): tables in global memory,
 * state in local memory.  Same arithmetic, same results; only slower. */
template <class Obj, int BT, int P>
__global__ void __launch_bounds__(BT)
sobol_eval_generic_kernel(const TkTablesG T, const TkScalars sc, const TkWave wv) {
    constexpr int S = (BT == 64) ? 6 : (BT == 128) ? 7 : 8;
    constexpr int CH = BT * P;
    constexpr int CPB = TIKTAK_SOBOL_BLOCK / CH;
    __shared__ unsigned int blk_cnt;
    const int nx = sc.nx;
    const int64_t lcb = blockIdx.x / CPB;
    const int64_t cbf = wv.cb0 + (((int64_t)wv.rank - wv.cb0 % wv.world) + wv.world) % wv.world;
    const int64_t cb = cbf + lcb * wv.world;
    if (cb >= wv.cb0 + wv.ncb) return;
    const int64_t base = cb * TIKTAK_SOBOL_BLOCK + (int64_t)(blockIdx.x % CPB) * CH;
    if (base > wv.N) return;
    const int t = threadIdx.x;
    if (t == 0) blk_cnt = 0;
    __syncthreads();
    const uint32_t h0 = (uint32_t)(base >> S);
    const uint32_t n0 = (uint32_t)base + (uint32_t)t;
    const uint32_t g0 = n0 ^ (n0 >> 1);
    uint32_t q[TK_NX_GENERIC_MAX];
    double xv[TK_NX_GENERIC_MAX];
    for (int d = 0; d < nx; d++) q[d] = tk_gray_point_coord<0>(T.V + (size_t)d * 32, g0);
    const tk_obj_ctx ctx = tk_make_ctx(sc, nx, T.bl, T.bu, T.aux);
    const int64_t lbase = tk_local_index(base, wv.world);
    unsigned int cnt = 0;
    for (int j = 0; j < P; j++) {
        const int64_t n = base + (int64_t)j * BT + t;
        for (int d = 0; d < nx; d++) xv[d] = __dadd_rn(T.lo[d], __dmul_rn(tk_q2unit(q[d]), T.w[d]));
        const TkMemX x{xv, 1};
        const double f = tk_objfun<Obj>(ctx, x);
        if (n >= 1 && n <= wv.N) {
            wv.fval[lbase + (int64_t)j * BT + t] = f;
            cnt += (f < sc.fvalmax) ? 1u : 0u;
        }
        const uint32_t h = h0 + (uint32_t)j;
        const int c = __ffs((int)~h) - 1;
        for (int d = 0; d < nx; d++) q[d] ^= T.V[(size_t)d * 32 + S - 1] ^ T.V[(size_t)d * 32 + S + c];
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((t & 31) == 0 && cnt) atomicAdd(&blk_cnt, cnt);
    __syncthreads();
    if (t == 0 && blk_cnt) atomicAdd(&wv.counts[cb], (unsigned long long)blk_cnt);
}

/* the points themselves, (count, nx) column-major like sobol_trial (:830-835) */
__global__ void sobol_points_kernel(const TkTablesG T, int nx, int64_t first, int64_t count, double* out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const uint32_t n = (uint32_t)(first + r);
    const uint32_t g = n ^ (n >> 1);
    for (int d = 0; d < nx; d++) {
        const uint32_t q = tk_gray_point_coord<0>(T.V + (size_t)d * 32, g);
        out[(size_t)d * count + r] = __dadd_rn(T.lo[d], __dmul_rn(tk_q2unit(q), T.w[d]));
    }
}

/* objFun at arbitrary points, x (n, nx) column-major */
template <class Obj>
__global__ void objfun_kernel(const TkTablesG T, const TkScalars sc, const double* x, int64_t n, double* f) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const tk_obj_ctx ctx = tk_make_ctx(sc, sc.nx, T.bl, T.bu, T.aux);
    const TkMemX xa{x + r, n};
    f[r] = tk_objfun<Obj>(ctx, xa);
}

/* exact position of the `need`-th valid point inside count block cb (the legitimate cut,
 * TiktakGlobalSearch.f90:429).  One block; result[0] = n of that point (0 if not reached). */
__global__ void find_cut_kernel(const double* fval, int64_t lbase, int64_t nbase, int64_t N, double fvalmax,
                                unsigned long long need, unsigned long long* result) {
    __shared__ unsigned int wsum[32];
    __shared__ unsigned long long running;
    const int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) { running = 0; result[0] = 0; }
    __syncthreads();
    for (int off = 0; off < TIKTAK_SOBOL_BLOCK; off += nt) {
        const int64_t n = nbase + off + t;
        const bool v = (n >= 1 && n <= N) && (fval[lbase + off + t] < fvalmax);
        const unsigned int b = __ballot_sync(0xffffffffu, v);
        const int lane = t & 31, w = t >> 5;
        if (lane == 0) wsum[w] = __popc(b);
        __syncthreads();
        unsigned long long before = running;
        for (int i = 0; i < w; i++) before += wsum[i];
        const unsigned long long mine = before + __popc(b & ((1u << lane) - 1u)) + (v ? 1 : 0);
        if (v && mine == need) result[0] = (unsigned long long)n;
        __syncthreads();
        if (t == 0) { unsigned long long s = 0; for (int i = 0; i < nt / 32; i++) s += wsum[i]; running += s; }
        __syncthreads();
        if (running >= need) break;
    }
}

/* bit-major direction numbers + step table + thread part of the seed (sobol_kernel.cuh), once per handle */
int build_sobtab(tiktak_handle* h) {
    if (h->d_sobtab) return TIKTAK_OK;
    const int NX = h->cfg.nx, NXP = TK_SOBOL_NXP(NX);
    std::vector<uint32_t> tab((size_t)(2 * TK_MAXBIT + (1 << TK_SOBOL_S)) * NXP, 0u);
    for (int d = 0; d < NX; d++) {
        const uint32_t* Vd = h->V.data() + (size_t)d * 32;
        for (int j = 0; j < TK_MAXBIT; j++) tab[(size_t)j * NXP + d] = Vd[j];
        for (int c = 0; c + TK_SOBOL_S < TK_MAXBIT; c++) {
            tab[(size_t)(TK_MAXBIT + c) * NXP + d] = Vd[TK_SOBOL_S - 1] ^ Vd[TK_SOBOL_S + c];
        }
    }
    for (int t = 0; t < (1 << TK_SOBOL_S); t++) { /* thread part of the seed: XOR of the direction numbers picked by gray(t) */
        const uint32_t g = (uint32_t)t ^ ((uint32_t)t >> 1);
        for (int d = 0; d < NX; d++) {
            uint32_t x = 0;
            for (int j = 0; j < TK_SOBOL_S; j++)
                if ((g >> j) & 1u) x ^= h->V[(size_t)d * 32 + j];
            tab[(size_t)(2 * TK_MAXBIT + t) * NXP + d] = x;
        }
    }
    TK_CUDA(cudaMalloc(&h->d_sobtab, tab.size() * sizeof(uint32_t)));
    TK_CUDA(cudaMemcpyAsync(h->d_sobtab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream)); /* tab is a local */
    return TIKTAK_OK;
}
int sobol_nbits(const tiktak_handle* h) { /* rows in use: bit length of the largest index a block may touch (qr_ndraw rounded up to a chunk) */
    const uint64_t top = (uint64_t)h->cfg.qr_ndraw + TIKTAK_SOBOL_BLOCK;
    int nb = 0;
    while ((top >> nb) != 0) nb++;
    return std::min(nb + 1, TK_MAXBIT);
}

template <class Obj, int NX>
int launch_wave_fixed(tiktak_handle* h, const TkWave& wv, int64_t owned_cbs) {
    TkTables<NX> T;
    memset(&T, 0, sizeof T);
    for (int d = 0; d < NX; d++) {
        T.lo[d] = h->range[d];
        T.w[d] = h->range[NX + d] - h->range[d];
        T.bl[d] = h->bound[d];
        T.bu[d] = h->bound[NX + d];
        T.aux[d] = h->aux[d];
    }
    { const int rcb = build_sobtab(h); if (rcb) return rcb; }
    T.tab = h->d_sobtab;
    T.nbits = sobol_nbits(h);
    /* points per thread P: the smallest P whose grid is at most TWO resident waves (blocks of the second wave start as
     * blocks of the first one retire, which evens out the SMs; measured at C2: P = 4 and 8 within 1 %, 16 is 15 % slower,
     * 2 pays the seeding twice as often); a job too big for that takes the longest stride. */
    bool inb = true; /* Sobol box inside the bounds (and inside the plugin's declared domain)? */
    for (int d = 0; d < NX; d++) inb = inb && (T.lo[d] >= T.bl[d]) && (T.lo[d] + T.w[d] <= T.bu[d]);
    if constexpr (requires(const double* p) { Obj::domain_ok(NX, p, p, p); }) {
        double hi[NX];
        for (int d = 0; d < NX; d++) hi[d] = T.lo[d] + T.w[d];
        inb = inb && Obj::domain_ok(NX, T.lo, hi, T.aux);
    }
    TkWave wvk = wv; /* with the owned-block origin filled in */
    wvk.cbf = wv.cb0 + (((int64_t)wv.rank - wv.cb0 % wv.world) + wv.world) % wv.world;
    wvk.cbf_q = wvk.cbf / wv.world;
    auto try_launch = [&]<int BT, int P>(bool force) -> int {
        static_assert(BT == (1 << TK_SOBOL_S), "the step table is built for blocks of 2^TK_SOBOL_S threads");
        const int64_t grid = owned_cbs * (TIKTAK_SOBOL_BLOCK / (BT * P));
        auto go = [&]<bool INB, int G>() -> int {
            static int per_sm = 0; /* per instantiation: the occupancy query is host work in front of every launch */
            if (per_sm == 0) TK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sobol_eval_kernel<Obj, NX, BT, P, INB, G>, BT, 0));
            if (!force && grid > 2 * (int64_t)per_sm * h->sm_count) return 1;
            sobol_eval_kernel<Obj, NX, BT, P, INB, G><<<(unsigned)grid, BT, 0, h->stream>>>(T, h->sc, wvk);
            return TIKTAK_OK;
        };
        /* terms per group of the breadth-first evaluation (0: one term after the other) */
        constexpr bool kHasTerms = requires(const tk_obj_ctx& c, const TkRegX<NX>& x, double (&a)[1], double (&b)[1]) { Obj::template terms<1, true>(c, x, 0, a, b); };
        constexpr int NT = Obj::nterms(NX);
        constexpr int GA = !kHasTerms ? 0 : (NT % TK_SOBOL_G == 0 ? TK_SOBOL_G : (NT % 5 == 0 ? 5 : (NT % 4 == 0 ? 4 : 0)));
        const int rc = !inb ? go.template operator()<false, 0>() : go.template operator()<true, GA>();
        if (rc != TIKTAK_OK) return rc;
        h->launches++;
        TK_CUDA(cudaGetLastError());
        return TIKTAK_OK;
    };
    /* experiment hook: TIKTAK_SOBOL_SHAPE=<BT>x<P> forces one shape */
    static const char* shape = getenv("TIKTAK_SOBOL_SHAPE");
    if (shape) {
        int bt = 0, pp = 0;
        if (sscanf(shape, "%dx%d", &bt, &pp) == 2) {
            if (bt == 128 && pp == 4) return try_launch.template operator()<128, 4>(true);
            if (bt == 128 && pp == 8) return try_launch.template operator()<128, 8>(true);
            if (bt == 128 && pp == 16) return try_launch.template operator()<128, 16>(true);
        }
    }
    int rc = try_launch.template operator()<128, 4>(false);
    if (rc == 1) rc = try_launch.template operator()<128, 8>(false);
    if (rc == 1) rc = try_launch.template operator()<128, 16>(false);
    if (rc == 1) rc = try_launch.template operator()<128, 32>(true);
    return rc;
}

/* any other (objective, nx <= TK_NX_PARAM_MAX): the same kernel compiled at run time (rtc.cu).  Returns 1 when that is not available. */
template <class Obj>
int launch_wave_rtc(tiktak_handle* h, const TkWave& wv, int64_t owned_cbs) {
    const int NX = h->cfg.nx;
    if (NX > TK_NX_PARAM_MAX) return 1;
    { const int rcb = build_sobtab(h); if (rcb) return rcb; }
    /* TkTables<NX> as the kernel sees it: five arrays of NX doubles, the table pointer, nbits (+ padding to 8 bytes) */
    std::vector<double> blob((size_t)5 * NX + 2, 0.0);
    double *lo = blob.data(), *w = lo + NX, *bl = w + NX, *bu = bl + NX, *aux = bu + NX;
    for (int d = 0; d < NX; d++) {
        lo[d] = h->range[d]; w[d] = h->range[NX + d] - h->range[d];
        bl[d] = h->bound[d]; bu[d] = h->bound[NX + d]; aux[d] = h->aux[d];
    }
    const uint32_t* tabp = h->d_sobtab;
    memcpy(&blob[(size_t)5 * NX], &tabp, sizeof tabp);
    const int nbits = sobol_nbits(h);
    memcpy(&blob[(size_t)5 * NX + 1], &nbits, sizeof nbits);
    bool inb = true;
    for (int d = 0; d < NX; d++) inb = inb && (lo[d] >= bl[d]) && (lo[d] + w[d] <= bu[d]);
    if constexpr (requires(const double* p) { Obj::domain_ok(1, p, p, p); }) {
        std::vector<double> hi(NX);
        for (int d = 0; d < NX; d++) hi[d] = lo[d] + w[d];
        inb = inb && Obj::domain_ok(NX, lo, hi.data(), aux);
    }
    constexpr bool kHasTerms = requires(const tk_obj_ctx& c, const TkRegX<4>& x, double (&a)[1], double (&b)[1]) { Obj::template terms<1, true>(c, x, 0, a, b); };
    const int NT = Obj::nterms(NX);
    const int G = (!inb || !kHasTerms) ? 0 : (NT % TK_SOBOL_G == 0 ? TK_SOBOL_G : (NT % 5 == 0 ? 5 : (NT % 4 == 0 ? 4 : 0)));
    TkWave wvk = wv;
    wvk.cbf = wv.cb0 + (((int64_t)wv.rank - wv.cb0 % wv.world) + wv.world) % wv.world;
    wvk.cbf_q = wvk.cbf / wv.world;
    /* points per thread by the rule of the compiled-in kernels, with the register budget's resident blocks per SM */
    int P = 32;
    for (int p : {4, 8, 16}) {
        const int64_t grid = owned_cbs * (TIKTAK_SOBOL_BLOCK / (128 * p));
        if (grid <= 2 * (int64_t)TK_SOBOL_MB * h->sm_count) { P = p; break; }
    }
    int per_sm = 0;
    void* fn = tk_rtc_sobol_kernel(h, P, inb ? 1 : 0, G, &per_sm);
    if (!fn) return 1;
    const int64_t grid = owned_cbs * (TIKTAK_SOBOL_BLOCK / (128 * P));
    void* args[] = {(void*)blob.data(), (void*)&h->sc, (void*)&wvk};
    const int rc = tk_rtc_launch(h, fn, (unsigned)grid, args);
    if (rc) return rc;
    h->launches++;
    return TIKTAK_OK;
}

TkTablesG tablesG(tiktak_handle* h) { return TkTablesG{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux}; }

} /* namespace */

int tk_launch_sobol_wave(tiktak_handle* h, const TkWave& wv) {
    /* count blocks of [cb0, cb0+ncb) owned by this rank */
    const int64_t cbf = wv.cb0 + (((int64_t)wv.rank - wv.cb0 % wv.world) + wv.world) % wv.world;
    if (cbf >= wv.cb0 + wv.ncb) return TIKTAK_OK;
    const int64_t owned = (wv.cb0 + wv.ncb - 1 - cbf) / wv.world + 1;
    const int nx = h->cfg.nx;
    return tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        /* compile-time-nx instantiations for the configurations BASELINE.json names */
        if (nx == 4 && Obj::kId == TIKTAK_OBJ_TEMPLATE) return launch_wave_fixed<ObjTemplate, 4>(h, wv, owned);
        if (nx == 10 && Obj::kId == TIKTAK_OBJ_GRIEWANK) return launch_wave_fixed<ObjGriewank, 10>(h, wv, owned);
        if (nx == 20 && Obj::kId == TIKTAK_OBJ_RASTRIGIN) return launch_wave_fixed<ObjRastrigin, 20>(h, wv, owned);
        if (nx == 50 && Obj::kId == TIKTAK_OBJ_ROSENBROCK) return launch_wave_fixed<ObjRosenbrock, 50>(h, wv, owned);
        if (nx > TK_NX_GENERIC_MAX) {
            tk_set_error("nx=%d exceeds the generic kernel limit %d", nx, TK_NX_GENERIC_MAX);
            return TIKTAK_ERR_UNSUPPORTED;
        }
        { /* the fixed-nx kernel for this pair, compiled at run time; 1 = not available here: the runtime-nx kernel below */
            const int rcr = launch_wave_rtc<Obj>(h, wv, owned);
            if (rcr != 1) return rcr;
        }
        constexpr int BT = 128, P = 4;
        const unsigned grid = (unsigned)(owned * (TIKTAK_SOBOL_BLOCK / (BT * P)));
        sobol_eval_generic_kernel<Obj, BT, P><<<grid, BT, 0, h->stream>>>(tablesG(h), h->sc, wv);
        h->launches++;
        TK_CUDA(cudaGetLastError());
        return TIKTAK_OK;
    });
}

int tk_launch_sobol_points(tiktak_handle* h, int64_t first, int64_t count, double* d_out) {
    if (count <= 0) return TIKTAK_OK;
    const unsigned grid = (unsigned)((count + 255) / 256);
    sobol_points_kernel<<<grid, 256, 0, h->stream>>>(tablesG(h), h->cfg.nx, first, count, d_out);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_launch_objfun(tiktak_handle* h, const double* d_x, int64_t n, double* d_f) {
    if (n <= 0) return TIKTAK_OK;
    if (h->cfg.nx > TK_NX_GENERIC_MAX) return TIKTAK_ERR_UNSUPPORTED;
    return tk_dispatch_obj(h->cfg.objective_id, [&]<class Obj>() -> int {
        const unsigned grid = (unsigned)((n + 127) / 128);
        objfun_kernel<Obj><<<grid, 128, 0, h->stream>>>(tablesG(h), h->sc, d_x, n, d_f);
        h->launches++;
        TK_CUDA(cudaGetLastError());
        return TIKTAK_OK;
    });
}

int tk_launch_find_cut(tiktak_handle* h, int64_t cb, unsigned long long need, unsigned long long* d_result) {
    const int64_t nbase = cb * TIKTAK_SOBOL_BLOCK;
    const int64_t lbase = tk_local_index(nbase, h->cfg.world);
    find_cut_kernel<<<1, 1024, 0, h->stream>>>(h->d_fval, lbase, nbase, h->cfg.qr_ndraw, h->cfg.fvalmax, need, d_result);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

/* ---------------------------------------------------------------------------------------------
 * FP64 peak microbenchmark: 8 independent DFMA chains per thread, no memory traffic.
 * --------------------------------------------------------------------------------------------- */
namespace {
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123456.789) out[0] = s; /* never true: keeps the chains alive */
}
} /* namespace */

int tk_fp64_peak(int device, double* flops, double* ms_out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { tk_set_error("no CUDA device"); return TIKTAK_ERR_NOGPU; }
    TK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TK_CUDA(cudaGetDeviceProperties(&prop, device));
    double* d_out = nullptr;
    TK_CUDA(cudaMalloc(&d_out, 8));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    TK_CUDA(cudaEventCreate(&e0));
    TK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        TK_CUDA(cudaEventRecord(e0));
        dfma_chain_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-7);
        TK_CUDA(cudaEventRecord(e1));
        TK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        TK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double fl = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    if (flops) *flops = fl / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    return TIKTAK_OK;
}
