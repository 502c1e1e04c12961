/* kernels_smm.cu -- TIKTAK_OBJ_SMM: host-side obj_initialize (common random numbers, data moments) and
 * the Sobol-stage / objFun kernels (one block per parameter vector).  See smm.cuh and
 * include/tiktak_cuda.h (tiktak_smm_params).
 */
#include <math.h>

#include <algorithm>

#include "smm.cuh"

namespace {

/* Philox4x32-10 (Salmon et al. 2011), counter = (e_lo, e_hi, stream, 0), key = (seed, 0) */
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
inline void draw(uint32_t seed, uint32_t stream, uint64_t e, double& unif, double& normal) {
    uint32_t c[4] = {(uint32_t)e, (uint32_t)(e >> 32), stream, 0u};
    philox4x32_10(c, seed, 0u);
    const double u1 = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6) + 0.5) / 9007199254740992.0; /* (0,1) */
    const double u2 = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6) + 0.5) / 9007199254740992.0;
    unif = u1;
    normal = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

struct SmmState {
    SmmDev dev;
    SmmDev* d_dev = nullptr;
    double *d_eta0 = nullptr, *d_eps0 = nullptr, *d_u = nullptr, *d_alpha0 = nullptr, *d_mdata = nullptr, *d_scale = nullptr;
};

int fill_and_upload(SmmState& st, int np, int T, uint32_t seed, cudaStream_t s) {
    const size_t n = (size_t)np * T;
    std::vector<double> eta(n), eps(n), uu(n), al(np);
    double un, no;
    for (int t = 0; t < T; t++)
        for (int i = 0; i < np; i++) {
            const uint64_t e = (uint64_t)i * TK_SMM_TMAX + t; /* element id independent of the layout */
            const size_t o = (size_t)t * np + i;
            draw(seed, 1u, e, un, no); eta[o] = no;
            draw(seed, 2u, e, un, no); eps[o] = no;
            draw(seed, 3u, e, un, no); uu[o] = un;
        }
    for (int i = 0; i < np; i++) { draw(seed, 4u, (uint64_t)i, un, no); al[i] = no; }
    TK_CUDA(cudaMemcpyAsync(st.d_eta0, eta.data(), n * 8, cudaMemcpyHostToDevice, s));
    TK_CUDA(cudaMemcpyAsync(st.d_eps0, eps.data(), n * 8, cudaMemcpyHostToDevice, s));
    TK_CUDA(cudaMemcpyAsync(st.d_u, uu.data(), n * 8, cudaMemcpyHostToDevice, s));
    TK_CUDA(cudaMemcpyAsync(st.d_alpha0, al.data(), (size_t)np * 8, cudaMemcpyHostToDevice, s));
    TK_CUDA(cudaStreamSynchronize(s));
    return TIKTAK_OK;
}

/* raw simulated moments (mdata = 0, scale = 1 while this runs) */
__global__ void __launch_bounds__(256) smm_moments_kernel(const SmmDev* S, const double* theta, double* out) {
    extern __shared__ double ysh[];
    tk_smm_dfovec_block(*S, theta, out, ysh);
}

/* objFun = sum F^2 (index order) + getPenalty, one block per point; x = column-major (n, 7) or Sobol index */
__device__ double smm_objfun_block(const SmmDev& S, const tk_obj_ctx& ctx, const double* theta, double* ysh, double* fsh) {
    __shared__ double s_pen, s_val;
    if (threadIdx.x == 0) s_pen = tk_get_penalty(ctx, TkMemX{theta, 1});
    __syncthreads();
    const double pen = s_pen;
    const int nm = ctx.nm;
    if (pen > TK_MYEPS) { /* dfovec's penalty branch (objective.f90:92-94) */
        if (threadIdx.x == 0) s_val = tk_objfun_finish(ctx, TK_DIV(pen, ctx.sqrt_nm), pen);
        __syncthreads();
        return s_val;
    }
    if (blockDim.x >= 512) tk_smm_dfovec_block_ws(S, theta, fsh, ysh);
    else tk_smm_dfovec_block(S, theta, fsh, ysh);
    if (threadIdx.x == 0) {
        double dot = 0.0;
        for (int m = 0; m < nm; m++) dot = TK_ADD(dot, TK_MUL(fsh[m], fsh[m]));
        s_val = TK_ADD(dot, pen);
    }
    __syncthreads();
    return s_val;
}

__global__ void __launch_bounds__(512)
smm_wave_kernel(const SmmDev* Sp, const TkTablesG T, const TkScalars sc, const TkWave wv, const int64_t* list, int64_t nlist) {
    extern __shared__ double sh[];
    __shared__ double theta[8];
    const SmmDev S = *Sp;
    double* ysh = sh;
    double* fsh = sh + 2 * TK_SMM_SMEM_DOUBLES(S.T);
    const tk_obj_ctx ctx = tk_make_ctx(sc, sc.nx, T.bl, T.bu, T.aux);
    for (int64_t w = blockIdx.x; w < nlist; w += gridDim.x) {
        const int64_t n = list[w];
        __syncthreads();
        if (threadIdx.x < sc.nx) {
            const uint32_t g = (uint32_t)n ^ ((uint32_t)n >> 1);
            uint32_t q = 0;
            for (int j = 0; j < TK_MAXBIT; j++)
                if ((g >> j) & 1u) q ^= T.V[(size_t)threadIdx.x * 32 + j];
            theta[threadIdx.x] = __dadd_rn(T.lo[threadIdx.x], __dmul_rn(tk_q2unit(q), T.w[threadIdx.x]));
        }
        __syncthreads();
        const double f = smm_objfun_block(S, ctx, theta, ysh, fsh);
        if (threadIdx.x == 0) {
            wv.fval[tk_local_index(n, wv.world)] = f;
            if (f < sc.fvalmax) atomicAdd(&wv.counts[n / TIKTAK_SOBOL_BLOCK], 1ULL);
        }
    }
}

__global__ void __launch_bounds__(512)
smm_objfun_kernel(const SmmDev* Sp, const TkTablesG T, const TkScalars sc, const double* x, int64_t n, double* f) {
    extern __shared__ double sh[];
    __shared__ double theta[8];
    const SmmDev S = *Sp;
    double* ysh = sh;
    double* fsh = sh + 2 * TK_SMM_SMEM_DOUBLES(S.T);
    const tk_obj_ctx ctx = tk_make_ctx(sc, sc.nx, T.bl, T.bu, T.aux);
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < sc.nx) theta[threadIdx.x] = x[(size_t)threadIdx.x * n + r];
        __syncthreads();
        const double v = smm_objfun_block(S, ctx, theta, ysh, fsh);
        if (threadIdx.x == 0) f[r] = v;
    }
}

size_t smm_shmem(int T, int nm) { return (size_t)(TK_SMM_SMEM_DOUBLES(T) + nm + 8) * sizeof(double); }       /* 256-thread blocks */
size_t smm_shmem_ws(int T, int nm) { return (size_t)(2 * TK_SMM_SMEM_DOUBLES(T) + nm + 8) * sizeof(double); } /* 512-thread blocks */

} /* namespace */

int tk_smm_init(tiktak_handle* h) {
    const tiktak_smm_params* p = (const tiktak_smm_params*)h->cfg.user;
    if (!p) { tk_set_error("TIKTAK_OBJ_SMM needs tiktak_config.user = tiktak_smm_params"); return TIKTAK_ERR_ARG; }
    const int np = p->np, T = p->T, nm = TK_SMM_NSERIES * TK_SMM_NSTAT * T;
    if (h->cfg.nx != 7 || h->cfg.nm != nm || np < TK_SMM_CHUNK || np % TK_SMM_CHUNK || T < 11 || T > TK_SMM_TMAX) {
        tk_set_error("TIKTAK_OBJ_SMM: need nx=7, nm=30*T, np multiple of 256, 11<=T<=40 (got nx=%d nm=%d np=%d T=%d)", h->cfg.nx,
                     h->cfg.nm, np, T);
        return TIKTAK_ERR_ARG;
    }
    SmmState* st = new SmmState();
    h->smm = st;
    const size_t n = (size_t)np * T;
    TK_CUDA(cudaMalloc(&st->d_eta0, n * 8)); TK_CUDA(cudaMalloc(&st->d_eps0, n * 8)); TK_CUDA(cudaMalloc(&st->d_u, n * 8));
    TK_CUDA(cudaMalloc(&st->d_alpha0, (size_t)np * 8)); TK_CUDA(cudaMalloc(&st->d_mdata, (size_t)nm * 8));
    TK_CUDA(cudaMalloc(&st->d_scale, (size_t)nm * 8)); TK_CUDA(cudaMalloc(&st->d_dev, sizeof(SmmDev)));
    st->dev = SmmDev{np, T, st->d_eta0, st->d_eps0, st->d_u, st->d_alpha0, st->d_mdata, st->d_scale};
    /* data moments: the same simulator at theta_star with seed + 1, mdata = 0 and scale = 1 meanwhile */
    std::vector<double> zeros(nm, 0.0), ones(nm, 1.0), md(nm), scl(nm);
    TK_CUDA(cudaMemcpyAsync(st->d_mdata, zeros.data(), (size_t)nm * 8, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaMemcpyAsync(st->d_scale, ones.data(), (size_t)nm * 8, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaMemcpyAsync(st->d_dev, &st->dev, sizeof(SmmDev), cudaMemcpyHostToDevice, h->stream));
    int rc = fill_and_upload(*st, np, T, (uint32_t)p->seed + 1u, h->stream);
    if (rc) return rc;
    double *d_th = nullptr, *d_out = nullptr;
    TK_CUDA(cudaMalloc(&d_th, 7 * 8)); TK_CUDA(cudaMalloc(&d_out, (size_t)nm * 8));
    TK_CUDA(cudaMemcpyAsync(d_th, p->theta_star, 7 * 8, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaFuncSetAttribute(smm_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smm_shmem(T, nm)));
    smm_moments_kernel<<<1, 256, smm_shmem(T, nm), h->stream>>>(st->d_dev, d_th, d_out);
    h->launches++;
    TK_CUDA(cudaMemcpyAsync(md.data(), d_out, (size_t)nm * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_th); cudaFree(d_out);
    for (int m = 0; m < nm; m++) scl[m] = 1.0 + fabs(md[m]);
    TK_CUDA(cudaMemcpyAsync(st->d_mdata, md.data(), (size_t)nm * 8, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaMemcpyAsync(st->d_scale, scl.data(), (size_t)nm * 8, cudaMemcpyHostToDevice, h->stream));
    rc = fill_and_upload(*st, np, T, (uint32_t)p->seed, h->stream); /* the simulation shocks */
    if (rc) return rc;
    h->sc.user = st->d_dev;
    h->smm_T = T;
    return TIKTAK_OK;
}

void tk_smm_free(tiktak_handle* h) {
    SmmState* st = (SmmState*)h->smm;
    if (!st) return;
    void* ptrs[] = {st->d_eta0, st->d_eps0, st->d_u, st->d_alpha0, st->d_mdata, st->d_scale, st->d_dev};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete st;
    h->smm = nullptr;
}

int tk_launch_smm_wave(tiktak_handle* h, const TkWave& wv) {
    /* the owned points of the wave, explicitly (evaluations are expensive: one block each) */
    std::vector<int64_t> list;
    for (int64_t cb = wv.cb0; cb < wv.cb0 + wv.ncb; cb++) {
        if (cb % wv.world != wv.rank) continue;
        int64_t lo = std::max<int64_t>(1, cb * (int64_t)TIKTAK_SOBOL_BLOCK);
        int64_t hi = std::min<int64_t>(wv.N, (cb + 1) * (int64_t)TIKTAK_SOBOL_BLOCK - 1);
        if (wv.n_hi >= 0) { lo = std::max(lo, wv.n_lo); hi = std::min(hi, wv.n_hi); }
        for (int64_t n = lo; n <= hi; n++) list.push_back(n);
    }
    if (list.empty()) return TIKTAK_OK;
    int64_t* d_list = nullptr;
    TK_CUDA(cudaMalloc(&d_list, list.size() * 8));
    TK_CUDA(cudaMemcpyAsync(d_list, list.data(), list.size() * 8, cudaMemcpyHostToDevice, h->stream));
    const size_t shm = smm_shmem_ws(h->smm_T, h->cfg.nm);
    TK_CUDA(cudaFuncSetAttribute(smm_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    const unsigned grid = (unsigned)std::min<int64_t>((int64_t)list.size(), (int64_t)h->sm_count); /* one 512-thread block per SM */
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    smm_wave_kernel<<<grid, 512, shm, h->stream>>>((const SmmDev*)h->sc.user, T, h->sc, wv, d_list, (int64_t)list.size());
    h->launches++;
    TK_CUDA(cudaGetLastError());
    TK_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_list);
    return TIKTAK_OK;
}

int tk_launch_smm_objfun(tiktak_handle* h, const double* d_x, int64_t n, double* d_f) {
    const size_t shm = smm_shmem_ws(h->smm_T, h->cfg.nm);
    TK_CUDA(cudaFuncSetAttribute(smm_objfun_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    const unsigned grid = (unsigned)std::min<int64_t>(n, (int64_t)h->sm_count);
    const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
    smm_objfun_kernel<<<grid, 512, shm, h->stream>>>((const SmmDev*)h->sc.user, T, h->sc, d_x, n, d_f);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}
