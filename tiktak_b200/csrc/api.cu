/* api.cu -- the C ABI of libtiktak_cuda (include/tiktak_cuda.h): handle, config, stage drivers.
 *
 * Host-side restatements in this file (small, run once per handle):
 *   - insobl's direction-number recurrence (utilities.f90:981-1031) -> 32-bit scaled table V
 *   - the blend weight w11 (TiktakGlobalSearch.f90:718, single precision)
 *   - simplex_coordinates2 (simplex.f90:346-438) and the simplex scale K**(1/nx) (:623)
 *   - parseConfig's defaults (genericParams.f90:307-338)
 * There is no CPU fallback: without a CUDA device tiktak_cuda_create fails with TIKTAK_ERR_NOGPU.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "tk_common.cuh"
#include "smm.cuh"
#include "../../tables/sobol_bf1111.inc"

static thread_local char g_err[1024] = "";
void tk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int tk_export_candidates(tiktak_handle* h, double* d_cand);
int tk_merge_candidates(tiktak_handle* h, const double* d_cand);

namespace {

/* direction numbers m_{d,j} (odd, < 2^j) from the primitive polynomial + initial values, then
 * V[d][j-1] = m_{d,j} * 2^(32-j): the reference's v(d,j)*2^(maxcol-j) and recipd = 2^-maxcol
 * (utilities.f90:1027-1035) give the same dyadic rational for every maxcol. */
void build_dirnums(int nx, std::vector<uint32_t>& V) {
    V.assign((size_t)nx * 32, 0u);
    for (int d = 0; d < nx; d++) {
        uint64_t m[33];
        if (d == 0) {
            for (int j = 1; j <= 32; j++) m[j] = 1; /* :977 */
        } else {
            const unsigned poly = tk_sobol_poly[d - 1];
            int deg = 0;
            for (unsigned t = poly >> 1; t; t >>= 1) deg++;
            for (int j = 1; j <= deg; j++) m[j] = tk_sobol_vinit[d - 1][j - 1];
            for (int j = deg + 1; j <= 32; j++) {
                uint64_t nv = m[j - deg];
                for (int k = 1; k <= deg; k++)
                    if ((poly >> (deg - k)) & 1u) nv ^= (m[j - k] << k); /* includ(k), :1013-1021 */
                m[j] = nv;
            }
        }
        for (int j = 1; j <= 32; j++) V[(size_t)d * 32 + (j - 1)] = (uint32_t)(m[j] << (32 - j));
    }
}

void build_simplex(int n, std::vector<double>& x) { /* simplex.f90:412-435; (n, n+1) column-major */
    x.assign((size_t)n * (n + 1), 0.0);
    for (int j = 0; j < n; j++) x[(size_t)j * n + j] = 1.0;
    const double a = (1.0 - sqrt(1.0 + (double)n)) / (double)n;
    for (int i = 0; i < n; i++) x[(size_t)n * n + i] = a;
    std::vector<double> c(n);
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = 0; j <= n; j++) s += x[(size_t)j * n + i];
        c[i] = s / (double)(n + 1);
    }
    for (int j = 0; j <= n; j++)
        for (int i = 0; i < n; i++) x[(size_t)j * n + i] -= c[i];
    double ss = 0.0;
    for (int i = 0; i < n; i++) ss += x[i] * x[i];
    const double s = sqrt(ss);
    for (size_t e = 0; e < x.size(); e++) x[e] /= s;
}

template <class T>
int dalloc(T** p, size_t n) {
    TK_CUDA(cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)));
    return TIKTAK_OK;
}
template <class T>
int upload(T** p, const std::vector<T>& v, cudaStream_t s) {
    int rc = dalloc(p, v.size());
    if (rc) return rc;
    if (!v.empty()) TK_CUDA(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return TIKTAK_OK;
}

int begin_timer(tiktak_handle* h) { TK_CUDA(cudaEventRecord(h->ev0, h->stream)); return TIKTAK_OK; }
int end_timer(tiktak_handle* h, double* ms) {
    TK_CUDA(cudaEventRecord(h->ev1, h->stream));
    TK_CUDA(cudaEventSynchronize(h->ev1));
    float f = 0;
    TK_CUDA(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    *ms = f;
    return TIKTAK_OK;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

} /* namespace */

extern "C" {

static int tk_est_dfpmin(tiktak_handle* h, std::vector<double>& p, double* fret_out, int* its_out, long* nfev);

const char* tiktak_cuda_last_error(void) { return g_err; }
const char* tiktak_cuda_version(void) { return "tiktak_b200 0.1 (sm_100a)"; }
const char* tiktak_cuda_strerror(int code) {
    switch (code) {
        case TIKTAK_OK: return "ok";
        case TIKTAK_ERR_ARG: return "invalid argument or inconsistent configuration";
        case TIKTAK_ERR_CUDA: return "CUDA runtime error";
        case TIKTAK_ERR_STATE: return "stage called out of order";
        case TIKTAK_ERR_NOGPU: return "no CUDA device available (libtiktak_cuda has no CPU fallback)";
        case TIKTAK_ERR_UNSUPPORTED: return "feature not supported by this build";
    }
    return "unknown status";
}

int tiktak_cuda_config_defaults(int32_t nx, int32_t* maxeval, int32_t* qr_ndraw, int32_t* legitimate,
                                double* fvalmax, int32_t* maxpoints, int32_t* lottery_points) {
    if (!maxeval || !qr_ndraw || !legitimate || !fvalmax || !maxpoints || !lottery_points) return TIKTAK_ERR_ARG;
    if (*maxeval < 0) *maxeval = 40 * (nx + 1);             /* genericParams.f90:307-309 */
    if (*qr_ndraw < 0) *qr_ndraw = 500;                      /* :313-315 */
    if (*legitimate < 0) *legitimate = *qr_ndraw + 10000;    /* :317-319 */
    if (*fvalmax < 0.0) *fvalmax = 100000000.0;              /* :321-323 */
    if (*qr_ndraw == 0) *maxpoints = 0;                      /* :325-326 */
    else if (*maxpoints < 0) *maxpoints = std::min(200, *qr_ndraw); /* :328-329 */
    else if (*maxpoints > *qr_ndraw) {                       /* :331-333 `stop 0` */
        tk_set_error("config: maxpoints > # sobol points");
        return TIKTAK_ERR_ARG;
    }
    if (*lottery_points < 0) *lottery_points = *maxpoints;   /* :336-338 */
    return TIKTAK_OK;
}

int tiktak_cuda_create(tiktak_handle** out, const tiktak_config* cfg) {
    if (!out || !cfg) return TIKTAK_ERR_ARG;
    *out = nullptr;
    if (cfg->nx < 1 || cfg->nx > TK_SOBOL_MAXDIM || cfg->nm < 1 || !cfg->range || !cfg->init || !cfg->bound ||
        cfg->maxpoints_plus1 < 1 || cfg->qr_ndraw < 0 || cfg->qr_ndraw >= (1 << TK_MAXBIT) || cfg->world < 1 ||
        cfg->rank < 0 || cfg->rank >= cfg->world) {
        tk_set_error("create: invalid configuration");
        return TIKTAK_ERR_ARG;
    }
    if (cfg->maxpoints_plus1 - 1 > cfg->qr_ndraw) { tk_set_error("create: maxpoints > qr_ndraw"); return TIKTAK_ERR_ARG; }
    if (cfg->search_type != 0 && cfg->search_type != 1) { /* getModifiedParam :709-715 `stop 0` */
        tk_set_error("create: unknown selectionMethod %d", cfg->search_type);
        return TIKTAK_ERR_ARG;
    }
    if (cfg->search_type == 1 && cfg->lottery_points < 1) { tk_set_error("create: lotteryPoints must be >= 1"); return TIKTAK_ERR_ARG; }
    if (cfg->sobol_quirk != 0 && cfg->qr_ndraw > 33554435) {
        /* the reference's rightmost-zero search goes through single-precision REAL() (utilities.f90:1090) and degenerates to a
         * 4-point cycle from Sobol index 33,554,436 on (SURVEY finding 2); the device generates the intended sequence only.
         * Up to that index both sequences are identical, so the flag is accepted there. */
        tk_set_error("create: sobol_quirk = 1 with qr_ndraw > 33554435 is not built (the degenerate 4-point cycle of the reference); use sobol_quirk = 0");
        return TIKTAK_ERR_UNSUPPORTED;
    }
    if (cfg->nx > TK_NX_GENERIC_MAX) { tk_set_error("create: nx > %d not supported by this build", TK_NX_GENERIC_MAX); return TIKTAK_ERR_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        tk_set_error("create: no CUDA device; libtiktak_cuda has no CPU fallback");
        return TIKTAK_ERR_NOGPU;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { tk_set_error("create: device %d of %d", cfg->device, ndev); return TIKTAK_ERR_ARG; }
    TK_CUDA(cudaSetDevice(cfg->device));
    tiktak_handle* h = new tiktak_handle();
    h->cfg = *cfg;
    const int nx = cfg->nx, K = cfg->maxpoints_plus1;
    h->range.assign(cfg->range, cfg->range + 2 * nx);
    h->init.assign(cfg->init, cfg->init + nx);
    h->bound.assign(cfg->bound, cfg->bound + 2 * nx);
    h->cfg.range = h->range.data(); h->cfg.init = h->init.data(); h->cfg.bound = h->bound.data();
    if (h->cfg.nm_itmax <= 0) h->cfg.nm_itmax = 5000;
    if (!(h->cfg.nm_ftol > 0)) h->cfg.nm_ftol = 1.0e-4;
    if (h->cfg.round_size < 1) h->cfg.round_size = 1;
    h->device = cfg->device;
    cudaDeviceProp prop;
    TK_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    TK_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TK_CUDA(cudaEventCreate(&h->ev0));
    TK_CUDA(cudaEventCreate(&h->ev1));
    /* obj_initialize (objective.f90:31-37) */
    h->sc.nx = nx; h->sc.nm = cfg->nm;
    h->sc.fvalmax = cfg->fvalmax;
    h->sc.penalty0 = cfg->fvalmax;
    h->sc.penaltyc = 10 * cfg->fvalmax;
    h->sc.max_penalty = 1000.0 * h->sc.penalty0;
    h->sc.sqrt_nm = sqrt((double)cfg->nm);
    h->sc.user = nullptr;
    h->aux.assign(nx, 0.0);
    int rc = TIKTAK_OK;
    if (cfg->objective_id != TIKTAK_OBJ_SMM)
        rc = tk_dispatch_obj(cfg->objective_id, [&]<class Obj>() -> int { Obj::host_init(nx, h->aux.data()); return TIKTAK_OK; });
    if (rc) { delete h; return rc; }
    build_dirnums(nx, h->V);
    build_simplex(nx, h->simplex);
    h->nm_scale = pow((double)K, 1.0 / nx); /* DBLE(p_maxpoints)**(1.0_dp/p_nx), :623 */
    h->w11.assign(K + 1, 0.0);
    for (int k = 1; k <= K; k++) { /* :718 -- REAL() is single precision */
        float wf = sqrtf((float)k / (float)K);
        wf = std::min(std::max(0.10f, wf), 0.95f);
        h->w11[k] = (double)wf;
    }
    std::vector<double> lo(h->range.begin(), h->range.begin() + nx), w(nx), bl(h->bound.begin(), h->bound.begin() + nx),
        bu(h->bound.begin() + nx, h->bound.end());
    for (int d = 0; d < nx; d++) w[d] = h->range[nx + d] - h->range[d];
#define TK_TRY(x) do { rc = (x); if (rc) { tiktak_cuda_destroy(h); return rc; } } while (0)
    TK_TRY(upload(&h->d_V, h->V, h->stream));
    TK_TRY(upload(&h->d_lo, lo, h->stream));
    TK_TRY(upload(&h->d_w, w, h->stream));
    TK_TRY(upload(&h->d_width, w, h->stream));
    TK_TRY(upload(&h->d_bl, bl, h->stream));
    TK_TRY(upload(&h->d_bu, bu, h->stream));
    TK_TRY(upload(&h->d_aux, h->aux, h->stream));
    TK_TRY(upload(&h->d_init, h->init, h->stream));
    TK_TRY(upload(&h->d_simplex, h->simplex, h->stream));
    TK_TRY(upload(&h->d_w11, h->w11, h->stream));
    h->n_cblocks = (int64_t)cfg->qr_ndraw / TIKTAK_SOBOL_BLOCK + 1; /* n = 0..N */
    h->n_local_cblocks = (h->n_cblocks > cfg->rank) ? (h->n_cblocks - cfg->rank + cfg->world - 1) / cfg->world : 0;
    TK_TRY(dalloc(&h->d_fval, (size_t)h->n_local_cblocks * TIKTAK_SOBOL_BLOCK));
    TK_TRY(dalloc(&h->d_counts, (size_t)h->n_cblocks + 1));
    TK_TRY(dalloc(&h->d_counts_spare, (size_t)h->n_cblocks + 1));
    TK_CUDA(cudaMemsetAsync(h->d_counts, 0, ((size_t)h->n_cblocks + 1) * sizeof(unsigned long long), h->stream));
    TK_CUDA(cudaMemsetAsync(h->d_counts_spare, 0, ((size_t)h->n_cblocks + 1) * sizeof(unsigned long long), h->stream));
    TK_CUDA(cudaMallocHost(&h->h_counts, ((size_t)h->n_cblocks + 1) * sizeof(unsigned long long)));
    memset(h->h_counts, 0, ((size_t)h->n_cblocks + 1) * sizeof(unsigned long long));
    TK_TRY(dalloc(&h->d_hist, 256));
    TK_TRY(dalloc(&h->d_selstate, 8));
    const size_t ncand = (size_t)std::max(K - 1, 1) * (size_t)std::max(cfg->world, 1);
    TK_TRY(dalloc(&h->d_cand_f, ncand));
    TK_TRY(dalloc(&h->d_cand_i, ncand));
    TK_TRY(dalloc(&h->d_cand_n, 1));
    TK_TRY(dalloc(&h->d_ticket, 1));
    TK_TRY(dalloc(&h->d_sorted_f, ncand));
    TK_TRY(dalloc(&h->d_sorted_i, ncand));
    TK_TRY(dalloc(&h->d_xstarts, (size_t)K * (nx + 1)));
    TK_TRY(dalloc(&h->d_res, (size_t)(K + 1) * (nx + 1)));
    TK_TRY(dalloc(&h->d_res_valid, (size_t)K + 1));
    TK_CUDA(cudaMemsetAsync(h->d_res_valid, 0, (size_t)(K + 1) * sizeof(int), h->stream));
    h->h_res_f.assign(K + 1, 0.0);
    h->h_res_valid.assign(K + 1, 0);
    h->h_res_imp.assign(K + 1, 0);
    TK_TRY(dalloc(&h->d_best, (size_t)nx + 4));
    TK_TRY(dalloc(&h->d_starts, (size_t)K * nx));
    TK_TRY(dalloc(&h->d_out_x, (size_t)K * nx));
    TK_TRY(dalloc(&h->d_out_f, (size_t)K));
    TK_TRY(dalloc(&h->d_out_evals, (size_t)K));
    TK_TRY(dalloc(&h->d_arch_starts, (size_t)K * nx));
    TK_TRY(dalloc(&h->d_arch_evals, (size_t)K + 1));
    TK_TRY(dalloc(&h->d_wshr, (size_t)nx));
    TK_TRY(dalloc(&h->d_shr_have, 1));
    TK_CUDA(cudaMemsetAsync(h->d_shr_have, 0, sizeof(int), h->stream));
    TK_CUDA(cudaMemsetAsync(h->d_arch_evals, 0, (size_t)(K + 1) * sizeof(int), h->stream));
    TK_CUDA(cudaEventCreate(&h->evl0));
    TK_CUDA(cudaEventCreate(&h->evl1));
    TK_CUDA(cudaEventCreate(&h->evp0));
    TK_CUDA(cudaEventCreate(&h->evp1));
#undef TK_TRY
    TK_CUDA(cudaStreamSynchronize(h->stream));
    if (cfg->objective_id == TIKTAK_OBJ_SMM) {
        rc = tk_smm_init(h);
        if (rc) { tiktak_cuda_destroy(h); return rc; }
    }
    *out = h;
    return TIKTAK_OK;
}

void tiktak_cuda_destroy(tiktak_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    for (auto& g : h->stage_graph) if (g) cudaGraphExecDestroy(g);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->h_fetch) cudaFreeHost(h->h_fetch);
    tk_comm_destroy(h);
    tk_smm_free(h);
    void* ptrs[] = {h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux, h->d_init, h->d_simplex, h->d_w11, h->d_fval,
                    h->d_counts, h->d_hist, h->d_selstate, h->d_cand_f, h->d_cand_i, h->d_cand_n, h->d_sorted_f,
                    h->d_sorted_i, h->d_xstarts, h->d_res, h->d_res_valid, h->d_best, h->d_starts, h->d_out_x,
                    h->d_out_f, h->d_width, h->d_out_evals, h->d_dfnls_slab, h->d_rho, h->d_out_status, h->d_wshr,
                    h->d_shr_have, h->d_arch_starts, h->d_arch_evals, h->d_lot_f, h->d_lot_sf, h->d_lot_cum,
                    h->d_lot_k, h->d_lot_sk, h->d_lot_point, h->d_ticket, h->d_accepted, h->d_nmq, h->d_counts_spare,
                    h->d_objfun_x, h->d_objfun_f, h->d_async_ev, h->d_async_clk, h->d_async_base, h->d_async_desc, h->d_nm_pslab, h->d_sobtab, h->d_sel_list_key, h->d_sel_list_n, h->d_xsend, h->d_xrecv};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evw0) cudaEventDestroy(h->evw0);
    if (h->evw1) cudaEventDestroy(h->evw1);
    if (h->evp0) cudaEventDestroy(h->evp0);
    if (h->evp1) cudaEventDestroy(h->evp1);
    if (h->evl0) cudaEventDestroy(h->evl0);
    if (h->evl1) cudaEventDestroy(h->evl1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

void* tiktak_cuda_stream(tiktak_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t tiktak_cuda_launch_count(tiktak_handle* h) { return h ? h->launches : 0; }

int tiktak_cuda_stage_ms(tiktak_handle* h, const char* stage, double* ms) {
    if (!h || !stage || !ms) return TIKTAK_ERR_ARG;
    if (!strcmp(stage, "pretest")) {
        if (h->pretest_timing_open) {
            TK_CUDA(cudaEventSynchronize(h->ev_stage_end));
            float w = 0; cudaEventElapsedTime(&w, h->evp0, h->ev_stage_end); h->ms_pretest = w; h->pretest_timing_open = false;
        }
        *ms = h->ms_pretest;
    }
    else if (!strcmp(stage, "pretest_eval")) {
        if (h->wave_pending) {
            TK_CUDA(cudaEventSynchronize(h->evw1));
            float w = 0; cudaEventElapsedTime(&w, h->ev_wave_start, h->evw1); h->ms_pretest_eval += w; h->wave_pending = false;
        }
        *ms = h->ms_pretest_eval;
    }
    else if (!strcmp(stage, "choose")) *ms = h->ms_choose;
    else if (!strcmp(stage, "local")) {
        if (h->local_timing_open) { /* close the interval opened by the first local_enqueue since the last query */
            TK_CUDA(cudaEventSynchronize(h->evl1));
            float w = 0; cudaEventElapsedTime(&w, h->evl0, h->evl1); h->ms_local = w; h->local_timing_open = false;
        }
        *ms = h->ms_local;
    }
    else return TIKTAK_ERR_ARG;
    return TIKTAK_OK;
}

int tiktak_cuda_device_buffer(tiktak_handle* h, const char* name, void** ptr, int64_t* bytes) {
    if (!h || !name || !ptr) return TIKTAK_ERR_ARG;
    const int nx = h->cfg.nx, K = h->cfg.maxpoints_plus1;
    int64_t b = 0;
    if (!strcmp(name, "fval")) { *ptr = h->d_fval; b = h->n_local_cblocks * (int64_t)TIKTAK_SOBOL_BLOCK * 8; }
    else if (!strcmp(name, "counts")) { *ptr = h->d_counts; b = h->n_cblocks * 8; }
    else if (!strcmp(name, "x_starts")) { *ptr = h->d_xstarts; b = (int64_t)K * (nx + 1) * 8; }
    else if (!strcmp(name, "results")) { *ptr = h->d_res; b = (int64_t)(K + 1) * (nx + 1) * 8; }
    else return TIKTAK_ERR_ARG;
    if (bytes) *bytes = b;
    return TIKTAK_OK;
}

int tiktak_cuda_sobol_points(tiktak_handle* h, int64_t first, int64_t count, double* out) {
    if (!h || !out || first < 1 || count < 0 || first + count - 1 >= ((int64_t)1 << TK_MAXBIT)) return TIKTAK_ERR_ARG;
    if (count == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    const size_t bytes = (size_t)count * h->cfg.nx * sizeof(double);
    if (is_device_ptr(out)) {
        int rc = tk_launch_sobol_points(h, first, count, out);
        if (rc) return rc;
        TK_CUDA(cudaStreamSynchronize(h->stream));
        return TIKTAK_OK;
    }
    double* d = nullptr;
    TK_CUDA(cudaMalloc(&d, bytes));
    int rc = tk_launch_sobol_points(h, first, count, d);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { tk_set_error("sobol_points copy: %s", cudaGetErrorString(e)); rc = TIKTAK_ERR_CUDA; }
    }
    cudaFree(d);
    return rc;
}

int tiktak_cuda_objfun(tiktak_handle* h, const double* x, int64_t n, double* fval) {
    if (!h || !x || !fval || n < 0) return TIKTAK_ERR_ARG;
    if (n == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    const size_t xb = (size_t)n * h->cfg.nx * sizeof(double), fb = (size_t)n * sizeof(double);
    const bool xdev = is_device_ptr(x), fdev = is_device_ptr(fval);
    /* staging for host arguments lives in the handle (lastSearch calls this hundreds of times with a few points) */
    if ((!xdev || !fdev) && h->objfun_scratch_pts < n) {
        if (h->d_objfun_x) cudaFree(h->d_objfun_x);
        if (h->d_objfun_f) cudaFree(h->d_objfun_f);
        h->d_objfun_x = h->d_objfun_f = nullptr; h->objfun_scratch_pts = 0;
        const int64_t cap = std::max<int64_t>(n, 256);
        TK_CUDA(cudaMalloc(&h->d_objfun_x, (size_t)cap * h->cfg.nx * sizeof(double)));
        TK_CUDA(cudaMalloc(&h->d_objfun_f, (size_t)cap * sizeof(double)));
        h->objfun_scratch_pts = cap;
    }
    double* dx = xdev ? nullptr : h->d_objfun_x;
    double* df = fdev ? nullptr : h->d_objfun_f;
    if (!xdev) TK_CUDA(cudaMemcpyAsync(dx, x, xb, cudaMemcpyHostToDevice, h->stream));
    int rc = (h->cfg.objective_id == TIKTAK_OBJ_SMM) ? tk_launch_smm_objfun(h, xdev ? x : dx, n, fdev ? fval : df)
                                                      : tk_launch_objfun(h, xdev ? x : dx, n, fdev ? fval : df);
    if (!rc && !fdev) {
        cudaError_t e = cudaMemcpyAsync(fval, df, fb, cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) { tk_set_error("objfun copy: %s", cudaGetErrorString(e)); rc = TIKTAK_ERR_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess && !rc) { tk_set_error("objfun: %s", cudaGetErrorString(e)); rc = TIKTAK_ERR_CUDA; }
    return rc;
}

/* evaluate the owned count blocks of [cb0, cb0+ncb) */
int tiktak_cuda_pretest_wave(tiktak_handle* h, int64_t cb0, int64_t ncb) {
    if (!h || cb0 < 0 || ncb < 0 || cb0 + ncb > h->n_cblocks) return TIKTAK_ERR_ARG;
    TK_CUDA(cudaSetDevice(h->device));
    const bool first_wave = (cb0 == 0);
    if (first_wave) { /* the stage starts here: its device time runs from evp0 to the evp1 of the last wave */
        std::swap(h->d_counts, h->d_counts_spare); /* a buffer zeroed behind the previous job's kernel: no memset in front */
        if (ncb == 0) { TK_CUDA(cudaEventRecord(h->evp0, h->stream)); TK_CUDA(cudaEventRecord(h->evp1, h->stream)); h->ev_stage_end = h->evp1; }
        h->pretest_timing_open = true;
        h->cblocks_done = 0; h->kcut = -1; h->pretest_done = false; h->starts_ready = false;
        h->ms_pretest_eval = 0; h->wave_pending = false;
    }
    if (ncb == 0) {
        if (first_wave) TK_CUDA(cudaMemsetAsync(h->d_counts_spare, 0, (size_t)h->n_cblocks * sizeof(unsigned long long), h->stream));
        return TIKTAK_OK;
    }
    TkWave wv;
    wv.N = h->cfg.qr_ndraw; wv.cb0 = cb0; wv.ncb = ncb; wv.rank = h->cfg.rank; wv.world = h->cfg.world;
    wv.fval = h->d_fval; wv.counts = h->d_counts;
    if (!h->evw0) { TK_CUDA(cudaEventCreate(&h->evw0)); TK_CUDA(cudaEventCreate(&h->evw1)); }
    if (h->wave_pending) { /* fold the previous wave's duration in before the events are reused */
        TK_CUDA(cudaEventSynchronize(h->evw1));
        float ms = 0; cudaEventElapsedTime(&ms, h->ev_wave_start, h->evw1); h->ms_pretest_eval += ms; h->wave_pending = false;
    }
    /* one event in front of the wave and one behind it: the first wave's start event is the stage's start event and the
     * last wave's end event is the stage's end event (an event record costs ~1.3 us of stream time on this GPU, so a
     * second pair around the same kernel would add 2.6 us to a 33 us stage) */
    h->ev_wave_start = first_wave ? h->evp0 : h->evw0;
    int rc = TIKTAK_OK;
    /* A stage that is ONE wave (no cut possible) goes to the device as one CUDA graph from its second run on: start event,
     * kernel, end event and the memset of the spare counts are four stream commands, and with an idle stream the host
     * time between them lies inside the event interval -- 2.4 us of a 31 us stage at C2 (tools/micro/launch_floor.cu,
     * profiles/r02_sobol_dealing.txt).  Two graphs: the job's counts buffer alternates between d_counts and its twin. */
    static const bool no_graph = [] { const char* e = getenv("TIKTAK_NO_GRAPH"); return e && atoi(e) != 0; }();
    const bool whole_stage = first_wave && ncb == h->n_cblocks && h->cfg.objective_id != TIKTAK_OBJ_SMM && !no_graph;
    const int gi = whole_stage ? ((h->d_counts == h->stage_graph_counts[0]) ? 0 : (h->d_counts == h->stage_graph_counts[1]) ? 1 : -1) : -1;
    if (gi >= 0 && h->stage_graph[gi]) {
        TK_CUDA(cudaGraphLaunch(h->stage_graph[gi], h->stream));
        h->launches++;
    } else {
        /* captured from the second run on (the first one builds tables and asks for the occupancy: not capturable) */
        int slot = -1;
        if (whole_stage && h->stage_runs >= 1) {
            slot = h->stage_graph_counts[0] ? (h->stage_graph_counts[1] ? -1 : 1) : 0;
            if (slot >= 0 && cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); slot = -1; }
        }
        if (slot >= 0) TK_CUDA(cudaEventRecordWithFlags(h->ev_wave_start, h->stream, cudaEventRecordExternal));
        else TK_CUDA(cudaEventRecord(h->ev_wave_start, h->stream));
        rc = (h->cfg.objective_id == TIKTAK_OBJ_SMM) ? tk_launch_smm_wave(h, wv) : tk_launch_sobol_wave(h, wv);
        if (slot >= 0) TK_CUDA(cudaEventRecordWithFlags(h->evw1, h->stream, cudaEventRecordExternal));
        else TK_CUDA(cudaEventRecord(h->evw1, h->stream));
        if (whole_stage) /* every count to the pinned mirror: the stage's callers read them right behind it */
            TK_CUDA(cudaMemcpyAsync(h->h_counts, h->d_counts, (size_t)h->n_cblocks * 8, cudaMemcpyDeviceToHost, h->stream));
        if (first_wave) /* clear the other counts buffer for the next job, behind this wave */
            TK_CUDA(cudaMemsetAsync(h->d_counts_spare, 0, (size_t)h->n_cblocks * sizeof(unsigned long long), h->stream));
        if (slot >= 0) {
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(h->stream, &g);
            if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&h->stage_graph[slot], g, 0);
            if (g) cudaGraphDestroy(g);
            if (e != cudaSuccess || rc) { tk_set_error("pre-test stage graph: %s", cudaGetErrorString(e)); cudaGetLastError(); return rc ? rc : TIKTAK_ERR_CUDA; }
            h->stage_graph_counts[slot] = h->d_counts;
            TK_CUDA(cudaGraphLaunch(h->stage_graph[slot], h->stream));
        }
        if (whole_stage) h->stage_runs++;
    }
    h->ev_stage_end = h->evw1;
    h->wave_pending = true;
    h->counts_copied = whole_stage;
    if (rc) return rc;
    h->cblocks_done = std::max(h->cblocks_done, cb0 + ncb);
    return TIKTAK_OK;
}

int tiktak_cuda_pretest_counts(tiktak_handle* h, int64_t* counts, int64_t nblocks) {
    if (!h || !counts || nblocks < 0 || nblocks > h->n_cblocks) return TIKTAK_ERR_ARG;
    TK_CUDA(cudaSetDevice(h->device));
    static_assert(sizeof(int64_t) == sizeof(unsigned long long), "count width");
    TK_CUDA(cudaMemcpyAsync(counts, h->d_counts, (size_t)nblocks * 8, cudaMemcpyDefault, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return TIKTAK_OK;
}

/* exact index of the `need`-th valid point of count block cb (owner only; others get 0) */
int tiktak_cuda_pretest_find(tiktak_handle* h, int64_t cb, int64_t need, int64_t* n_out) {
    if (!h || !n_out || cb < 0 || cb >= h->n_cblocks || need < 1) return TIKTAK_ERR_ARG;
    *n_out = 0;
    if (cb % h->cfg.world != h->cfg.rank) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    int rc = tk_launch_find_cut(h, cb, (unsigned long long)need, h->d_selstate);
    if (rc) return rc;
    unsigned long long r = 0;
    TK_CUDA(cudaMemcpyAsync(&r, h->d_selstate, 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    *n_out = (int64_t)r;
    return TIKTAK_OK;
}

int tiktak_cuda_pretest_set_cut(tiktak_handle* h, int64_t kcut, int64_t n_legit) {
    if (!h || kcut < 0 || kcut > h->cfg.qr_ndraw) return TIKTAK_ERR_ARG;
    h->kcut = kcut;
    h->nlegit = n_legit;
    h->pretest_done = true;
    return TIKTAK_OK;
}

/* exchange buffers of the handle's communicator: send `n` doubles per rank, receive world * n */
static int comm_buffers(tiktak_handle* h, size_t n) {
    if (h->xbuf_doubles >= n) return TIKTAK_OK;
    if (h->d_xsend) cudaFree(h->d_xsend);
    if (h->d_xrecv) cudaFree(h->d_xrecv);
    h->d_xsend = h->d_xrecv = nullptr; h->xbuf_doubles = 0;
    TK_CUDA(cudaMalloc(&h->d_xsend, n * sizeof(double)));
    TK_CUDA(cudaMalloc(&h->d_xrecv, n * sizeof(double) * (size_t)h->cfg.world));
    h->xbuf_doubles = n;
    return TIKTAK_OK;
}

/* world > 1 with the handle's communicator: count blocks are dealt block-cyclically; when the `legitimate` cut cannot
 * fire every rank evaluates its blocks and nothing is exchanged (the valid count travels with choose's candidate rows);
 * otherwise blocks are evaluated in waves of 64 per rank with one all-reduce of the per-block valid counts per wave, and
 * the owner of the block that holds the L-th valid point finds its index (solveAtSobolPoints :429-464 on shards) */
static int pretest_sharded(tiktak_handle* h, int64_t* n_evaluated, int64_t* n_legit) {
    const int64_t N = h->cfg.qr_ndraw, L = h->cfg.legitimate;
    int rc;
    h->counts_global = false;
    if (L > N) {
        if ((rc = tiktak_cuda_pretest_wave(h, 0, h->n_cblocks))) return rc;
        h->kcut = N; h->nlegit = -1; h->pretest_done = true;
        if (n_evaluated) *n_evaluated = N;
        if (n_legit) { /* the caller wants the count now: one all-reduce of the shard totals */
            if ((rc = tiktak_cuda_pretest_result(h, nullptr, n_legit))) return rc;
        }
        return TIKTAK_OK;
    }
    const int64_t wave = 64 * (int64_t)h->cfg.world;
    int64_t cum = 0, kcut = -1, nleg = 0;
    for (int64_t cb0 = 0; cb0 < h->n_cblocks && kcut < 0; cb0 += wave) {
        const int64_t ncb = std::min(wave, h->n_cblocks - cb0);
        if ((rc = tiktak_cuda_pretest_wave(h, cb0, ncb))) return rc;
        if ((rc = tk_comm_allreduce_sum_u64(h, h->d_counts + cb0, (size_t)ncb))) return rc;
        TK_CUDA(cudaMemcpyAsync(h->h_counts + cb0, h->d_counts + cb0, (size_t)ncb * 8, cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t cb = cb0; cb < cb0 + ncb; cb++) {
            const int64_t c = (int64_t)h->h_counts[cb];
            if (cum + c >= L) { /* the L-th valid point lives in this count block: its owner finds the index */
                TK_CUDA(cudaMemsetAsync(h->d_selstate, 0, 8, h->stream));
                if (cb % h->cfg.world == h->cfg.rank && (rc = tk_launch_find_cut(h, cb, (unsigned long long)(L - cum), h->d_selstate))) return rc;
                if ((rc = tk_comm_allreduce_max_u64(h, h->d_selstate, 1))) return rc;
                unsigned long long r = 0;
                TK_CUDA(cudaMemcpyAsync(&r, h->d_selstate, 8, cudaMemcpyDeviceToHost, h->stream));
                TK_CUDA(cudaStreamSynchronize(h->stream));
                kcut = (int64_t)r; nleg = L;
                break;
            }
            cum += c;
        }
    }
    if (kcut < 0) { kcut = N; nleg = cum; }
    h->counts_global = true;
    h->kcut = kcut; h->nlegit = nleg; h->pretest_done = true;
    if (n_evaluated) *n_evaluated = kcut;
    if (n_legit) *n_legit = nleg;
    TK_CUDA(cudaEventRecord(h->evp1, h->stream));
    h->ev_stage_end = h->evp1;
    h->pretest_timing_open = true;
    return TIKTAK_OK;
}

int tiktak_cuda_pretest(tiktak_handle* h, int64_t* n_evaluated, int64_t* n_legit) {
    if (!h) return TIKTAK_ERR_ARG;
    TK_CUDA(cudaSetDevice(h->device));
    const int64_t N = h->cfg.qr_ndraw, L = h->cfg.legitimate;
    int rc = TIKTAK_OK;
    auto end_stage = [&]() -> int { /* the find_cut / count reads after the last wave belong to the stage */
        TK_CUDA(cudaEventRecord(h->evp1, h->stream));
        h->ev_stage_end = h->evp1;
        h->pretest_timing_open = true;
        return TIKTAK_OK;
    };
    if (N == 0 || L <= 0) { /* :429: nothing is evaluated */
        tiktak_cuda_pretest_wave(h, 0, 0);
        h->kcut = 0; h->nlegit = 0; h->pretest_done = true;
        if (n_evaluated) *n_evaluated = 0;
        if (n_legit) *n_legit = 0;
        return end_stage();
    }
    if (h->cfg.world > 1 && tk_comm_ready(h)) return pretest_sharded(h, n_evaluated, n_legit);
    const bool can_cut = (L <= N) && h->cfg.world == 1;
    const int64_t wave = can_cut ? 64 : h->n_cblocks;
    int64_t cum = 0, kcut = -1, nleg = 0;
    if (!can_cut && h->cfg.world == 1 && !n_legit) {
        /* legitimate > qr_ndraw: the cut cannot fire, every point is evaluated and nobody waits for the valid
         * count -- queue the evaluation and return; the count is read when it is first asked for */
        rc = tiktak_cuda_pretest_wave(h, 0, h->n_cblocks);
        if (rc) return rc;
        h->kcut = N; h->nlegit = -1; h->pretest_done = true;
        if (n_evaluated) *n_evaluated = N;
        return TIKTAK_OK;
    }
    if (h->cfg.objective_id == TIKTAK_OBJ_SMM && can_cut) {
        /* one evaluation simulates a whole panel: stop as soon as the cut is reached, in sub-waves of a few
         * points per SM instead of whole count blocks (same prefix semantics, :429) */
        if ((rc = tiktak_cuda_pretest_wave(h, 0, 0))) return rc; /* reset state + counts */
        const int64_t chunk = (int64_t)h->sm_count * 8;
        std::vector<unsigned long long> cnt(h->n_cblocks);
        for (int64_t n0 = 1; n0 <= N && kcut < 0; n0 += chunk) {
            TkWave wv;
            wv.N = N; wv.cb0 = 0; wv.ncb = h->n_cblocks; wv.rank = 0; wv.world = 1; wv.fval = h->d_fval; wv.counts = h->d_counts;
            wv.n_lo = n0; wv.n_hi = std::min<int64_t>(N, n0 + chunk - 1);
            if (!h->evw0) { TK_CUDA(cudaEventCreate(&h->evw0)); TK_CUDA(cudaEventCreate(&h->evw1)); }
            TK_CUDA(cudaEventRecord(h->evw0, h->stream));
            if ((rc = tk_launch_smm_wave(h, wv))) return rc;
            TK_CUDA(cudaEventRecord(h->evw1, h->stream));
            TK_CUDA(cudaMemcpyAsync(cnt.data(), h->d_counts, (size_t)h->n_cblocks * 8, cudaMemcpyDeviceToHost, h->stream));
            TK_CUDA(cudaStreamSynchronize(h->stream));
            float wms = 0; cudaEventElapsedTime(&wms, h->evw0, h->evw1); h->ms_pretest_eval += wms;
            h->cblocks_done = h->n_cblocks;
            cum = 0;
            for (int64_t cb = 0; cb < h->n_cblocks; cb++) {
                if (cum + (int64_t)cnt[cb] >= L) {
                    int64_t n = 0;
                    if ((rc = tiktak_cuda_pretest_find(h, cb, L - cum, &n))) return rc;
                    kcut = n; nleg = L;
                    break;
                }
                cum += (int64_t)cnt[cb];
            }
        }
        if (kcut < 0) { kcut = N; nleg = cum; }
        h->kcut = kcut; h->nlegit = nleg; h->pretest_done = true;
        if (n_evaluated) *n_evaluated = kcut;
        if (n_legit) *n_legit = nleg;
        return end_stage();
    }
    for (int64_t cb0 = 0; cb0 < h->n_cblocks && kcut < 0; cb0 += wave) {
        const int64_t ncb = std::min(wave, h->n_cblocks - cb0);
        rc = tiktak_cuda_pretest_wave(h, cb0, ncb);
        if (rc) return rc;
        if (!h->counts_copied) TK_CUDA(cudaMemcpyAsync(h->h_counts + cb0, h->d_counts + cb0, (size_t)ncb * 8, cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream));
        if (h->cfg.world > 1) continue;
        for (int64_t cb = cb0; cb < cb0 + ncb; cb++) {
            const int64_t c = (int64_t)h->h_counts[cb];
            if (can_cut && cum + c >= L) { /* the L-th valid point lives in this count block */
                int64_t n = 0;
                rc = tiktak_cuda_pretest_find(h, cb, L - cum, &n);
                if (rc) return rc;
                kcut = n; nleg = L;
                break;
            }
            cum += c;
        }
    }
    if (h->cfg.world == 1) {
        if (kcut < 0) { kcut = N; nleg = cum; }
        h->kcut = kcut; h->nlegit = nleg; h->pretest_done = true;
    }
    if (n_evaluated) *n_evaluated = h->cfg.world == 1 ? kcut : -1;
    if (n_legit) *n_legit = h->cfg.world == 1 ? nleg : -1;
    return end_stage();
}

/* valid count of the evaluated prefix when it was not needed to place the cut (sum of the per-block counts) */
static int resolve_legit(tiktak_handle* h) {
    if (h->nlegit >= 0 || !h->pretest_done) return TIKTAK_OK;
    if (!h->counts_copied) TK_CUDA(cudaMemcpyAsync(h->h_counts, h->d_counts, (size_t)h->n_cblocks * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    int64_t tot = 0;
    for (int64_t cb = 0; cb < h->n_cblocks; cb++) tot += (int64_t)h->h_counts[cb];
    if (h->cfg.world == 1 || h->counts_global) h->nlegit = tot;
    else if (tk_comm_ready(h)) { /* shard totals summed over the communicator (a collective: every rank calls it) */
        unsigned long long v = (unsigned long long)tot;
        TK_CUDA(cudaMemcpyAsync(h->d_selstate, &v, 8, cudaMemcpyHostToDevice, h->stream));
        int rc = tk_comm_allreduce_sum_u64(h, h->d_selstate, 1);
        if (rc) return rc;
        TK_CUDA(cudaMemcpyAsync(&v, h->d_selstate, 8, cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream));
        h->nlegit = (int64_t)v;
    } /* else: a shard only knows its own count; the caller sums the shards */
    return TIKTAK_OK;
}

int tiktak_cuda_pretest_result(tiktak_handle* h, int64_t* n_evaluated, int64_t* n_legit) {
    if (!h) return TIKTAK_ERR_ARG;
    if (!h->pretest_done) return TIKTAK_ERR_STATE;
    TK_CUDA(cudaSetDevice(h->device));
    int rc = resolve_legit(h);
    if (rc) return rc;
    if (n_evaluated) *n_evaluated = h->kcut;
    if (n_legit) *n_legit = h->nlegit;
    return TIKTAK_OK;
}

int tiktak_cuda_pretest_values_shard(tiktak_handle* h, double* fval_local, int64_t capacity, int64_t* n_local) {
    if (!h || !fval_local) return TIKTAK_ERR_ARG;
    if (!h->pretest_done) return TIKTAK_ERR_STATE;
    const int64_t lastcb = h->kcut / TIKTAK_SOBOL_BLOCK;
    const int64_t nl = (lastcb >= h->cfg.rank) ? ((lastcb - h->cfg.rank) / h->cfg.world + 1) : 0; /* owned blocks touching 0..kcut */
    const int64_t n = nl * (int64_t)TIKTAK_SOBOL_BLOCK;
    if (n_local) *n_local = n;
    if (capacity < n) { tk_set_error("pretest_values_shard: capacity %lld < %lld", (long long)capacity, (long long)n); return TIKTAK_ERR_ARG; }
    TK_CUDA(cudaSetDevice(h->device));
    if (n > 0) TK_CUDA(cudaMemcpyAsync(fval_local, h->d_fval, (size_t)n * 8, cudaMemcpyDefault, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return TIKTAK_OK;
}

/* the same copy on a second stream, behind the evaluation and beside whatever the handle's stream does next (selection,
 * restarts): sobolFnVal.dat is the only consumer of the values on the host, nothing on the path waits for it */
int tiktak_cuda_pretest_values_shard_async(tiktak_handle* h, double* fval_local, int64_t capacity, int64_t* n_local) {
    if (!h || !fval_local) return TIKTAK_ERR_ARG;
    if (!h->pretest_done) return TIKTAK_ERR_STATE;
    const int64_t lastcb = h->kcut / TIKTAK_SOBOL_BLOCK;
    const int64_t nl = (lastcb >= h->cfg.rank) ? ((lastcb - h->cfg.rank) / h->cfg.world + 1) : 0;
    const int64_t n = nl * (int64_t)TIKTAK_SOBOL_BLOCK;
    if (n_local) *n_local = n;
    if (capacity < n) { tk_set_error("pretest_values_shard_async: capacity %lld < %lld", (long long)capacity, (long long)n); return TIKTAK_ERR_ARG; }
    TK_CUDA(cudaSetDevice(h->device));
    if (!h->copy_stream) {
        TK_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        TK_CUDA(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    }
    TK_CUDA(cudaEventRecord(h->ev_copy, h->stream)); /* the values exist once everything queued so far has run */
    TK_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_copy, 0));
    if (n > 0) TK_CUDA(cudaMemcpyAsync(fval_local, h->d_fval, (size_t)n * 8, cudaMemcpyDefault, h->copy_stream));
    return TIKTAK_OK;
}
int tiktak_cuda_pretest_values_wait(tiktak_handle* h) {
    if (!h) return TIKTAK_ERR_ARG;
    if (h->copy_stream) { TK_CUDA(cudaSetDevice(h->device)); TK_CUDA(cudaStreamSynchronize(h->copy_stream)); }
    return TIKTAK_OK;
}

int tiktak_cuda_pretest_values(tiktak_handle* h, int64_t first, int64_t count, double* fval) {
    if (!h || !fval || first < 1 || count < 0) return TIKTAK_ERR_ARG;
    if (!h->pretest_done) return TIKTAK_ERR_STATE;
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    if (h->cfg.world == 1 && first + count - 1 <= h->kcut) { /* one shard, inside the evaluated prefix: one direct copy */
        TK_CUDA(cudaMemcpy(fval, h->d_fval + first, (size_t)count * 8, cudaMemcpyDefault));
        return TIKTAK_OK;
    }
    std::vector<double> tmp((size_t)count, NAN);
    int64_t n = first;
    while (n < first + count) {
        const int64_t cb = n / TIKTAK_SOBOL_BLOCK;
        const int64_t end = std::min<int64_t>((cb + 1) * (int64_t)TIKTAK_SOBOL_BLOCK, first + count);
        const int64_t hi = std::min<int64_t>(end, h->kcut + 1);
        if (cb % h->cfg.world == h->cfg.rank && hi > n && cb < h->cblocks_done)
            TK_CUDA(cudaMemcpy(tmp.data() + (n - first), h->d_fval + tk_local_index(n, h->cfg.world), (size_t)(hi - n) * 8,
                               cudaMemcpyDeviceToHost));
        n = end;
    }
    TK_CUDA(cudaMemcpy(fval, tmp.data(), (size_t)count * 8, cudaMemcpyDefault));
    return TIKTAK_OK;
}

/* the k=0 record [seqNo, 0, 0, fval(p_init), p_init] (TiktakGlobalSearch.f90:496-501) */
static int push_init_row(tiktak_handle* h) {
    /* on the device, without a host round trip; the host mirror of row 0 is read back only if push_results needs it */
    const int K = h->cfg.maxpoints_plus1;
    std::fill(h->h_res_valid.begin(), h->h_res_valid.end(), 0);
    TK_CUDA(cudaMemsetAsync(h->d_res_valid, 0, (size_t)(K + 1) * sizeof(int), h->stream));
    int rc = tk_launch_init_results(h);
    if (rc) return rc;
    h->n_rows = 1;
    h->row0_on_device_only = true;
    h->best_dirty = false;
    return TIKTAK_OK;
}

static int export_starts(tiktak_handle* h, double* x_starts, int32_t* sobol_idx) {
    const int nx = h->cfg.nx, K = h->cfg.maxpoints_plus1;
    if (x_starts) TK_CUDA(cudaMemcpyAsync(x_starts, h->d_xstarts, (size_t)K * (nx + 1) * 8, cudaMemcpyDefault, h->stream));
    std::vector<unsigned int> si(std::max<int64_t>(h->n_cand, 1));
    if (sobol_idx && h->n_cand > 0)
        TK_CUDA(cudaMemcpyAsync(si.data(), h->d_sorted_i, (size_t)h->n_cand * 4, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    if (sobol_idx) {
        std::vector<int32_t> idx(K, 0);
        for (int r = 0; r < K - 1 && r < h->n_cand; r++) idx[r + 1] = (int32_t)si[r];
        TK_CUDA(cudaMemcpy(sobol_idx, idx.data(), (size_t)K * 4, cudaMemcpyDefault));
    }
    return TIKTAK_OK;
}

int tiktak_cuda_choose(tiktak_handle* h, double* x_starts, int32_t* sobol_idx) {
    if (!h) return TIKTAK_ERR_ARG;
    if (!h->pretest_done) { tk_set_error("choose: pretest has not completed"); return TIKTAK_ERR_STATE; }
    TK_CUDA(cudaSetDevice(h->device));
    int rc = begin_timer(h); /* the valid count (legitSobol.dat) is not needed to select: tiktak_cuda_pretest_result reads it */
    if (rc) return rc;
    rc = tk_select_top(h, h->cfg.maxpoints_plus1 - 1);
    if (rc) return rc;
    if (h->cfg.world > 1 && tk_comm_ready(h)) {
        /* every shard's K candidate rows [fval, idx] (+ trailer [valid count, rows filled]) in ONE all-gather; the merged
         * list is built identically on every rank (chooseSobol :840-899 over the union of the shards) */
        const int K = h->cfg.maxpoints_plus1, W = h->cfg.world;
        if ((rc = comm_buffers(h, (size_t)2 * K))) return rc;
        if ((rc = tk_export_candidates(h, h->d_xsend))) return rc;
        if ((rc = tk_comm_allgather_f64(h, h->d_xsend, h->d_xrecv, (size_t)2 * K))) return rc;
        unsigned int tot = 0;
        if (K > 1) {
            if ((rc = tk_merge_candidates(h, h->d_xrecv))) return rc;
            TK_CUDA(cudaMemcpyAsync(&tot, h->d_cand_n, sizeof tot, cudaMemcpyDeviceToHost, h->stream));
        }
        std::vector<double> trailer((size_t)W);
        for (int r = 0; r < W; r++) /* the shards' valid counts ride in the trailer rows */
            TK_CUDA(cudaMemcpyAsync(&trailer[r], h->d_xrecv + ((size_t)r * K + (K - 1)) * 2, 8, cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream));
        h->n_cand = tot;
        if (h->nlegit < 0) { double sum = 0; for (double v : trailer) sum += v; h->nlegit = (int64_t)sum; }
    }
    rc = tk_build_starts(h);
    if (rc) return rc;
    rc = end_timer(h, &h->ms_choose);
    if (rc) return rc;
    h->starts_ready = true;
    rc = push_init_row(h);
    if (rc) return rc;
    return export_starts(h, x_starts, sobol_idx);
}

int tiktak_cuda_choose_candidates(tiktak_handle* h, double* cand) {
    if (!h || !cand) return TIKTAK_ERR_ARG;
    if (!h->starts_ready) return TIKTAK_ERR_STATE;
    TK_CUDA(cudaSetDevice(h->device));
    const size_t bytes = (size_t)2 * h->cfg.maxpoints_plus1 * 8; /* (M+1) rows of 2 doubles */
    if (is_device_ptr(cand)) return tk_export_candidates(h, cand); /* stays on the stream: no host round trip */
    double* d = nullptr;
    TK_CUDA(cudaMalloc(&d, bytes));
    int rc = tk_export_candidates(h, d);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(cand, d, bytes, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { tk_set_error("choose_candidates copy: %s", cudaGetErrorString(e)); rc = TIKTAK_ERR_CUDA; }
    }
    cudaFree(d);
    return rc;
}

int tiktak_cuda_choose_merge(tiktak_handle* h, const double* cand, int64_t nrows, double* x_starts, int32_t* sobol_idx) {
    if (!h || !cand) return TIKTAK_ERR_ARG;
    const int M = h->cfg.maxpoints_plus1 - 1;
    if (nrows != (int64_t)(M + 1) * h->cfg.world) { tk_set_error("choose_merge: expected world*(K) rows"); return TIKTAK_ERR_ARG; }
    TK_CUDA(cudaSetDevice(h->device));
    int rc = TIKTAK_OK;
    double* d = nullptr;
    const double* src = cand;
    if (!is_device_ptr(cand)) {
        TK_CUDA(cudaMalloc(&d, (size_t)nrows * 16));
        TK_CUDA(cudaMemcpyAsync(d, cand, (size_t)nrows * 16, cudaMemcpyHostToDevice, h->stream));
        src = d;
    }
    unsigned int tot = 0;
    if (M > 0) {
        rc = tk_merge_candidates(h, src);
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(&tot, h->d_cand_n, sizeof tot, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { tk_set_error("choose_merge: %s", cudaGetErrorString(e)); rc = TIKTAK_ERR_CUDA; }
        }
    }
    if (d) cudaFree(d);
    if (rc) return rc;
    h->n_cand = tot;
    rc = tk_build_starts(h);
    if (rc) return rc;
    h->starts_ready = true;
    rc = push_init_row(h);
    if (rc) return rc;
    return export_starts(h, x_starts, sobol_idx);
}

int tiktak_cuda_push_results(tiktak_handle* h, const double* rows, int32_t n) {
    if (!h || !rows || n < 0) return TIKTAK_ERR_ARG;
    const int nx = h->cfg.nx, K = h->cfg.maxpoints_plus1;
    TK_CUDA(cudaSetDevice(h->device));
    std::vector<double> host((size_t)n * (nx + 4));
    if (n == 0) return TIKTAK_OK;
    if (h->row0_on_device_only) { /* the #improvements bookkeeping below needs the k = 0 value on the host */
        double f0 = 0;
        TK_CUDA(cudaMemcpyAsync(&f0, h->d_res, 8, cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream));
        h->h_res_f[0] = f0; h->h_res_valid[0] = 1; h->h_res_imp[0] = 0;
        h->row0_on_device_only = false;
    }
    h->best_dirty = true;
    TK_CUDA(cudaMemcpy(host.data(), rows, host.size() * 8, cudaMemcpyDefault));
    /* running best, for the #improvements column (completeSearch :585-589) */
    double best_f = 1.0e7; /* pos_miss, getBestLine :791-796 */
    int best_imp = 0;
    bool any = false;
    for (int k = 0; k <= K; k++)
        if (h->h_res_valid[k] && (!any || h->h_res_f[k] < best_f)) { best_f = h->h_res_f[k]; best_imp = h->h_res_imp[k]; any = true; }
    /* rows usually arrive with consecutive k: stage runs of consecutive rows and copy each run once */
    std::vector<double> stage;
    std::vector<int> ones;
    int run_k0 = -1, run_len = 0;
    auto flush = [&]() -> int {
        if (run_len == 0) return TIKTAK_OK;
        ones.assign(run_len, 1);
        TK_CUDA(cudaMemcpyAsync(h->d_res + (size_t)run_k0 * (nx + 1), stage.data(), stage.size() * 8, cudaMemcpyHostToDevice, h->stream));
        TK_CUDA(cudaMemcpyAsync(h->d_res_valid + run_k0, ones.data(), (size_t)run_len * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream)); /* staging buffers are reused */
        stage.clear(); run_len = 0; run_k0 = -1;
        return TIKTAK_OK;
    };
    for (int r = 0; r < n; r++) {
        const double fv = host[(size_t)3 * n + r];
        const int k = (int)host[(size_t)1 * n + r];
        if (!(fv == fv) || k < 0 || k > K) continue;
        int imp = best_imp;
        if (k > 0 && fv < best_f) imp = best_imp + 1;
        if (!any || fv < best_f) { best_f = fv; best_imp = imp; any = true; }
        if (!h->h_res_valid[k]) h->n_rows++;
        h->h_res_f[k] = fv; h->h_res_valid[k] = 1; h->h_res_imp[k] = imp;
        if (run_len > 0 && k != run_k0 + run_len) { int rc = flush(); if (rc) return rc; }
        if (run_len == 0) run_k0 = k;
        stage.push_back(fv);
        for (int d = 0; d < nx; d++) stage.push_back(host[(size_t)(4 + d) * n + r]);
        run_len++;
    }
    return flush();
}

/* completeSearch CASE (2) (TiktakGlobalSearch.f90:575-576): EST_dfpmin from the blended start of every restart this rank owns.
 * The CLI's algorithm `d`; the BFGS recurrences are host arithmetic in the reference's order (minimize.f90:2572-2816), every
 * objective evaluation goes through the device plugin -- the same code as lastSearch.  Restarts run one after the other: this
 * is the reference's rarely used third option, provided for completeness, not a throughput path. */
static int local_dfpmin(tiktak_handle* h, int n_k) {
    const int nx = h->cfg.nx;
    std::vector<double> st((size_t)n_k * nx), ox((size_t)n_k * nx, 0.0), of(n_k, NAN);
    std::vector<int> oe(n_k, 0);
    TK_CUDA(cudaMemcpyAsync(st.data(), h->d_starts, st.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    for (int s = h->cfg.rank; s < n_k; s += h->cfg.world) {
        std::vector<double> p(nx);
        for (int d = 0; d < nx; d++) p[d] = st[(size_t)d * n_k + s];
        double fret = 0, fn = 0;
        int its = 0;
        long nfev = 0;
        int rc = tk_est_dfpmin(h, p, &fret, &its, &nfev);
        if (rc) return rc;
        if ((rc = tiktak_cuda_objfun(h, p.data(), 1, &fn))) return rc; /* :582 fn_val = objFun(evalParam) */
        for (int d = 0; d < nx; d++) ox[(size_t)d * n_k + s] = p[d];
        of[s] = fn;
        oe[s] = (int)nfev + 1;
    }
    TK_CUDA(cudaMemcpyAsync(h->d_out_x, ox.data(), ox.size() * 8, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaMemcpyAsync(h->d_out_f, of.data(), of.size() * 8, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaMemcpyAsync(h->d_out_evals, oe.data(), oe.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream)); /* the staging vectors are locals */
    return TIKTAK_OK;
}

int tiktak_cuda_local_enqueue(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, double* send) {
    if (!h || first_k < 1 || n_k < 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (!h->starts_ready) { tk_set_error("local: choose has not run"); return TIKTAK_ERR_STATE; }
    if (alg != TIKTAK_ALG_AMOEBA && alg != TIKTAK_ALG_DFNLS && alg != TIKTAK_ALG_DFPMIN) { tk_set_error("local: unknown search method %d", alg); return TIKTAK_ERR_ARG; }
    if (n_k == 0) return TIKTAK_OK;
    if (h->cfg.world > 1 && !send) { tk_set_error("local: world > 1 needs the send buffer of the round's all-gather"); return TIKTAK_ERR_ARG; }
    TK_CUDA(cudaSetDevice(h->device));
    h->async_rows = false; /* rows in restart order from here on */
    int rc;
    if (!h->local_timing_open) {
        TK_CUDA(cudaEventRecord(h->evl0, h->stream));
        h->local_timing_open = true;
    }
    if (h->best_dirty) { /* rows were pushed from the host: recompute the best row once */
        if ((rc = tk_launch_best(h))) return rc;
        h->best_dirty = false;
    }
    if ((rc = h->cfg.search_type == 1 ? tk_launch_blend_lottery(h, first_k, n_k, h->n_rows, h->blend_w_idx) : tk_launch_blend(h, first_k, n_k, h->blend_w_idx))) return rc;
    if (alg == TIKTAK_ALG_AMOEBA && h->cfg.nm_shrink_mode == 1 && (rc = tk_launch_shrink_width(h))) return rc;
    if (alg == TIKTAK_ALG_DFPMIN) rc = local_dfpmin(h, n_k);
    else rc = (alg == TIKTAK_ALG_AMOEBA) ? tk_launch_nm(h, first_k, n_k) : tk_launch_dfnls(h, first_k, n_k);
    if (rc) return rc;
    if (h->cfg.world > 1 && (rc = tk_launch_pack(h, n_k, send))) return rc;
    TK_CUDA(cudaEventRecord(h->evl1, h->stream));
    return TIKTAK_OK;
}

int tiktak_cuda_local_ingest(tiktak_handle* h, int32_t first_k, int32_t n_k, const double* recv) {
    if (!h || first_k < 1 || n_k < 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (h->cfg.world > 1 && !recv) { tk_set_error("local_ingest: world > 1 needs the gathered rows"); return TIKTAK_ERR_ARG; }
    if (n_k == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    int rc = tk_launch_ingest(h, first_k, n_k, n_k, h->cfg.world > 1 ? recv : nullptr, nullptr);
    if (rc) return rc;
    h->n_rows += n_k; /* upper bound (rows with a NaN value are skipped on the device); exact after ingest_rounds / local_launch */
    TK_CUDA(cudaEventRecord(h->evl1, h->stream));
    return TIKTAK_OK;
}

int tiktak_cuda_local_ingest_rounds(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t round_size, const double* recv,
                                    int32_t* accepted) {
    if (!h || !accepted || first_k < 1 || n_k < 0 || round_size < 1 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (h->cfg.world > 1 && !recv) { tk_set_error("local_ingest_rounds: world > 1 needs the gathered rows"); return TIKTAK_ERR_ARG; }
    *accepted = 0;
    if (n_k == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    if (!h->d_accepted) TK_CUDA(cudaMalloc(&h->d_accepted, sizeof(int)));
    int rc = tk_launch_ingest(h, first_k, n_k, round_size, h->cfg.world > 1 ? recv : nullptr, h->d_accepted);
    if (rc) return rc;
    TK_CUDA(cudaEventRecord(h->evl1, h->stream));
    int acc = 0;
    double rows = 0; /* the row count comes from the device: restarts whose value is NaN are not rows (ingest_kernel skips them) */
    TK_CUDA(cudaMemcpyAsync(&acc, h->d_accepted, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaMemcpyAsync(&rows, h->d_best + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    h->n_rows = (int)rows;
    *accepted = acc;
    return TIKTAK_OK;
}

int tiktak_cuda_local_launch(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, int32_t round_size, int32_t* accepted) {
    if (!h || !accepted || first_k < 1 || n_k < 0 || round_size < 1 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    *accepted = 0;
    if (n_k == 0) return TIKTAK_OK;
    if (h->cfg.world > 1) {
        /* slot s of the launch is run by rank s % world; one all-gather of the ranks' rows [fval, x, #evaluations], then every
         * rank ingests the same rows in the same order (replicated results table and best row) */
        if (!tk_comm_ready(h)) { tk_set_error("local_launch: world > 1 needs tiktak_cuda_comm_init (or enqueue / all-gather / ingest_rounds by the caller)"); return TIKTAK_ERR_STATE; }
        const size_t cnt = (size_t)((n_k + h->cfg.world - 1) / h->cfg.world) * (h->cfg.nx + 2);
        int rc = comm_buffers(h, std::max(cnt, (size_t)2 * h->cfg.maxpoints_plus1));
        if (rc) return rc;
        if ((rc = tiktak_cuda_local_enqueue(h, alg, first_k, n_k, h->d_xsend))) return rc;
        if ((rc = tk_comm_allgather_f64(h, h->d_xsend, h->d_xrecv, cnt))) return rc;
        return tiktak_cuda_local_ingest_rounds(h, first_k, n_k, round_size, h->d_xrecv, accepted);
    }
    const bool fused = alg == TIKTAK_ALG_AMOEBA && h->cfg.objective_id != TIKTAK_OBJ_SMM && tk_nm_use_w4(h);
    if (!fused) { /* DFNLS, or a configuration only the latency form is built for: kernel, then the ingest kernel */
        int rc = tiktak_cuda_local_enqueue(h, alg, first_k, n_k, nullptr);
        if (rc) return rc;
        return tiktak_cuda_local_ingest_rounds(h, first_k, n_k, round_size, nullptr, accepted);
    }
    if (!h->starts_ready) { tk_set_error("local: choose has not run"); return TIKTAK_ERR_STATE; }
    TK_CUDA(cudaSetDevice(h->device));
    h->async_rows = false;
    int rc;
    if (!h->local_timing_open) {
        TK_CUDA(cudaEventRecord(h->evl0, h->stream));
        h->local_timing_open = true;
    }
    if (h->best_dirty) {
        if ((rc = tk_launch_best(h))) return rc;
        h->best_dirty = false;
    }
    if ((rc = h->cfg.search_type == 1 ? tk_launch_blend_lottery(h, first_k, n_k, h->n_rows, h->blend_w_idx) : tk_launch_blend(h, first_k, n_k, h->blend_w_idx))) return rc;
    if (h->cfg.nm_shrink_mode == 1 && (rc = tk_launch_shrink_width(h))) return rc;
    if ((rc = tk_launch_nm(h, first_k, n_k, round_size))) return rc;
    TK_CUDA(cudaEventRecord(h->evl1, h->stream));
    int acc = 0;
    double rows = 0;
    TK_CUDA(cudaMemcpyAsync(&acc, h->d_nmq + 4 /* NMQ_ACCEPTED */, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaMemcpyAsync(&rows, h->d_best + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    if (acc <= 0 || acc > n_k) { tk_set_error("local_launch: the kernel committed %d of %d restarts", acc, n_k); return TIKTAK_ERR_STATE; }
    h->n_rows = (int)rows;
    *accepted = acc;
    return TIKTAK_OK;
}

/* Asynchronous instances: restarts first_k .. first_k + n_k - 1 run by `instances` blocks that take the next restart number the
 * moment they finish one, each new start blended toward the best row finished by then -- the reference's own multi-process mode
 * (TiktakGlobalSearch.f90:546-600), made reproducible by a virtual clock (kernels_nm.cu, nm_quad_async_kernel).
 * One launch, no rounds, no host round trip; *used = instances actually run (<= the number of resident blocks). */
int tiktak_cuda_local_async(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, int32_t instances, int32_t* used) {
    if (!h || !used || first_k < 1 || n_k < 0 || instances == 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    *used = 0;
    const bool free_run = instances < 0; /* -n: n free-running instances (no virtual clock: not reproducible, see the header) */
    if (free_run) instances = -instances;
    if (free_run && alg != TIKTAK_ALG_AMOEBA) { tk_set_error("local_async: free-running instances are built for Nelder-Mead"); return TIKTAK_ERR_UNSUPPORTED; }
    if (alg != TIKTAK_ALG_AMOEBA && alg != TIKTAK_ALG_DFNLS) { tk_set_error("local_async: Nelder-Mead or DFNLS"); return TIKTAK_ERR_UNSUPPORTED; }
    if (!h->starts_ready) { tk_set_error("local: choose has not run"); return TIKTAK_ERR_STATE; }
    if (n_k == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    if (!h->local_timing_open) {
        TK_CUDA(cudaEventRecord(h->evl0, h->stream));
        h->local_timing_open = true;
    }
    int u = 0;
    int rc = alg == TIKTAK_ALG_AMOEBA ? tk_launch_nm_async(h, first_k, n_k, instances, &u, free_run) : tk_launch_dfnls(h, first_k, n_k, instances, &u);
    if (rc) return rc;
    TK_CUDA(cudaEventRecord(h->evl1, h->stream));
    h->async_rows = true;
    h->best_dirty = false;
    h->n_rows = first_k + n_k; /* (rows with a NaN value are not rows; the exact count is d_best[2]) */
    *used = u;
    return TIKTAK_OK;
}

/* what an asynchronous run did with restart k: the row its start was blended toward (-1: none, the raw Sobol start) and its place in
 * the order of the commits / virtual finish times (the order of the file; ties of the virtual clock share a value) */
int tiktak_cuda_local_async_info(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t* base, int32_t* order) {
    if (!h || first_k < 1 || n_k < 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (!h->async_rows || !h->d_async_base) { tk_set_error("local_async_info: the rows are not those of an asynchronous run"); return TIKTAK_ERR_STATE; }
    if (n_k == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    if (base) TK_CUDA(cudaMemcpy(base, h->d_async_base + first_k, (size_t)n_k * sizeof(int), cudaMemcpyDefault));
    if (order) {
        std::vector<TkAsyncRow> ev(n_k);
        TK_CUDA(cudaMemcpy(ev.data(), h->d_async_ev + first_k, (size_t)n_k * sizeof(TkAsyncRow), cudaMemcpyDeviceToHost));
        std::vector<int> o(n_k);
        for (int i = 0; i < n_k; i++) o[i] = ev[i].fin;
        TK_CUDA(cudaMemcpy(order, o.data(), (size_t)n_k * sizeof(int), cudaMemcpyDefault));
    }
    return TIKTAK_OK;
}

/* experiment hook (TIKTAK_ASYNC_DEBUG=1): per instance wait / scan / run cycles, restarts, polls of the last asynchronous launch */
extern "C" int tiktak_cuda_debug_async(tiktak_handle* h, long long* out, int n) {
    if (!h || !h->async_dbg) return TIKTAK_ERR_STATE;
    cudaStreamSynchronize(h->stream);
    return (int)cudaMemcpy(out, h->async_dbg, (size_t)n * 5 * sizeof(long long), cudaMemcpyDeviceToHost);
}

/* which instances tiktak_cuda_local_async would run: 0 = the configuration is not built for it */
int tiktak_cuda_local_async_supported(tiktak_handle* h, int32_t alg, int32_t* yes) {
    if (!h || !yes) return TIKTAK_ERR_ARG;
    *yes = ((alg == TIKTAK_ALG_AMOEBA && tk_nm_async_supported(h)) ||
            (alg == TIKTAK_ALG_DFNLS && h->cfg.world == 1 && h->cfg.search_type == 0 && h->cfg.nx <= 100)) ? 1 : 0;
    return TIKTAK_OK;
}

/* the virtual clock tiktak_cuda_local_async(h, alg, ., ., instances, .) runs on (what a model of the run has to charge a restart) */
int tiktak_cuda_local_async_clock(tiktak_handle* h, int32_t alg, int32_t instances, int32_t* clock) {
    if (!h || !clock) return TIKTAK_ERR_ARG;
    int32_t yes = 0;
    int rc = tiktak_cuda_local_async_supported(h, alg, &yes);
    if (rc) return rc;
    if (!yes) { tk_set_error("local_async_clock: asynchronous instances are not built for this configuration"); return TIKTAK_ERR_UNSUPPORTED; }
    *clock = (alg == TIKTAK_ALG_DFNLS || tk_nm_async_form(h, instances) == 1) ? TIKTAK_ASYNC_CLOCK_EVALS : TIKTAK_ASYNC_CLOCK_TRIPS;
    return TIKTAK_OK;
}

/* SolveMissingSearch (TiktakGlobalSearch.f90:1104-1146): a restart that an earlier (killed) instance never reported.  Its start
 * is getModifiedParam(p_maxpoints, x_starts(k,2:), ...) (:1121) -- the blend weight of the LAST restart number, toward the best
 * row known now -- then completeSearch(alg, k, ...) as usual (the DFNLS radii follow k).  One restart per call, like the
 * reference's sweep; every rank of a multi-GPU run makes the call (rank 0 runs the search, the row is exchanged).
 * start (nx), row (nx+4 = seqNo, k, #improvements, fval, x) and evals may be NULL. */
int tiktak_cuda_local_missing(tiktak_handle* h, int32_t alg, int32_t k, double* start, double* row, int32_t* evals) {
    if (!h || k < 1 || k > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    const int nx = h->cfg.nx, K = h->cfg.maxpoints_plus1;
    /* the best row BEFORE the search, for the #improvements column (completeSearch :585-589 sees every row of the file) */
    std::vector<double> hres((size_t)(K + 1) * (nx + 1));
    std::vector<int> hval(K + 1);
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaMemcpyAsync(hres.data(), h->d_res, hres.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaMemcpyAsync(hval.data(), h->d_res_valid, hval.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    double best_f = 1.0e7; int best_imp = 0; bool any = false;
    for (int q = 0; q <= K; q++)
        if (q != k && hval[q] && (!any || hres[(size_t)q * (nx + 1)] < best_f)) { best_f = hres[(size_t)q * (nx + 1)]; best_imp = h->h_res_imp[q]; any = true; }
    h->blend_w_idx = K;
    int32_t acc = 0;
    int rc = tiktak_cuda_local_launch(h, alg, k, 1, 1, &acc);
    h->blend_w_idx = -1;
    if (rc) return rc;
    if (acc != 1) { tk_set_error("local_missing: restart %d was not committed", k); return TIKTAK_ERR_STATE; }
    std::vector<double> r1(nx + 1), st(nx);
    int ev = 0;
    TK_CUDA(cudaMemcpyAsync(r1.data(), h->d_res + (size_t)k * (nx + 1), (size_t)(nx + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaMemcpyAsync(&ev, h->d_arch_evals + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    for (int d = 0; d < nx; d++) TK_CUDA(cudaMemcpyAsync(&st[d], h->d_arch_starts + (size_t)d * K + (k - 1), 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    const int imp = best_imp + ((any && r1[0] < best_f) ? 1 : 0);
    h->h_res_f[k] = r1[0]; h->h_res_valid[k] = 1; h->h_res_imp[k] = imp;
    if (start) for (int d = 0; d < nx; d++) start[d] = st[d];
    if (row) { row[0] = 1.0; row[1] = (double)k; row[2] = (double)imp; row[3] = r1[0]; for (int d = 0; d < nx; d++) row[4 + d] = r1[1 + d]; }
    if (evals) *evals = ev & 0xffffff;
    return TIKTAK_OK;
}

int tiktak_cuda_local_capacity(tiktak_handle* h, int32_t alg, int32_t* restarts) {
    if (!h || !restarts) return TIKTAK_ERR_ARG;
    TK_CUDA(cudaSetDevice(h->device));
    /* later rounds only depend on earlier ones through the best row, so they can be started early when the search
     * mode makes the start a function of that row alone (not of the 10 best rows / the rank lottery) */
    const bool speculate = h->cfg.search_type == 0 && !(alg == TIKTAK_ALG_AMOEBA && h->cfg.nm_shrink_mode == 1);
    int cap = tk_nm_wave_capacity(h) * h->cfg.world;
    /* nm_w4_kernel bounds the early starts itself (restart queue + in-kernel commit): a launch may hold all that is left */
    if (alg == TIKTAK_ALG_AMOEBA && h->cfg.world == 1 && h->cfg.objective_id != TIKTAK_OBJ_SMM && tk_nm_use_w4(h)) cap = std::max(cap, (int)h->cfg.maxpoints_plus1);
    *restarts = speculate ? cap : 0;
    return TIKTAK_OK;
}

int tiktak_cuda_local_fetch(tiktak_handle* h, int32_t first_k, int32_t n_k, double* starts, double* results, int32_t* evals) {
    if (!h || first_k < 1 || n_k < 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (n_k == 0) return TIKTAK_OK;
    const int nx = h->cfg.nx, K = h->cfg.maxpoints_plus1, last = first_k + n_k - 1;
    TK_CUDA(cudaSetDevice(h->device));
    /* one pinned staging area per handle: the four tables (and the rows of an asynchronous run) come over in one go, asynchronously
     * (pageable destinations made every copy a staged, synchronous one: 4.5 ms of C3's restart stage) */
    const size_t n_res = (size_t)(K + 1) * (nx + 1), n_st = (size_t)K * nx, n_i = (size_t)K + 1;
    if (!h->h_fetch) TK_CUDA(cudaMallocHost(&h->h_fetch, (n_res + n_st) * sizeof(double) + 2 * n_i * sizeof(int) + n_i * sizeof(TkAsyncRow) + 64));
    double* hres = (double*)h->h_fetch;
    double* hst = hres + n_res;
    TkAsyncRow* hrow = (TkAsyncRow*)(hst + n_st + (n_st & 1)); /* 16-byte aligned */
    int* hval = (int*)(hrow + n_i);
    int* hev = hval + n_i;
    int* herr = hev + n_i;
    TK_CUDA(cudaMemcpyAsync(hres, h->d_res, (size_t)(last + 1) * (nx + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaMemcpyAsync(hval, h->d_res_valid, (size_t)(last + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaMemcpyAsync(hev, h->d_arch_evals, (size_t)(last + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (starts) TK_CUDA(cudaMemcpyAsync(hst, h->d_arch_starts, n_st * 8, cudaMemcpyDeviceToHost, h->stream));
    if (h->async_rows) {
        TK_CUDA(cudaMemcpyAsync(herr, h->d_async_clk + 60 * 1024, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaMemcpyAsync(hrow, h->d_async_ev, (size_t)(last + 1) * sizeof(TkAsyncRow), cudaMemcpyDeviceToHost, h->stream));
    }
    TK_CUDA(cudaStreamSynchronize(h->stream));
    /* the #improvements column (completeSearch :585-589): previous best row's counter, +1 when strictly lower;
     * rows are taken in restart order k = 0, 1, 2, ... (ties: the earlier row stays the best) */
    double best_f = 1.0e7; /* pos_miss, getBestLine :791-796 */
    int best_imp = 0;
    bool any = false;
    /* caller buffers: host memory (the usual case: written in place) or device memory (staged) */
    auto put = [&](void* dst, const void* src, size_t bytes) -> int {
        if (!is_device_ptr(dst)) { memcpy(dst, src, bytes); return TIKTAK_OK; }
        TK_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
        return TIKTAK_OK;
    };
    const bool res_host = results && !is_device_ptr(results), st_host = starts && !is_device_ptr(starts);
    std::vector<double> rows_v, st_v;
    if (!res_host) rows_v.resize((size_t)n_k * (nx + 4));
    if (starts && !st_host) st_v.resize((size_t)n_k * nx);
    double* rows = res_host ? results : rows_v.data();
    double* st = st_host ? starts : st_v.data();
    std::vector<int> order(last + 1), impv(last + 1, 0);
    for (int k = 0; k <= last; k++) order[k] = k;
    if (h->async_rows) { /* the file of an asynchronous run is in the order of the (virtual) finish times, ties by instance */
        if (*herr) { tk_set_error("local_async: an instance gave up waiting for the others (the launch was not resident as a whole?)"); return TIKTAK_ERR_STATE; }
        std::vector<int> fin(last + 1), slt(last + 1);
        for (int k = 0; k <= last; k++) { fin[k] = hrow[k].fin; slt[k] = hrow[k].slt; }
        std::stable_sort(order.begin() + 1, order.end(), [&](int x, int y) {
            const int fx = std::max(fin[x], 0), fy = std::max(fin[y], 0);
            if (fx != fy) return fx < fy;
            if (slt[x] != slt[y]) return slt[x] < slt[y];
            return x < y;
        });
    }
    for (int kk = 0; kk <= last; kk++) { /* the counters, along the file */
        const int k = order[kk];
        if (hval[k] == 0) continue;
        const double fv = hres[(size_t)k * (nx + 1)];
        int imp = best_imp;
        if (k > 0 && fv < best_f) imp = best_imp + 1;
        if (!any || fv < best_f) { best_f = fv; best_imp = imp; any = true; }
        h->h_res_f[k] = fv; h->h_res_valid[k] = 1; h->h_res_imp[k] = imp;
        impv[k] = imp;
    }
    /* the rows, (n_k, nx+4) column-major: a block of restarts at a time, so that the transposition reads from cache and writes runs */
    constexpr int KB = 128;
    for (int s0 = 0; s0 < n_k; s0 += KB) {
        const int s1 = std::min(n_k, s0 + KB);
        for (int s_ = s0; s_ < s1; s_++) {
            const int k = first_k + s_;
            const bool valid = hval[k] != 0;
            rows[(size_t)0 * n_k + s_] = 1.0; /* seqNo */
            rows[(size_t)1 * n_k + s_] = (double)k;
            rows[(size_t)2 * n_k + s_] = valid ? (double)impv[k] : 0.0;
            rows[(size_t)3 * n_k + s_] = valid ? hres[(size_t)k * (nx + 1)] : NAN;
        }
        for (int d = 0; d < nx; d++) {
            double* dst = rows + (size_t)(4 + d) * n_k;
            for (int s_ = s0; s_ < s1; s_++) {
                const int k = first_k + s_;
                dst[s_] = hval[k] != 0 ? hres[(size_t)k * (nx + 1) + 1 + d] : NAN;
            }
        }
    }
    if (starts) /* (K, nx) column-major on the device, (n_k, nx) column-major for the caller: a run per coordinate */
        for (int d = 0; d < nx; d++) memcpy(st + (size_t)d * n_k, hst + (size_t)d * K + (first_k - 1), (size_t)n_k * 8);
    int rcp = TIKTAK_OK;
    if (starts && !st_host && (rcp = put(starts, st_v.data(), st_v.size() * 8))) return rcp;
    if (results && !res_host && (rcp = put(results, rows_v.data(), rows_v.size() * 8))) return rcp;
    if (evals) {
        std::vector<int> ev2(n_k);
        for (int s = 0; s < n_k; s++) ev2[s] = hev[first_k + s] & 0xffffff; /* the high byte is the local search's status (tiktak_cuda_local_status) */
        if ((rcp = put(evals, ev2.data(), (size_t)n_k * sizeof(int)))) return rcp;
    }
    return TIKTAK_OK;
}

int tiktak_cuda_local_status(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t* status) {
    if (!h || !status || first_k < 1 || n_k < 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (n_k == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    std::vector<int> hev(n_k);
    TK_CUDA(cudaMemcpyAsync(hev.data(), h->d_arch_evals + first_k, (size_t)n_k * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    for (int& e : hev) e = (e >> 24) & 0x3;
    TK_CUDA(cudaMemcpy(status, hev.data(), (size_t)n_k * sizeof(int), cudaMemcpyDefault));
    return TIKTAK_OK;
}

/* how often restart k called RESCUE_H (minimize.f90:643-650; DFNLS only, saturates at 31) */
int tiktak_cuda_local_nrescue(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t* nrescue) {
    if (!h || !nrescue || first_k < 1 || n_k < 0 || first_k + n_k - 1 > h->cfg.maxpoints_plus1) return TIKTAK_ERR_ARG;
    if (n_k == 0) return TIKTAK_OK;
    TK_CUDA(cudaSetDevice(h->device));
    std::vector<int> hev(n_k);
    TK_CUDA(cudaMemcpyAsync(hev.data(), h->d_arch_evals + first_k, (size_t)n_k * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    for (int& e : hev) e = (e >> 26) & 0x1f;
    TK_CUDA(cudaMemcpy(nrescue, hev.data(), (size_t)n_k * sizeof(int), cudaMemcpyDefault));
    return TIKTAK_OK;
}

int tiktak_cuda_local(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, double* starts, double* results,
                      int32_t* evals) {
    if (!h) return TIKTAK_ERR_ARG;
    if (h->cfg.world > 1) {
        tk_set_error("local: with world > 1 drive the rounds with local_enqueue / all-gather / local_ingest");
        return TIKTAK_ERR_UNSUPPORTED;
    }
    int rc = tiktak_cuda_local_enqueue(h, alg, first_k, n_k, nullptr);
    if (rc) return rc;
    if ((rc = tiktak_cuda_local_ingest(h, first_k, n_k, nullptr))) return rc;
    return tiktak_cuda_local_fetch(h, first_k, n_k, starts, results, evals);
}

int tiktak_cuda_best(tiktak_handle* h, double* fbest, double* xbest, int32_t* k) {
    if (!h) return TIKTAK_ERR_ARG;
    const int nx = h->cfg.nx;
    TK_CUDA(cudaSetDevice(h->device));
    int rc = tk_launch_best(h);
    if (rc) return rc;
    std::vector<double> b(nx + 3);
    TK_CUDA(cudaMemcpyAsync(b.data(), h->d_best, b.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    if (b[1] < 0) { /* getBestLine with no rows (:791-796) */
        if (fbest) *fbest = 1.0e7;
        if (xbest) for (int d = 0; d < nx; d++) xbest[d] = h->init[d];
        if (k) *k = 0;
        return TIKTAK_OK;
    }
    if (fbest) *fbest = b[0];
    if (k) *k = (int32_t)b[1];
    if (xbest) for (int d = 0; d < nx; d++) xbest[d] = b[3 + d];
    return TIKTAK_OK;
}

/* lastSearch (TiktakGlobalSearch.f90:635-681) = EST_dfpmin + dograd + lnsrch (minimize.f90:2572-2816) from the
 * best row.  It runs once per job; the BFGS recurrences (O(nx^2) per iteration) are host arithmetic in the
 * reference's order, every objective evaluation goes through the device plugin: the 2 nx points of a central-
 * difference gradient as one batch, line-search trial points one by one. */
static int tk_est_dfpmin(tiktak_handle* h, std::vector<double>& p, double* fret_out, int* its_out, long* nfev) {
    const int nn = h->cfg.nx, itmax = 1000;
    const double gtol = 1.0e-6; /* itmax_DFPMIN, gtol_DFPMIN (genericParams.f90:27-28) */
    int rc = TIKTAK_OK;
    const double stpmx = 100.0, eps = 2.220446049250313e-16, tolx = 4.0 * eps, xinc = 1.0e-5;
    auto func = [&](const std::vector<double>& xx, double& out) -> int { if (nfev) (*nfev)++; return tiktak_cuda_objfun(h, xx.data(), 1, &out); };
    auto dograd = [&](const std::vector<double>& xx, std::vector<double>& grad) -> int {
        std::vector<double> dh(nn), pts((size_t)2 * nn * nn), f((size_t)2 * nn);
        const double tolera = 10.0 * xinc;
        for (int i = 0; i < nn; i++) dh[i] = (fabs(xx[i]) >= tolera) ? xx[i] * xinc : xinc;
        const int n = 2 * nn; /* point 2i: x - dh e_i, point 2i+1: x + dh e_i ; column-major (n, nn) */
        for (int i = 0; i < nn; i++)
            for (int j = 0; j < nn; j++) {
                pts[(size_t)j * n + 2 * i] = xx[j] - (i == j ? dh[i] : 0.0);
                pts[(size_t)j * n + 2 * i + 1] = xx[j] + (i == j ? dh[i] : 0.0);
            }
        int r = tiktak_cuda_objfun(h, pts.data(), n, f.data());
        if (r) return r;
        if (nfev) *nfev += n;
        grad.resize(nn);
        for (int i = 0; i < nn; i++) grad[i] = (f[2 * i + 1] - f[2 * i]) / (2.0 * dh[i]);
        return TIKTAK_OK;
    };
    auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) { double sm = 0.0; for (int i = 0; i < nn; i++) sm = sm + a[i] * b[i]; return sm; };
    double fp = 0, fret = 0;
    if ((rc = func(p, fp))) return rc;
    std::vector<double> g, dg(nn), hdg(nn), pnew(nn), xi(nn), hessin((size_t)nn * nn, 0.0);
    if ((rc = dograd(p, g))) return rc;
    for (int i = 0; i < nn; i++) hessin[(size_t)i * nn + i] = 1.0;
    for (int i = 0; i < nn; i++) xi[i] = -g[i];
    const double stpmax = stpmx * std::max(sqrt(dot(p, p)), (double)nn);
    fret = fp;
    int its_done = itmax;
    for (int its = 1; its <= itmax; its++) {
        { /* lnsrch :2719-2814 */
            const double alf = 1.0e-4;
            double a, alam, alam2 = 0, alamin, b, disc, f2 = 0, pabs, rhs1, rhs2, slope, tmplam, f;
            pabs = sqrt(dot(xi, xi));
            if (pabs > stpmax) for (int i = 0; i < nn; i++) xi[i] = xi[i] * stpmax / pabs;
            slope = dot(g, xi);
            pnew = p; f = fp;
            if (!(slope >= 0.0)) {
                double mx = 0.0;
                for (int i = 0; i < nn; i++) mx = std::max(mx, fabs(xi[i]) / std::max(fabs(p[i]), 1.0));
                alamin = eps / mx;
                alam = 1.0;
                for (;;) {
                    for (int i = 0; i < nn; i++) pnew[i] = p[i] + alam * xi[i];
                    if ((rc = func(pnew, f))) return rc;
                    if (alam < alamin) { pnew = p; break; }
                    else if (f <= fp + alf * alam * slope) break;
                    else {
                        if (alam == 1.0) tmplam = -slope / (2.0 * (f - fp - slope));
                        else {
                            rhs1 = f - fp - alam * slope;
                            rhs2 = f2 - fp - alam2 * slope;
                            a = (rhs1 / (alam * alam) - rhs2 / (alam2 * alam2)) / (alam - alam2);
                            b = (-alam2 * rhs1 / (alam * alam) + alam * rhs2 / (alam2 * alam2)) / (alam - alam2);
                            if (a == 0.0) tmplam = -slope / (2.0 * b);
                            else {
                                disc = b * b - 3.0 * a * slope;
                                if (disc < 0.0) tmplam = 0.5 * alam;
                                else if (b <= 0.0) tmplam = (-b + sqrt(disc)) / (3.0 * a);
                                else tmplam = -slope / (b + sqrt(disc));
                            }
                            if (tmplam > 0.5 * alam) tmplam = 0.5 * alam;
                        }
                    }
                    alam2 = alam;
                    f2 = f;
                    alam = std::max(tmplam, 0.1 * alam);
                }
            }
            fret = f;
        }
        fp = fret;
        for (int i = 0; i < nn; i++) xi[i] = pnew[i] - p[i];
        p = pnew;
        double t = 0.0;
        for (int i = 0; i < nn; i++) t = std::max(t, fabs(xi[i]) / std::max(fabs(p[i]), 1.0));
        if (t < tolx) { its_done = its; break; }
        dg = g;
        if ((rc = dograd(p, g))) return rc;
        const double den = std::max(fret, 1.0);
        t = 0.0;
        for (int i = 0; i < nn; i++) t = std::max(t, fabs(g[i]) * std::max(fabs(p[i]), 1.0) / den);
        if (t < gtol) { its_done = its; break; }
        for (int i = 0; i < nn; i++) dg[i] = g[i] - dg[i];
        for (int i = 0; i < nn; i++) { double sm = 0.0; for (int j = 0; j < nn; j++) sm = sm + hessin[(size_t)i * nn + j] * dg[j]; hdg[i] = sm; }
        double fac = dot(dg, xi), fae = dot(dg, hdg), sumdg = dot(dg, dg), sumxi = dot(xi, xi);
        if (fac > sqrt(eps * sumdg * sumxi)) {
            fac = 1.0 / fac;
            const double fad = 1.0 / fae;
            for (int i = 0; i < nn; i++) dg[i] = fac * xi[i] - fad * hdg[i];
            for (int i = 0; i < nn; i++)
                for (int j = 0; j < nn; j++)
                    hessin[(size_t)i * nn + j] = hessin[(size_t)i * nn + j] + fac * (xi[i] * xi[j]) - fad * (hdg[i] * hdg[j]) + fae * (dg[i] * dg[j]);
        }
        for (int i = 0; i < nn; i++) { double sm = 0.0; for (int j = 0; j < nn; j++) sm = sm + hessin[(size_t)i * nn + j] * g[j]; xi[i] = -sm; }
    }
    if (fret_out) *fret_out = fret;
    if (its_out) *its_out = its_done;
    return TIKTAK_OK;
}

int tiktak_cuda_last_search(tiktak_handle* h, double* fval, double* x, int32_t* iters) {
    if (!h) return TIKTAK_ERR_ARG;
    const int nn = h->cfg.nx;
    double fbest = 0, fret = 0;
    std::vector<double> p(nn);
    int kb = 0, its = 0;
    int rc = tiktak_cuda_best(h, &fbest, p.data(), &kb);
    if (rc) return rc;
    if ((rc = tk_est_dfpmin(h, p, &fret, &its, nullptr))) return rc;
    if (fval) *fval = fret;
    if (x) for (int i = 0; i < nn; i++) x[i] = p[i];
    if (iters) *iters = its;
    return TIKTAK_OK;
}

int tiktak_cuda_fp64_peak(int device, double* flops_per_s, double* ms) { return tk_fp64_peak(device, flops_per_s, ms); }

} /* extern "C" */
