/* kernels_select.cu -- start selection on the device.
 *
 * Replaces chooseSobol (TiktakGlobalSearch.f90:840-899): the reference re-reads every evaluated row
 * from sobolFnVal.dat, index-sorts ALL of them with indexx (utilities.f90:1334-1424) and keeps the
 * first K-1.  Only the K-1 smallest are ever used, so the device does an exact radix SELECT on the
 * order-preserving 64-bit image of fval (8 histogram passes of 8 bits over the fval array, most
 * significant digit first), resolves a tie at the threshold by Sobol index (4 more passes, skipped
 * when there is none), compacts the <= K-1 winners and orders them by (fval, index) with a
 * rank-by-counting pass.  Order among exactly equal fvals is ascending Sobol index (indexx's order
 * among equal keys is an artefact of its partitioning; DESIGN.md "ties").
 */
#include "tk_common.cuh"
#include "smm.cuh"

namespace {

__device__ __forceinline__ unsigned long long tk_order_key(double f) {
    unsigned long long u = (unsigned long long)__double_as_longlong(f);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}

/* selstate layout (unsigned long long): [0] key prefix, [1] remaining, [2] idx prefix,
 * [3] tie flag (1 = idx passes needed), [4] count of keys == threshold */
struct SelView {
    const double* fval;
    int64_t nlocal; /* local elements (multiple of the count block) */
    int64_t kcut;
    int rank, world;
};
__device__ __forceinline__ int64_t tk_global_n(const SelView& v, int64_t li) {
    return ((li / TIKTAK_SOBOL_BLOCK) * v.world + v.rank) * (int64_t)TIKTAK_SOBOL_BLOCK + (li % TIKTAK_SOBOL_BLOCK);
}

/* phase 0: digits of the key; phase 1: digits of n among keys == threshold */
__global__ void __launch_bounds__(256)
select_hist_kernel(const SelView v, const unsigned long long* state, unsigned long long* hist, int phase, int shift) {
    __shared__ unsigned int sh[256];
    if (phase == 1 && state[3] == 0) return;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = state[0], iprefix = state[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; li < v.nlocal; li += stride) {
        const int64_t n = tk_global_n(v, li);
        if (n < 1 || n > v.kcut) continue;
        const unsigned long long key = tk_order_key(v.fval[li]);
        if (phase == 0) {
            const bool match = (shift == 56) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
            if (match) atomicAdd(&sh[(key >> shift) & 255u], 1u);
        } else {
            if (key != prefix) continue;
            const unsigned long long un = (unsigned long long)n;
            const bool match = (shift == 24) || ((un >> (shift + 8)) == (iprefix >> (shift + 8)));
            if (match) atomicAdd(&sh[(un >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void select_scan_kernel(unsigned long long* state, unsigned long long* hist, int phase, int shift) {
    if (threadIdx.x != 0) return;
    if (phase == 1 && state[3] == 0) return;
    unsigned long long rem = state[1], cum = 0;
    int b = 0;
    for (; b < 256; b++) {
        if (cum + hist[b] >= rem) break;
        cum += hist[b];
    }
    if (b == 256) b = 255; /* fewer elements than requested: everything qualifies */
    const unsigned long long inbin = hist[b];
    rem -= cum;
    if (phase == 0) {
        state[0] |= ((unsigned long long)b) << shift;
        state[1] = rem;
        if (shift == 0) { /* threshold key fixed: ties iff more keys equal it than still needed */
            state[4] = inbin;
            state[3] = (inbin > rem) ? 1 : 0;
            state[2] = (inbin > rem) ? 0 : 0xffffffffULL;
        }
    } else {
        state[2] |= ((unsigned long long)b) << shift;
        state[1] = rem;
    }
    for (int i = 0; i < 256; i++) hist[i] = 0;
}

__global__ void __launch_bounds__(256)
select_collect_kernel(const SelView v, const unsigned long long* state, double* cand_f, unsigned int* cand_i,
                      unsigned int* cand_n, unsigned int cap) {
    const unsigned long long T = state[0], nT = state[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; li < v.nlocal; li += stride) {
        const int64_t n = tk_global_n(v, li);
        if (n < 1 || n > v.kcut) continue;
        const double f = v.fval[li];
        const unsigned long long key = tk_order_key(f);
        if (key < T || (key == T && (unsigned long long)n <= nT)) {
            const unsigned int pos = atomicAdd(cand_n, 1u);
            if (pos < cap) { cand_f[pos] = f; cand_i[pos] = (unsigned int)n; }
        }
    }
}

/* order m candidates by (key, n): rank = number of strictly smaller composites (all distinct) */
__global__ void __launch_bounds__(256)
rank_sort_kernel(const double* cf, const unsigned int* ci, unsigned int m, double* sf, unsigned int* si) {
    __shared__ unsigned long long tk[256];
    __shared__ unsigned int tn[256];
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mykey = 0;
    unsigned int myn = 0;
    double myf = 0;
    if (i < m) { myf = cf[i]; mykey = tk_order_key(myf); myn = ci[i]; }
    unsigned int rank = 0;
    for (unsigned int base = 0; base < m; base += 256) {
        const unsigned int j = base + threadIdx.x;
        if (j < m) { tk[threadIdx.x] = tk_order_key(cf[j]); tn[threadIdx.x] = ci[j]; }
        __syncthreads();
        const unsigned int lim = min(256u, m - base);
        if (i < m)
            for (unsigned int u = 0; u < lim; u++)
                rank += (tk[u] < mykey || (tk[u] == mykey && tn[u] < myn)) ? 1u : 0u;
        __syncthreads();
    }
    if (i < m) { sf[rank] = myf; si[rank] = myn; }
}

/* x_starts rows 2..: [fval, Sobol point n], through the f40.20 record when text_roundtrip is on
 * (sobolFnVal.dat is written :467 and re-read :860).  Column-major (K, nx+1). */
__global__ void build_starts_kernel(const TkTablesG T, int nx, int K, unsigned int m, const double* sf,
                                    const unsigned int* si, int roundtrip, double* xs) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; /* candidate r -> row r+1 */
    if (r >= (unsigned int)(K - 1)) return;
    const int row = r + 1;
    if (r >= m) { /* the reference leaves these rows uninitialised (:851, 884-892) */
        for (int c = 0; c <= nx; c++) xs[(size_t)c * K + row] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const double f = sf[r];
    xs[row] = roundtrip ? tk_f4020_roundtrip(f) : f;
    const uint32_t n = si[r], g = n ^ (n >> 1);
    for (int d = 0; d < nx; d++) {
        uint32_t q = 0;
        for (int j = 0; j < TK_MAXBIT; j++)
            if ((g >> j) & 1u) q ^= T.V[(size_t)d * 32 + j];
        const double x = __dadd_rn(T.lo[d], __dmul_rn(tk_q2unit(q), T.w[d]));
        xs[(size_t)(d + 1) * K + row] = roundtrip ? tk_f4020_roundtrip(x) : x;
    }
}

__global__ void init_row_kernel(const double* init, const double* finit, int nx, int K, double* xs) {
    /* x_starts(1,:) = [objFun(p_init), p_init] (:853-854): in-memory values, no text round trip */
    if (threadIdx.x == 0) xs[0] = finit[0];
    for (int d = threadIdx.x; d < nx; d += blockDim.x) xs[(size_t)(d + 1) * K] = init[d];
}

} /* namespace */

/* select the M smallest evaluated points of this shard into d_sorted_f / d_sorted_i (n_cand of them) */
int tk_select_top(tiktak_handle* h, int64_t M) {
    const int world = h->cfg.world, rank = h->cfg.rank;
    /* how many evaluated points does this shard own? */
    int64_t owned = 0;
    {
        const int64_t kc = h->kcut;
        const int64_t full = (kc + 1) / TIKTAK_SOBOL_BLOCK; /* count blocks completely inside [0, kc] */
        for (int64_t cb = rank; cb <= kc / TIKTAK_SOBOL_BLOCK; cb += world) {
            int64_t lo = cb * (int64_t)TIKTAK_SOBOL_BLOCK, hi = lo + TIKTAK_SOBOL_BLOCK - 1;
            if (lo < 1) lo = 1;
            if (hi > kc) hi = kc;
            if (hi >= lo) owned += hi - lo + 1;
        }
        (void)full;
    }
    const int64_t m = M < owned ? M : owned;
    h->n_cand = m;
    if (m <= 0) return TIKTAK_OK;
    SelView v{h->d_fval, h->n_local_cblocks * (int64_t)TIKTAK_SOBOL_BLOCK, h->kcut, rank, world};
    /* only scan the count blocks that can hold n <= kcut */
    {
        const int64_t lastcb = h->kcut / TIKTAK_SOBOL_BLOCK;
        const int64_t nl = (lastcb >= rank) ? ((lastcb - rank) / world + 1) : 0;
        if (nl * (int64_t)TIKTAK_SOBOL_BLOCK < v.nlocal) v.nlocal = nl * (int64_t)TIKTAK_SOBOL_BLOCK;
    }
    unsigned long long st[8] = {0, (unsigned long long)m, 0, 0, 0, 0, 0, 0};
    TK_CUDA(cudaMemcpyAsync(h->d_selstate, st, sizeof st, cudaMemcpyHostToDevice, h->stream));
    TK_CUDA(cudaMemsetAsync(h->d_hist, 0, 256 * sizeof(unsigned long long), h->stream));
    TK_CUDA(cudaMemsetAsync(h->d_cand_n, 0, sizeof(unsigned int), h->stream));
    int64_t want = (v.nlocal + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 16;
    const unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
    for (int phase = 0; phase < 2; phase++) {
        const int top = phase == 0 ? 56 : 24;
        for (int shift = top; shift >= 0; shift -= 8) {
            select_hist_kernel<<<grid, 256, 0, h->stream>>>(v, h->d_selstate, h->d_hist, phase, shift);
            select_scan_kernel<<<1, 32, 0, h->stream>>>(h->d_selstate, h->d_hist, phase, shift);
            h->launches += 2;
        }
    }
    select_collect_kernel<<<grid, 256, 0, h->stream>>>(v, h->d_selstate, h->d_cand_f, h->d_cand_i, h->d_cand_n,
                                                       (unsigned int)m);
    rank_sort_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(h->d_cand_f, h->d_cand_i, (unsigned int)m,
                                                                          h->d_sorted_f, h->d_sorted_i);
    h->launches += 2;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

/* order an externally supplied candidate list (multi-GPU merge) and keep the best M */
int tk_sort_candidates(tiktak_handle* h, const double* d_f, const unsigned int* d_i, int64_t n, double* d_sf,
                       unsigned int* d_si) {
    if (n <= 0) return TIKTAK_OK;
    rank_sort_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_f, d_i, (unsigned int)n, d_sf, d_si);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_build_starts(tiktak_handle* h) {
    const int K = h->cfg.maxpoints_plus1, nx = h->cfg.nx;
    /* row 1: objFun(p_init) evaluated by the same plugin kernel */
    int rc = (h->cfg.objective_id == TIKTAK_OBJ_SMM) ? tk_launch_smm_objfun(h, h->d_init, 1, h->d_best)
                                                      : tk_launch_objfun(h, h->d_init, 1, h->d_best); /* d_best[0] used as scratch */
    if (rc) return rc;
    init_row_kernel<<<1, 64, 0, h->stream>>>(h->d_init, h->d_best, nx, K, h->d_xstarts);
    h->launches++;
    if (K > 1) {
        const TkTablesG T{h->d_V, h->d_lo, h->d_w, h->d_bl, h->d_bu, h->d_aux};
        build_starts_kernel<<<(unsigned)((K - 1 + 127) / 128), 128, 0, h->stream>>>(
            T, nx, K, (unsigned int)h->n_cand, h->d_sorted_f, h->d_sorted_i, h->cfg.text_roundtrip, h->d_xstarts);
        h->launches++;
    }
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}
