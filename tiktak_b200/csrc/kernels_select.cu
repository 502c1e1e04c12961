/* kernels_select.cu -- start selection on the device.
 *
 * Replaces chooseSobol (TiktakGlobalSearch.f90:840-899): the reference re-reads every evaluated row
 * from sobolFnVal.dat, index-sorts ALL of them with indexx (utilities.f90:1334-1424) and keeps the
 * first K-1.  Only the K-1 smallest are ever used, so the device does an exact radix SELECT on the
 * order-preserving 64-bit image of fval (8 histogram passes of 8 bits, most significant digit first; the
 * last block of a pass scans the bins).  Only the first two passes read the fval array: once the top 16
 * bits of the threshold are known, the elements that share them -- the only ones the later digits depend
 * on -- are copied to a short (key, index) list and the remaining passes read that list (four sweeps over
 * the array per selection instead of thirteen).  A tie at the threshold is resolved by Sobol index (4 more
 * passes over the list, skipped when there is none); then the <= K-1 winners are compacted and ordered by
 * (fval, index) with a chunk sort + merge-by-rank passes.  Order among exactly equal fvals is ascending Sobol index (indexx's order
 * among equal keys is an artefact of its partitioning; DESIGN.md "ties").
 */
#include <utility>

#include "tk_common.cuh"
#include "smm.cuh"

namespace {

/* order-preserving image of a double.  Every NaN (either sign, any payload: CUDA's canonical NaN has the sign bit set, and
 * a plugin objective may return one, e.g. tk_cos outside its domain) maps to ONE key above +inf and below the padding key
 * ~0: NaN rows sort last, after every number, and are selected only when fewer than K-1 numbers exist.  (The reference's
 * indexx has no defined order for NaN; README.md:34 asks objectives to return a large finite value instead.) */
#define TK_KEY_NAN 0xFFFFFFFFFFFFFFFEULL
__device__ __forceinline__ unsigned long long tk_order_key(double f) {
    unsigned long long u = (unsigned long long)__double_as_longlong(f);
    if (f != f) return TK_KEY_NAN;
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}

/* selstate layout (unsigned long long): [0] key prefix, [1] remaining, [2] idx prefix,
 * [3] tie flag (1 = idx passes needed), [4] count of keys == threshold */
struct SelView {
    const double* fval;
    int64_t nlocal; /* local elements (multiple of the count block) */
    int64_t kcut;
    int rank, world;
    unsigned long long* list_key; /* elements whose key shares the threshold's top 16 bits (filled after pass 2) */
    unsigned int* list_n;
    unsigned long long list_cap;
};
/* selstate[5] = elements in the list, selstate[6] = 1 when they did not fit (later passes read the array again) */
__device__ __forceinline__ int64_t tk_global_n(const SelView& v, int64_t li) {
    return ((li / TIKTAK_SOBOL_BLOCK) * v.world + v.rank) * (int64_t)TIKTAK_SOBOL_BLOCK + (li % TIKTAK_SOBOL_BLOCK);
}

/* One radix pass = one kernel: every block adds its shared-memory histogram to the global one and takes a ticket;
 * the block that draws the last ticket scans the 256 bins, fixes the next digit of the threshold in `state`, and
 * clears histogram and ticket for the next pass (no separate scan launch, no host round trip).
 * phase 0: digits of the key; phase 1: digits of n among keys == threshold */
__global__ void __launch_bounds__(256)
select_hist_kernel(const SelView v, unsigned long long* state, unsigned long long* hist, unsigned int* ticket, int phase, int shift) {
    __shared__ unsigned int sh[256];
    __shared__ bool last;
    if (phase == 1 && state[3] == 0) return;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = state[0], iprefix = state[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool from_list = !(phase == 0 && shift >= 48) && state[6] == 0; /* the list exists and is complete */
    const int64_t count = from_list ? (int64_t)state[5] : v.nlocal;
    for (int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; li < count; li += stride) {
        unsigned long long key;
        int64_t n;
        if (from_list) {
            key = v.list_key[li];
            n = (int64_t)v.list_n[li];
        } else {
            n = tk_global_n(v, li);
            if (n < 1 || n > v.kcut) continue;
            key = tk_order_key(v.fval[li]);
        }
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
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ unsigned long long vh[256];
    vh[threadIdx.x] = __ldcg(&hist[threadIdx.x]); /* all bins in one parallel read from L2 */
    __syncthreads();
    if (threadIdx.x == 0) { /* 256 bins, one thread, shared memory */
        unsigned long long rem = state[1], cum = 0;
        int b = 0;
        for (; b < 256; b++) {
            if (cum + vh[b] >= rem) break;
            cum += vh[b];
        }
        if (b == 256) b = 255; /* fewer elements than requested: everything qualifies */
        const unsigned long long inbin = vh[b];
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
        *ticket = 0;
    }
    __syncthreads();
    hist[threadIdx.x] = 0;
}

/* after the pass with shift 48: copy the elements whose key has the threshold's top 16 bits to the list (one atomic per
 * warp); the order in the list is arbitrary, the passes that read it only count */
__global__ void __launch_bounds__(256)
select_compact_kernel(const SelView v, unsigned long long* state) {
    const unsigned long long top = state[0] >> 48;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (v.nlocal + stride - 1) / stride;
    const int lane = threadIdx.x & 31;
    int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = 0; r < rounds; r++, li += stride) { /* whole warps stay in the loop: the ballot below is warp-wide */
        bool take = false;
        unsigned long long key = 0;
        int64_t n = 0;
        if (li < v.nlocal) {
            n = tk_global_n(v, li);
            if (n >= 1 && n <= v.kcut) {
                key = tk_order_key(v.fval[li]);
                take = (key >> 48) == top;
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, take);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == __ffs((int)m) - 1) base = atomicAdd(&state[5], (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
        if (take) {
            const unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
            if (pos < v.list_cap) { v.list_key[pos] = key; v.list_n[pos] = (unsigned int)n; }
            else state[6] = 1; /* does not fit: the later passes fall back to the array */
        }
    }
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

/* Ordering the winners by (key, n) -- all composites are distinct because every n occurs once.
 * Runs of SORT_CHUNK candidates are sorted in shared memory (bitonic network), then merged pairwise by
 * rank: an element's position in the merged run is its own index plus the number of partner-run elements
 * below it (binary search), so every pass is one fully parallel kernel and the whole sort is
 * O(m log^2 m) work at any m (the multi-GPU merge of world x (K-1) candidates included). */
#define SORT_CHUNK 1024
struct Comp { unsigned long long key; unsigned int n; };
__device__ __forceinline__ bool comp_less(unsigned long long ka, unsigned int na, unsigned long long kb, unsigned int nb) {
    return ka < kb || (ka == kb && na < nb);
}
__global__ void __launch_bounds__(256)
chunk_sort_kernel(const double* cf, const unsigned int* ci, unsigned int m, double* sf, unsigned int* si) {
    __shared__ unsigned long long k[SORT_CHUNK];
    __shared__ unsigned int n[SORT_CHUNK];
    const unsigned int base = blockIdx.x * SORT_CHUNK;
    for (unsigned int t = threadIdx.x; t < SORT_CHUNK; t += blockDim.x) {
        const unsigned int e = base + t;
        /* padding sorts last: every NaN has the key TK_KEY_NAN, so the all-ones key is free */
        k[t] = (e < m) ? tk_order_key(cf[e]) : ~0ULL;
        n[t] = (e < m) ? ci[e] : 0xffffffffu;
    }
    __syncthreads();
    for (unsigned int size = 2; size <= SORT_CHUNK; size <<= 1)
        for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned int t = threadIdx.x; t < SORT_CHUNK / 2; t += blockDim.x) {
                const unsigned int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long ka = k[lo], kb = k[hi];
                const unsigned int na = n[lo], nb = n[hi];
                if (comp_less(kb, nb, ka, na) == up) { k[lo] = kb; k[hi] = ka; n[lo] = nb; n[hi] = na; }
            }
            __syncthreads();
        }
    for (unsigned int t = threadIdx.x; t < SORT_CHUNK; t += blockDim.x) {
        const unsigned int e = base + t;
        if (e < m) {
            const unsigned long long u = k[t];
            sf[e] = __longlong_as_double((long long)((u >> 63) ? (u ^ 0x8000000000000000ULL) : ~u));
            si[e] = n[t];
        }
    }
}
/* merge neighbouring sorted runs of length `run` (the last one may be short) */
__global__ void __launch_bounds__(256)
merge_pass_kernel(const double* inf, const unsigned int* ini, unsigned int m, unsigned int run, double* outf,
                  unsigned int* outi) {
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const unsigned int pair = e / (2 * run), base = pair * 2 * run;
    const bool inA = (e - base) < run;
    const unsigned int ob = inA ? base + run : base;               /* partner run */
    const unsigned int oe = inA ? min(base + 2 * run, m) : base + run;
    const double f = inf[e];
    const unsigned long long key = tk_order_key(f);
    const unsigned int n = ini[e];
    unsigned int lo = ob, hi = (ob < oe) ? oe : ob; /* count partner elements below (key, n) */
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (comp_less(tk_order_key(inf[mid]), ini[mid], key, n)) lo = mid + 1; else hi = mid;
    }
    const unsigned int below = lo - ob;
    const unsigned int pos = base + (inA ? (e - base) : (e - base - run)) + below;
    outf[pos] = f;
    outi[pos] = n;
}

/* multi-GPU merge: `world` candidate lists, each (M+1) rows of [fval, (double)n]: rows 0..cnt-1 sorted by
 * (fval, n), NaN padding, and a trailer row [#valid points of the shard, cnt].  Position of an element in
 * the union = its index in its own list + the number of elements below it in every other list. */
__global__ void __launch_bounds__(256)
merge_lists_kernel(const double* cand, int world, unsigned int M, double* outf, unsigned int* outi, unsigned int* total) {
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) {
        unsigned int t = 0;
        for (int r = 0; r < world; r++) t += (unsigned int)cand[((size_t)r * (M + 1) + M) * 2 + 1];
        *total = t < M ? t : M;
    }
    if (e >= (unsigned int)world * M) return;
    const int r = e / M;
    const unsigned int i = e - r * M;
    const double* mine = cand + (size_t)r * (M + 1) * 2;
    if (i >= (unsigned int)mine[(size_t)M * 2 + 1]) return;
    const double f = mine[2 * i];
    const unsigned long long key = tk_order_key(f);
    const unsigned int n = (unsigned int)mine[2 * i + 1];
    unsigned int pos = i;
    for (int q = 0; q < world; q++) {
        if (q == r) continue;
        const double* oth = cand + (size_t)q * (M + 1) * 2;
        unsigned int lo = 0, hi = (unsigned int)oth[(size_t)M * 2 + 1];
        while (lo < hi) {
            const unsigned int mid = (lo + hi) >> 1;
            if (comp_less(tk_order_key(oth[2 * mid]), (unsigned int)oth[2 * mid + 1], key, n)) lo = mid + 1; else hi = mid;
        }
        pos += lo;
    }
    if (pos < M) { outf[pos] = f; outi[pos] = n; }
}
/* this shard's list in that format */
__global__ void export_candidates_kernel(const double* sf, const unsigned int* si, unsigned int cnt, unsigned int M,
                                         const unsigned long long* counts, int64_t ncb, double* cand) {
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < M) {
        cand[2 * e] = (e < cnt) ? sf[e] : __longlong_as_double(0x7ff8000000000000LL);
        cand[2 * e + 1] = (e < cnt) ? (double)si[e] : 0.0;
    } else if (e == M) {
        unsigned long long v = 0;
        for (int64_t b = 0; b < ncb; b++) v += counts[b];
        cand[2 * e] = (double)v;
        cand[2 * e + 1] = (double)cnt;
    }
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

/* the k = 0 record [fval(p_init), p_init] (TiktakGlobalSearch.f90:496-501) as the only row of the results table, through
 * the f40.20 record like every searchResults.dat row; it is also the best row so far.  valid[] was cleared before. */
__global__ void init_results_kernel(const double* xs, int nx, int K, int roundtrip, double* res, int* valid, double* best) {
    const double f = roundtrip ? tk_f4020_roundtrip(xs[0]) : xs[0];
    if (threadIdx.x == 0) { res[0] = f; valid[0] = 1; best[0] = f; best[1] = 0.0; best[2] = 1.0; }
    for (int d = threadIdx.x; d < nx; d += blockDim.x) {
        const double x = roundtrip ? tk_f4020_roundtrip(xs[(size_t)(d + 1) * K]) : xs[(size_t)(d + 1) * K];
        res[1 + d] = x;
        best[3 + d] = x;
    }
}

} /* namespace */

int tk_launch_init_results(tiktak_handle* h) {
    init_results_kernel<<<1, 64, 0, h->stream>>>(h->d_xstarts, h->cfg.nx, h->cfg.maxpoints_plus1, h->cfg.text_roundtrip, h->d_res,
                                                 h->d_res_valid, h->d_best);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_sort_candidates(tiktak_handle* h, double* d_f, unsigned int* d_i, int64_t n, double* d_sf, unsigned int* d_si);

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
    if (!h->d_sel_list_key) { /* list of the elements that share the threshold's top 16 bits: an eighth of the shard (+ slack) */
        h->sel_list_cap = (unsigned long long)(h->n_local_cblocks * (int64_t)TIKTAK_SOBOL_BLOCK) / 8 + 65536;
        TK_CUDA(cudaMalloc(&h->d_sel_list_key, h->sel_list_cap * sizeof(unsigned long long)));
        TK_CUDA(cudaMalloc(&h->d_sel_list_n, h->sel_list_cap * sizeof(unsigned int)));
    }
    SelView v{h->d_fval, h->n_local_cblocks * (int64_t)TIKTAK_SOBOL_BLOCK, h->kcut, rank, world,
              h->d_sel_list_key, h->d_sel_list_n, h->sel_list_cap};
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
    TK_CUDA(cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int), h->stream));
    int64_t want = (v.nlocal + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 16;
    const unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
    for (int phase = 0; phase < 2; phase++) {
        const int top = phase == 0 ? 56 : 24;
        for (int shift = top; shift >= 0; shift -= 8) {
            select_hist_kernel<<<grid, 256, 0, h->stream>>>(v, h->d_selstate, h->d_hist, h->d_ticket, phase, shift);
            h->launches++;
            if (phase == 0 && shift == 48) {
                select_compact_kernel<<<grid, 256, 0, h->stream>>>(v, h->d_selstate);
                h->launches++;
            }
        }
    }
    select_collect_kernel<<<grid, 256, 0, h->stream>>>(v, h->d_selstate, h->d_cand_f, h->d_cand_i, h->d_cand_n,
                                                       (unsigned int)m);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return tk_sort_candidates(h, h->d_cand_f, h->d_cand_i, m, h->d_sorted_f, h->d_sorted_i);
}

/* order n candidates by (fval, n): d_f/d_i are clobbered (ping-pong), the result lands in d_sf/d_si */
int tk_sort_candidates(tiktak_handle* h, double* d_f, unsigned int* d_i, int64_t n, double* d_sf, unsigned int* d_si) {
    if (n <= 0) return TIKTAK_OK;
    const unsigned int m = (unsigned int)n;
    int passes = 0;
    for (unsigned int run = SORT_CHUNK; run < m; run <<= 1) passes++;
    /* an even number of merge passes must end in (d_sf, d_si): start there when `passes` is even */
    double* af = (passes % 2 == 0) ? d_sf : d_f;
    unsigned int* ai = (passes % 2 == 0) ? d_si : d_i;
    double* bf = (passes % 2 == 0) ? d_f : d_sf;
    unsigned int* bi = (passes % 2 == 0) ? d_i : d_si;
    /* chunk sort reads the input; when it writes in place that is safe (each block loads its chunk first) */
    chunk_sort_kernel<<<(m + SORT_CHUNK - 1) / SORT_CHUNK, 256, 0, h->stream>>>(d_f, d_i, m, af, ai);
    h->launches++;
    for (unsigned int run = SORT_CHUNK; run < m; run <<= 1) {
        merge_pass_kernel<<<(m + 255) / 256, 256, 0, h->stream>>>(af, ai, m, run, bf, bi);
        h->launches++;
        std::swap(af, bf); std::swap(ai, bi);
    }
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_export_candidates(tiktak_handle* h, double* d_cand) {
    const unsigned int M = (unsigned int)(h->cfg.maxpoints_plus1 - 1);
    export_candidates_kernel<<<(M + 1 + 127) / 128, 128, 0, h->stream>>>(h->d_sorted_f, h->d_sorted_i, (unsigned int)h->n_cand, M,
                                                                         h->d_counts, h->n_cblocks, d_cand);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return TIKTAK_OK;
}

int tk_merge_candidates(tiktak_handle* h, const double* d_cand) {
    const unsigned int M = (unsigned int)(h->cfg.maxpoints_plus1 - 1);
    const unsigned int tot = M * (unsigned int)h->cfg.world;
    merge_lists_kernel<<<(tot + 255) / 256, 256, 0, h->stream>>>(d_cand, h->cfg.world, M, h->d_sorted_f, h->d_sorted_i, h->d_cand_n);
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
