/* tk_common.cuh -- shared declarations of libtiktak_cuda's translation units. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/tiktak_cuda.h"
#include "../../include/tiktak_obj.cuh"
#include "tk_device.cuh"

void tk_set_error(const char* fmt, ...);
#define TK_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            tk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return TIKTAK_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

__host__ __device__ inline int64_t tk_local_index(int64_t n, int world) {
    const int64_t cb = n / TIKTAK_SOBOL_BLOCK;
    return (cb / world) * TIKTAK_SOBOL_BLOCK + (n % TIKTAK_SOBOL_BLOCK);
}

/* one row of an asynchronous run as the scans of nm_quad_async_kernel read it */
struct alignas(16) TkAsyncRow { double fv; int fin; int slt; };

struct tiktak_handle {
    tiktak_config cfg;
    std::vector<double> range, init, bound, aux;
    std::vector<uint32_t> V; /* [nx][32] */
    std::vector<double> w11; /* blend weight per restart k (index k, 1-based) */
    std::vector<double> simplex; /* simplex_coordinates2, (nx, nx+1) column-major */
    double nm_scale = 1.0;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evw0 = nullptr, evw1 = nullptr;
    bool wave_pending = false;
    cudaGraphExec_t stage_graph[2] = {nullptr, nullptr}; /* the one-wave pre-test stage as a graph, per counts buffer (api.cu pretest_wave) */
    const void* stage_graph_counts[2] = {nullptr, nullptr};
    int stage_runs = 0;
    int sm_count = 148;
    int64_t launches = 0;
    TkScalars sc;
    /* device copies */
    uint32_t* d_V = nullptr;
    uint32_t* d_sobtab = nullptr; /* bit-major tables of the fixed-nx Sobol kernel (kernels_sobol.cu) */
    double *d_lo = nullptr, *d_w = nullptr, *d_bl = nullptr, *d_bu = nullptr, *d_aux = nullptr, *d_init = nullptr;
    double *d_simplex = nullptr, *d_w11 = nullptr;
    /* pre-test state */
    int64_t n_cblocks = 0, n_local_cblocks = 0;
    double* d_fval = nullptr;             /* local fvals */
    unsigned long long* d_counts = nullptr; /* [n_cblocks] */
    unsigned long long* d_counts_spare = nullptr; /* zeroed twin: the next job swaps to it */
    unsigned long long* h_counts = nullptr; /* pinned host mirror of d_counts */
    void* h_fetch = nullptr;                /* pinned staging of tiktak_cuda_local_fetch */
    bool counts_copied = false;            /* the whole-stage path has queued the copy of all counts to h_counts already */
    int64_t cblocks_done = 0;
    int64_t kcut = -1, nlegit = 0;
    bool pretest_done = false;
    /* selection state */
    unsigned long long* d_hist = nullptr; /* 256 bins */
    unsigned long long* d_selstate = nullptr;
    double* d_cand_f = nullptr;           /* K-1 candidates */
    unsigned int* d_cand_i = nullptr;
    unsigned int* d_cand_n = nullptr;
    unsigned int* d_ticket = nullptr;   /* last-block ticket of the radix-select passes */
    unsigned long long* d_sel_list_key = nullptr; /* radix select: elements sharing the threshold's top 16 bits */
    unsigned int* d_sel_list_n = nullptr;
    unsigned long long sel_list_cap = 0;
    double* d_sorted_f = nullptr;
    unsigned int* d_sorted_i = nullptr;
    int64_t n_cand = 0;
    /* starts + results */
    double* d_xstarts = nullptr; /* (K, nx+1) column-major */
    bool starts_ready = false;
    double* d_res = nullptr;     /* (K+1) rows x (nx+1): [fval, x], row k (row 0 = the k=0 record) */
    int* d_res_valid = nullptr;
    std::vector<double> h_res_f;
    std::vector<int> h_res_valid, h_res_imp;
    double* d_best = nullptr;    /* [0]=fval [1]=k [2]=count, [3..3+nx) = x */
    /* local stage scratch */
    double *d_starts = nullptr, *d_out_x = nullptr, *d_out_f = nullptr, *d_width = nullptr;
    int* d_out_evals = nullptr;
    double* d_wshr = nullptr;        /* nx: simplex widths from the best minima so far (nm_shrink_mode = 1) */
    int* d_shr_have = nullptr;
    double* d_arch_starts = nullptr; /* (K, nx) column-major: the searchStart.dat rows of every restart */
    int* d_arch_evals = nullptr;     /* [K+1]: objective evaluations of restart k */
    bool row0_on_device_only = false; /* the k = 0 record was written by init_results_kernel; host mirror not filled yet */
    bool best_dirty = true;          /* d_best must be recomputed from the rows (after push_results) */
    cudaEvent_t evl0 = nullptr, evl1 = nullptr, evp0 = nullptr, evp1 = nullptr;
    bool local_timing_open = false, pretest_timing_open = false;
    cudaEvent_t ev_wave_start = nullptr, ev_stage_end = nullptr; /* which of the events above bracket the current wave / close the stage */
    bool rho_ready = false;
    /* selectType = 1 (lottery) */
    double *d_lot_f = nullptr, *d_lot_sf = nullptr, *d_lot_cum = nullptr;
    unsigned int *d_lot_k = nullptr, *d_lot_sk = nullptr;
    int* d_lot_point = nullptr;       /* [K+1]: lotPoint of restart k (the row it was blended toward) */
    std::vector<double> lot_cum;
    int lot_cum_n = 0;
    int n_rows = 0;                   /* result rows known to the handle (k = 0 record included) */
    int* d_accepted = nullptr;        /* restarts ingested by the last multi-round ingest */
    /* multi-GPU: NCCL communicator owned by the handle (comm.cu) + exchange buffers */
    void* comm = nullptr;
    cudaStream_t copy_stream = nullptr; /* asynchronous copy of the stored values to the host (pretest_values_shard_async) */
    cudaEvent_t ev_copy = nullptr;
    double *d_xsend = nullptr, *d_xrecv = nullptr;
    size_t xbuf_doubles = 0;
    bool counts_global = false;       /* d_counts holds the all-reduced (global) per-block counts */
    double* d_nm_pslab = nullptr;     /* simplices of nm_w4_kernel's resident warps (nx >= 32) */
    size_t nm_pslab_doubles = 0;
    int* d_nmq = nullptr;             /* control words of the throughput Nelder-Mead kernel's restart queue (nm_common.cuh) */
    int blend_w_idx = -1;             /* > 0: the next launch blends with the weight of restart number blend_w_idx (missing-search sweep) */
    TkAsyncRow* d_async_ev = nullptr; /* asynchronous instances (kernels_nm.cu): value, virtual finish time and instance of row k */
    int* d_async_clk = nullptr;       /* ... and the instances' clocks */
    int* d_async_base = nullptr;      /* ... and the row each restart was blended toward */
    void* d_async_desc = nullptr;     /* ... and the device copy of the run's descriptor (one-warp instances, kernels_nm_w4.cu) */
    long long* async_dbg = nullptr;   /* experiment counters (TIKTAK_ASYNC_DEBUG) */
    bool async_rows = false;          /* the rows of the table were written by an asynchronous run: local_fetch orders them by (finish time, instance) */
    int nm_form = -1;                 /* 1: one warp per restart (kernels_nm_w4.cu), 0: four warps per restart (kernels_nm.cu), -1: not decided yet */
    double *d_objfun_x = nullptr, *d_objfun_f = nullptr; /* staging of tiktak_cuda_objfun for host arguments */
    int64_t objfun_scratch_pts = 0;
    /* DFNLS */
    double* d_dfnls_slab = nullptr;
    size_t dfnls_slab_doubles = 0;
    double* d_rho = nullptr;
    int* d_out_status = nullptr;
    /* TIKTAK_OBJ_SMM */
    void* smm = nullptr;
    int smm_T = 0;
    /* stage timings */
    double ms_pretest = 0, ms_choose = 0, ms_local = 0, ms_pretest_eval = 0;
};

/* kernel launchers (one per translation unit) */
int tk_launch_sobol_wave(tiktak_handle* h, const TkWave& wv);
int tk_launch_sobol_points(tiktak_handle* h, int64_t first, int64_t count, double* d_out);
int tk_launch_objfun(tiktak_handle* h, const double* d_x, int64_t n, double* d_f);
void* tk_rtc_sobol_kernel(tiktak_handle* h, int P, int inb, int G, int* per_sm); /* rtc.cu: the fixed-nx Sobol kernel compiled at run time */
int tk_rtc_launch(tiktak_handle* h, void* fn, unsigned grid, void** args);
int tk_launch_find_cut(tiktak_handle* h, int64_t cb, unsigned long long need, unsigned long long* d_result);
int tk_select_top(tiktak_handle* h, int64_t M);
int tk_build_starts(tiktak_handle* h);
int tk_launch_init_results(tiktak_handle* h);
int tk_launch_best(tiktak_handle* h);
int tk_launch_blend(tiktak_handle* h, int first_k, int n_k, int w_idx = -1);
int tk_launch_blend_lottery(tiktak_handle* h, int first_k, int n_k, int count, int w_idx = -1);
int tk_launch_nm(tiktak_handle* h, int first_k, int n_k, int commit_round_size = 0);
int tk_launch_nm_async(tiktak_handle* h, int first_k, int n_k, int instances, int* used, bool free_run = false);
bool tk_nm_async_supported(tiktak_handle* h);
int tk_nm_async_form(tiktak_handle* h, int instances); /* -1: not built for the configuration; 0: four-warp blocks, tick = loop trip; 1: one warp, tick = evaluation */
bool tk_nm_use_w4(tiktak_handle* h);
int tk_launch_dfnls(tiktak_handle* h, int first_k, int n_k, int instances = 0, int* used = nullptr);
int tk_launch_ingest(tiktak_handle* h, int first_k, int n_k, int round_size, const double* pack, int* d_accepted);
int tk_nm_wave_capacity(tiktak_handle* h);
int tk_launch_pack(tiktak_handle* h, int n_k, double* pack);
int tk_launch_shrink_width(tiktak_handle* h);
int tk_fp64_peak(int device, double* flops, double* ms);
bool tk_comm_ready(const tiktak_handle* h);
int tk_comm_allgather_f64(tiktak_handle* h, const double* send, double* recv, size_t count);
int tk_comm_allreduce_sum_u64(tiktak_handle* h, unsigned long long* buf, size_t count);
int tk_comm_allreduce_max_u64(tiktak_handle* h, unsigned long long* buf, size_t count);
void tk_comm_destroy(tiktak_handle* h);

/* objective dispatch: calls f.template operator()<Obj>() for the configured plugin */
#include "objectives.cuh"
template <class F>
inline int tk_dispatch_obj(int id, F&& f) {
    switch (id) {
        case TIKTAK_OBJ_TEMPLATE: return f.template operator()<ObjTemplate>();
        case TIKTAK_OBJ_GRIEWANK: return f.template operator()<ObjGriewank>();
        case TIKTAK_OBJ_RASTRIGIN: return f.template operator()<ObjRastrigin>();
        case TIKTAK_OBJ_ROSENBROCK: return f.template operator()<ObjRosenbrock>();
    }
    tk_set_error("objective id %d has no scalar plugin", id);
    return TIKTAK_ERR_UNSUPPORTED;
}

