/* comm.cu -- the handle's NCCL communicator ("the handle owns all device memory and NCCL communicators").
 *
 * The reference's "communication backend" is lock files on a shared directory (stateControl.f90:12-216); here the
 * shards of one job (Sobol count blocks and restart slots dealt over `world` GPUs) exchange three things, all on the
 * handle's stream and all inside the stage calls of the C ABI:
 *   pretest  per wave of count blocks, only when the `legitimate` cut can fire: all-reduce of the per-block valid counts
 *   choose   one all-gather of each shard's K candidate rows [fval, idx] (+ trailer: the shard's valid count)
 *   local    one all-gather per launch of each shard's result rows [fval, x, #evaluations]
 * NCCL is loaded with dlopen at communicator creation, not linked: a process that already carries a libnccl.so.2 (torch
 * bundles its own) shares that copy, the stand-alone driver picks the system's; a build box without NCCL still links.
 *
 * Two ways to get a communicator:
 *   tiktak_cuda_comm_unique_id + tiktak_cuda_comm_init   one process per GPU (torchrun / mpirun): rank 0 creates the id,
 *                                                        the launcher's own channel carries its 128 bytes to the others
 *   tiktak_cuda_comm_init_all                            one process, one handle per GPU (driver/TiktakGlobalSearch with
 *                                                        TIKTAK_NGPU): ncclCommInitAll; one host thread per handle then
 *                                                        drives the stage calls
 */
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "tk_common.cuh"

namespace {

struct TkNccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

TkNccl* tk_nccl() {
    static TkNccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) return;
#define TK_SYM(field, name) *(void**)(&n.field) = dlsym(n.lib, name)
        TK_SYM(GetUniqueId, "ncclGetUniqueId");
        TK_SYM(CommInitRank, "ncclCommInitRank");
        TK_SYM(CommInitAll, "ncclCommInitAll");
        TK_SYM(CommDestroy, "ncclCommDestroy");
        TK_SYM(AllGather, "ncclAllGather");
        TK_SYM(AllReduce, "ncclAllReduce");
        TK_SYM(GetErrorString, "ncclGetErrorString");
#undef TK_SYM
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommInitAll && n.CommDestroy && n.AllGather && n.AllReduce && n.GetErrorString;
    });
    return n.ok ? &n : nullptr;
}

#define TK_NCCL(call)                                                                                 \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) {                                                                      \
            tk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, tk_nccl()->GetErrorString(r_)); \
            return TIKTAK_ERR_CUDA;                                                                   \
        }                                                                                             \
    } while (0)

} /* namespace */

/* internal: used by the stage drivers in api.cu */
bool tk_comm_ready(const tiktak_handle* h) { return h->comm != nullptr; }
int tk_comm_allgather_f64(tiktak_handle* h, const double* send, double* recv, size_t count) {
    TK_NCCL(tk_nccl()->AllGather(send, recv, count, ncclDouble, (ncclComm_t)h->comm, h->stream));
    return TIKTAK_OK;
}
int tk_comm_allreduce_sum_u64(tiktak_handle* h, unsigned long long* buf, size_t count) {
    TK_NCCL(tk_nccl()->AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)h->comm, h->stream));
    return TIKTAK_OK;
}
int tk_comm_allreduce_max_u64(tiktak_handle* h, unsigned long long* buf, size_t count) {
    TK_NCCL(tk_nccl()->AllReduce(buf, buf, count, ncclUint64, ncclMax, (ncclComm_t)h->comm, h->stream));
    return TIKTAK_OK;
}
void tk_comm_destroy(tiktak_handle* h) {
    if (h->comm && tk_nccl()) tk_nccl()->CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
}

extern "C" {

int tiktak_cuda_comm_unique_id(void* id, int32_t bytes) {
    if (!id || bytes < (int32_t)sizeof(ncclUniqueId)) { tk_set_error("comm_unique_id: the buffer must hold %d bytes", (int)sizeof(ncclUniqueId)); return TIKTAK_ERR_ARG; }
    TkNccl* n = tk_nccl();
    if (!n) { tk_set_error("comm: libnccl.so.2 not found (dlopen)"); return TIKTAK_ERR_UNSUPPORTED; }
    ncclUniqueId u;
    TK_NCCL(n->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return TIKTAK_OK;
}

int tiktak_cuda_comm_init(tiktak_handle* h, const void* id, int32_t bytes) {
    if (!h || !id || bytes < (int32_t)sizeof(ncclUniqueId)) return TIKTAK_ERR_ARG;
    if (h->cfg.world <= 1) return TIKTAK_OK; /* nothing to exchange */
    TkNccl* n = tk_nccl();
    if (!n) { tk_set_error("comm: libnccl.so.2 not found (dlopen)"); return TIKTAK_ERR_UNSUPPORTED; }
    TK_CUDA(cudaSetDevice(h->device));
    tk_comm_destroy(h);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    TK_NCCL(n->CommInitRank(&c, h->cfg.world, u, h->cfg.rank));
    h->comm = c;
    return TIKTAK_OK;
}

int tiktak_cuda_comm_init_all(tiktak_handle** hs, int32_t n_handles) {
    if (!hs || n_handles < 1) return TIKTAK_ERR_ARG;
    if (n_handles == 1) return TIKTAK_OK;
    TkNccl* n = tk_nccl();
    if (!n) { tk_set_error("comm: libnccl.so.2 not found (dlopen)"); return TIKTAK_ERR_UNSUPPORTED; }
    std::vector<int> devs(n_handles);
    for (int r = 0; r < n_handles; r++) {
        if (!hs[r] || hs[r]->cfg.world != n_handles || hs[r]->cfg.rank != r) {
            tk_set_error("comm_init_all: handle %d must be rank %d of world %d", r, r, n_handles);
            return TIKTAK_ERR_ARG;
        }
        devs[r] = hs[r]->device;
        tk_comm_destroy(hs[r]);
    }
    std::vector<ncclComm_t> comms(n_handles, nullptr);
    TK_NCCL(n->CommInitAll(comms.data(), n_handles, devs.data()));
    for (int r = 0; r < n_handles; r++) hs[r]->comm = comms[r];
    return TIKTAK_OK;
}

} /* extern "C" */
