// lv_dist.cu -- multi-GPU plumbing of the strip decomposition: one process per GPU, NCCL over NVLink.
//
// The rectangle is cut into y-strips of cell-list bucket rows.  Each rank clips the polygons of the
// generators it owns; generators of the neighbouring strips that can influence them are present as
// ghost entries of the local cell list (the Python host exchanges them with torch.distributed after
// every move, lagrangianvoronoi.jl_b200/distributed.py).  This file holds what runs inside the
// Krylov loop and therefore must not bounce through Python:
//   * the halo exchange that fills the ghost slots of a slot-ordered vector from their owners
//     (pack kernel -> grouped ncclSend/ncclRecv -> unpack kernel), and
//   * the 2-scalar ncclAllReduce behind every CG dot product.
// NCCL is bound at run time (dlopen of the libnccl.so.2 PyTorch already loaded), the communicator is
// created from a unique id that the host broadcasts with torch.distributed.
#include "lv_internal.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string &err) {
    if (g_nccl.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (lib) break; } // the copy torch loaded
    if (!lib) for (const char *nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define LOAD(sym) *(void **)(&g_nccl.sym) = dlsym(lib, "nccl" #sym); if (!g_nccl.sym) { err = "libnccl lacks nccl" #sym; return false; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(AllReduce) LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
    g_nccl.lib = lib;
    return true;
}
} // namespace

#define LV_NCCL(c, expr)                                                                                          \
    do {                                                                                                          \
        ncclResult_t _r = (expr);                                                                                 \
        if (_r != ncclSuccess) return lv_set_error((c), LV_ECUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

template <int NC>
__global__ void __launch_bounds__(256) k_halo_pack(int64_t n, const int *__restrict__ slots, const double *__restrict__ vec, double *__restrict__ buf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slots[i];
#pragma unroll
    for (int k = 0; k < NC; k++) buf[(size_t)NC * i + k] = vec[(size_t)NC * s + k];
}
template <int NC>
__global__ void __launch_bounds__(256) k_halo_unpack(int64_t n, const int *__restrict__ slots, const double *__restrict__ buf, double *__restrict__ vec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slots[i];
#pragma unroll
    for (int k = 0; k < NC; k++) vec[(size_t)NC * s + k] = buf[(size_t)NC * i + k];
}

int lv_halo_exchange(LvContext *c, double *vec, int ncomp) { // ghost slots of a slot-ordered vector from their owners
    if (lv_strip_peer_mode(c)) return lv_strip_halo_exchange(c, vec, ncomp); // NVLink pulls, no NCCL
    return lv_halo_exchange_nccl(c, vec, ncomp);
}

int lv_halo_exchange_nccl(LvContext *c, double *vec, int ncomp) {
    if (!c->comm || c->peers.empty()) return LV_OK;
    ncclComm_t comm = (ncclComm_t)c->comm;
    cudaStream_t st = c->stream;
    if (c->halo_send_total > 0) {
        const int nb = (int)((c->halo_send_total + 255) / 256);
        if (ncomp == 1) k_halo_pack<1><<<nb, 256, 0, st>>>(c->halo_send_total, c->d_send_slots, vec, c->d_send_buf);
        else k_halo_pack<2><<<nb, 256, 0, st>>>(c->halo_send_total, c->d_send_slots, vec, c->d_send_buf);
        c->launches++;
    }
    LV_NCCL(c, g_nccl.GroupStart());
    for (const auto &p : c->peers) {
        if (p.nsend > 0) LV_NCCL(c, g_nccl.Send(c->d_send_buf + (size_t)ncomp * p.send_off, (size_t)ncomp * p.nsend, ncclDouble, p.rank, comm, st));
        if (p.nrecv > 0) LV_NCCL(c, g_nccl.Recv(c->d_recv_buf + (size_t)ncomp * p.recv_off, (size_t)ncomp * p.nrecv, ncclDouble, p.rank, comm, st));
    }
    LV_NCCL(c, g_nccl.GroupEnd());
    if (c->halo_recv_total > 0) {
        const int nb = (int)((c->halo_recv_total + 255) / 256);
        if (ncomp == 1) k_halo_unpack<1><<<nb, 256, 0, st>>>(c->halo_recv_total, c->d_recv_slots, c->d_recv_buf, vec);
        else k_halo_unpack<2><<<nb, 256, 0, st>>>(c->halo_recv_total, c->d_recv_slots, c->d_recv_buf, vec);
        c->launches++;
    }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// mailbox layout per rank: [parity 2][rank 64] of {double v[2]; int flag; int pad} = 24 -> 32 bytes
typedef LvMailSlot MailSlot;
#define MB_MAX_RANKS LV_MB_MAX_RANKS

// one block, one thread per rank: post my two partial sums into everybody's mailbox (release), wait until
// everybody's partials for this sequence number have arrived in mine (acquire), sum them in rank order
__global__ void k_peer_allreduce(MailSlot *const *mailboxes, int nranks, int rank, int seq, double *vals) {
    __shared__ double s0[MB_MAX_RANKS], s1[MB_MAX_RANKS];
    const int q = threadIdx.x;
    const int par = seq & 1;
    if (q < nranks) {
        MailSlot *dst = mailboxes[q] + par * MB_MAX_RANKS + rank;
        dst->v[0] = vals[0];
        dst->v[1] = vals[1];
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(&dst->flag), "r"(seq) : "memory");
        MailSlot *src = mailboxes[rank] + par * MB_MAX_RANKS + q;
        while (ld_acquire_sys(&src->flag) < seq) { }
        s0[q] = __ldcv(&src->v[0]);
        s1[q] = __ldcv(&src->v[1]);
    }
    __syncthreads();
    if (q == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < nranks; k++) { a += s0[k]; b += s1[k]; }
        vals[0] = a;
        vals[1] = b;
    }
}

int lv_allreduce_sum(LvContext *c, double *dev, int count) {
    if (!c->comm) return LV_OK;
    if (c->mailbox_ready && count == 2) {
        c->ar_seq++;
        k_peer_allreduce<<<1, MB_MAX_RANKS, 0, c->stream>>>((MailSlot *const *)c->d_mailbox_ptrs, c->nranks, c->rank, c->ar_seq, dev);
        c->launches++;
        return LV_OK;
    }
    LV_NCCL(c, g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
    return LV_OK;
}

static void close_mailboxes(LvContext *c) {
    for (size_t q = 0; q < c->mailbox_of.size(); q++)
        if ((int)q != c->rank && c->mailbox_of[q]) cudaIpcCloseMemHandle(c->mailbox_of[q]);
    c->mailbox_of.clear();
    c->mailbox_ready = false;
}

void lv_dist_destroy(LvContext *c) {
    lv_strip_destroy(c);
    close_mailboxes(c);
    cudaFree(c->d_mailbox); c->d_mailbox = nullptr;
    cudaFree(c->d_mailbox_ptrs); c->d_mailbox_ptrs = nullptr;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    cudaFree(c->d_send_slots); cudaFree(c->d_recv_slots); cudaFree(c->d_send_buf); cudaFree(c->d_recv_buf);
    c->d_send_slots = c->d_recv_slots = nullptr;
    c->d_send_buf = c->d_recv_buf = nullptr;
}

extern "C" {

int32_t lv_comm_unique_id(uint8_t *out128) {
    std::string err;
    if (!out128) return LV_EINVAL;
    if (!load_nccl(err)) return lv_set_error(nullptr, LV_ECUDA, "%s", err.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return lv_set_error(nullptr, LV_ECUDA, "ncclGetUniqueId failed");
    memcpy(out128, &id, 128);
    return LV_OK;
}

int32_t lv_comm_init(LvHandle c, int32_t rank, int32_t nranks, const uint8_t *id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return lv_set_error(c, LV_EINVAL, "bad communicator arguments");
    LV_ENTER(c);
    std::string err;
    if (!load_nccl(err)) return lv_set_error(c, LV_ECUDA, "%s", err.c_str());
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    LV_NCCL(c, g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    c->rank = rank;
    c->nranks = nranks;
    return LV_OK;
}

// Halo plan of the current mesh: for every peer rank, the slots whose values this rank sends (its own
// primary slots that are ghosts over there) and the ghost slots it receives into, both in the order the
// two sides agreed on (the receiver's request order).  Slot lists are device arrays.
int32_t lv_halo_plan(LvHandle c, int32_t npeers, const int32_t *peer_rank, const int64_t *send_count, const int32_t *send_slots_dev,
                     const int64_t *recv_count, const int32_t *recv_slots_dev) {
    if (!c || npeers < 0) return LV_EINVAL;
    LV_ENTER(c);
    c->peers.clear();
    int64_t so = 0, ro = 0;
    for (int k = 0; k < npeers; k++) {
        LvContext::HaloPeer p{peer_rank[k], send_count[k], recv_count[k], so, ro};
        so += send_count[k];
        ro += recv_count[k];
        c->peers.push_back(p);
    }
    c->halo_send_total = so;
    c->halo_recv_total = ro;
    if (so > c->cap_halo_send) {
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_send_slots); cudaFree(c->d_send_buf);
        c->cap_halo_send = so + so / 8 + 1024;
        LV_CUDA(c, cudaMalloc((void **)&c->d_send_slots, sizeof(int) * (size_t)c->cap_halo_send));
        LV_CUDA(c, cudaMalloc((void **)&c->d_send_buf, sizeof(double) * 2 * (size_t)c->cap_halo_send));
    }
    if (ro > c->cap_halo_recv) {
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_recv_slots); cudaFree(c->d_recv_buf);
        c->cap_halo_recv = ro + ro / 8 + 1024;
        LV_CUDA(c, cudaMalloc((void **)&c->d_recv_slots, sizeof(int) * (size_t)c->cap_halo_recv));
        LV_CUDA(c, cudaMalloc((void **)&c->d_recv_buf, sizeof(double) * 2 * (size_t)c->cap_halo_recv));
    }
    if (so > 0) LV_CUDA(c, cudaMemcpyAsync(c->d_send_slots, send_slots_dev, sizeof(int) * (size_t)so, cudaMemcpyDeviceToDevice, c->stream));
    if (ro > 0) LV_CUDA(c, cudaMemcpyAsync(c->d_recv_slots, recv_slots_dev, sizeof(int) * (size_t)ro, cudaMemcpyDeviceToDevice, c->stream));
    return LV_OK;
}

// CUDA IPC handle (64 B) of this rank's allreduce mailbox; lv_mailbox_plan maps the mailboxes of all ranks
int32_t lv_mailbox_export(LvHandle c, uint8_t *out64) {
    if (!c || !out64) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->d_mailbox) {
        LV_CUDA(c, cudaMalloc(&c->d_mailbox, sizeof(MailSlot) * 2 * MB_MAX_RANKS));
        LV_CUDA(c, cudaMemset(c->d_mailbox, 0, sizeof(MailSlot) * 2 * MB_MAX_RANKS));
    }
    cudaIpcMemHandle_t hm;
    LV_CUDA(c, cudaIpcGetMemHandle(&hm, c->d_mailbox));
    memcpy(out64, &hm, 64);
    return LV_OK;
}

int32_t lv_mailbox_plan(LvHandle c, int32_t nranks, const uint8_t *handles /* nranks x 64 */) {
    if (!c || !c->comm || nranks != c->nranks || nranks > MB_MAX_RANKS) return lv_set_error(c, LV_EINVAL, "lv_mailbox_plan: bad arguments");
    LV_ENTER(c);
    if (c->mailbox_ready && !memcmp(c->mailbox_handles, handles, 64 * (size_t)nranks)) return LV_OK;
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    close_mailboxes(c);
    c->mailbox_of.assign((size_t)nranks, nullptr);
    for (int q = 0; q < nranks; q++) {
        if (q == c->rank) { c->mailbox_of[q] = c->d_mailbox; continue; }
        cudaIpcMemHandle_t hm;
        memcpy(&hm, handles + 64 * q, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&c->mailbox_of[q], hm, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            c->mailbox_of[q] = nullptr;
            close_mailboxes(c);
            return lv_set_error(c, LV_ECUDA, "cudaIpcOpenMemHandle(mailbox of rank %d) failed: %s", q, cudaGetErrorString(e));
        }
    }
    if (!c->d_mailbox_ptrs) LV_CUDA(c, cudaMalloc((void **)&c->d_mailbox_ptrs, sizeof(void *) * MB_MAX_RANKS));
    LV_CUDA(c, cudaMemcpy(c->d_mailbox_ptrs, c->mailbox_of.data(), sizeof(void *) * (size_t)nranks, cudaMemcpyHostToDevice));
    memcpy(c->mailbox_handles, handles, 64 * (size_t)nranks);
    c->mailbox_ready = true;
    return LV_OK;
}

// fall back to the NCCL halo / allreduce (used when a rank could not map a peer: all ranks must switch together)
int32_t lv_peer_disable(LvHandle c) {
    if (!c) return LV_EINVAL;
    c->strip.mapped = false;
    c->mailbox_ready = false;
    return LV_OK;
}

// unmap everything mapped from other ranks (peer vectors, flags, mailboxes).  Every rank calls this, then a barrier,
// then lv_destroy: memory exported over CUDA IPC must not be freed while an importer still has it mapped.
int32_t lv_peer_close(LvHandle c) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    lv_strip_unmap(c);
    close_mailboxes(c);
    return LV_OK;
}

// exchange a caller's slot-ordered device vector (tests; the solver calls lv_halo_exchange directly)
int32_t lv_halo_exchange_dev(LvHandle c, double *vec_dev, int32_t ncomp) {
    if (!c || !vec_dev || (ncomp != 1 && ncomp != 2)) return LV_EINVAL;
    LV_ENTER(c);
    return lv_halo_exchange(c, vec_dev, ncomp);
}

// device pointers of the slot-ordered cell list, for the host to build the halo plan
// which: 0 ent_label (uint32[nslot]), 1 prim_of_label (int32[n]), 2 own (uint8[nslot]), 3 ent_xy (double2[nslot]),
//        4 P (double[nslot]), 5 area (double[nslot])
int32_t lv_device_array(LvHandle c, int32_t which, void **ptr, int64_t *count) {
    if (!c || !ptr || !count) return LV_EINVAL;
    switch (which) {
    case 0: *ptr = c->d_ent_label; *count = c->nslot; break;
    case 1: *ptr = c->d_prim_of_label; *count = c->n; break;
    case 2: *ptr = c->d_own; *count = c->nslot; break;
    case 3: *ptr = c->d_ent_xy; *count = c->nslot; break;
    case 4: *ptr = c->d_P; *count = c->nslot; break;
    case 5: *ptr = c->d_area; *count = c->nslot; break;
    case 6: *ptr = c->strip.loc_xy; *count = c->strip.n_loc; break;  // strip mode: local generators (owned first, then ghosts)
    case 7: *ptr = c->strip.loc_key; *count = c->strip.n_loc; break; // ... and their global labels (int32)
    case 8: *ptr = c->strip.loc_key; *count = c->strip.n_own; break; // the owned generators' global labels
    default: return lv_set_error(c, LV_EINVAL, "unknown device array %d", which);
    }
    return LV_OK;
}

// remesh on the generators present on this rank (owned + ghosts); buckets are ordered by order_key (the
// global labels) so that the candidate order equals the single-GPU order; only polygons with
// owned_mask != 0 are clipped
int32_t lv_remesh_owned_dev(LvHandle c, int64_t n_local, const double *xy_dev, const uint8_t *owned_mask_dev,
                            const int32_t *order_key_dev) {
    if (!c) return LV_EINVAL;
    c->owned_mask = owned_mask_dev;
    c->order_key = order_key_dev;
    int32_t st = lv_remesh_dev(c, n_local, xy_dev);
    return st;
}

} // extern "C"
