// lv_internal.cuh -- shared declarations of liblvb200 (sm_100a only).
//
// Data layout in HBM (everything structure-of-arrays, indexed by "slot"):
//   A slot is one entry of the bucket-sorted cell list: every generator contributes one
//   primary slot plus one slot per periodic image that falls inside the padded cell list
//   (voronoigrid.jl:130-147).  Slots are ordered by (bucket, label) -- exactly the order in
//   which `julia -t 1` finds labels inside a bucket -- so slot order is also a spatial order
//   and all per-cell data (mesh rows, pressure vectors) live in it.  Labels are mapped back
//   only at the C-ABI boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <functional>
#include "../../include/lv_capi.h"

#define LV_SP_MAXP 4 // strip neighbours per rank (2 in practice)
#define LV_STATE_MAX 16 // resident state fields (lv_step.cu)
#define LV_IMAGE_BIT 0x80000000u // set in ent_label[] for periodic-image slots

struct LvPathNode { // neighborlist.jl:6-9, truncated table (see lv_capi.cu:build_magic_path)
    int i1, i2;
    double rr;
};

struct LvGridParams { // passed by value to kernels
    double h;            // cell-list bucket size (neighborlist.jl:16)
    double rr_max;       // r_max^2 (voronoigrid.jl:40)
    double ox, oy;       // cell-list origin (neighborlist.jl:23)
    double cminx, cminy, cmaxx, cmaxy; // cropping_rect (voronoigrid.jl:29-33)
    double xperiod, yperiod;           // voronoigrid.jl:31-32
    int xper, yper;
    int n1, n2;          // buckets per axis (neighborlist.jl:24-25)
    int npath;           // nodes in the truncated magic_path table
};

struct LvProfSlot {
    double ms = 0.0;
    int64_t launches = 0;
};

struct LvPipe; // lv_pipeline.cu

struct LvContext {
    int device = 0;
    // pipelined host-buffer mode (lv_set_async_edges(h, 3), lv_pipeline.cu): pipe_pending = the last remesh is only
    // queued; lv_pipe_finish completes it (every entry point does that first, see LV_ENTER)
    LvPipe *pipe = nullptr;
    bool pipe_mode = false, pipe_pending = false;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::string err;
    // grid (voronoigrid.jl:14-25)
    double dr = 0, h = 0, r_max = 0;
    double bmin[2], bmax[2], cmin[2], cmax[2];
    LvGridParams gp{};
    LvPathNode *d_path = nullptr;
    LvPathNode *h_path = nullptr;
    int64_t ncell = 0;
    // generators (label order)
    int64_t n = 0;
    double2 *d_xy = nullptr;     // owned copy of the positions (host-buffer path)
    const double2 *xy = nullptr; // positions used by the last remesh (d_xy or caller's)
    int64_t cap_n = 0;
    // cell list
    int *d_cell_cnt = nullptr;   // [ncell+1] counts, then fill cursors
    int *d_cell_start = nullptr; // [ncell+1] exclusive scan
    int64_t nslot = 0, cap_slot = 0;
    unsigned *d_ent_label = nullptr; // [nslot] 0-based label | LV_IMAGE_BIT
    double2 *d_ent_xy = nullptr;     // [nslot] ORIGINAL position of the label (q.x)
    int *d_prim_of_label = nullptr;  // [n] label -> primary slot (-1: no primary slot in the local cell list)
    unsigned char *d_own = nullptr;  // [nslot] 1 = primary slot of a generator this rank owns (a real row)
    const unsigned char *owned_mask = nullptr; // [n] caller's device mask of owned generators (NULL: all)
    const int *order_key = nullptr;            // [n] caller's device ordering key (global labels) or NULL = the label itself
    // mesh (slot order)
    int *d_rowptr = nullptr; // [nslot] first edge of the row (rows are NOT stored in slot order, see lv_clip_fast.cu)
    unsigned char *d_deg = nullptr; // [nslot] number of edges of the row
    int64_t nnz = 0, cap_nnz = 0;
    int *d_col = nullptr;    // [nnz] neighbour primary slot, or wall code -1..-4
    double2 *d_v1 = nullptr, *d_v2 = nullptr; // [nnz] edge end points (clockwise, polygon.jl:51-97)
    double *d_area = nullptr;                 // [nslot]
    double2 *d_cen = nullptr;                 // [nslot]
    unsigned long long *d_tile_state = nullptr; // look-back scan state of the clip kernel
    void *d_park_v = nullptr, *d_park_l = nullptr, *d_park_nxt = nullptr, *d_park_hdr = nullptr; // parked rings of the lane-refill clip kernel
    int64_t cap_park_v = 0, cap_park_l = 0, cap_park_n = 0, cap_park_h = 0;
    int64_t cap_tiles = 0;
    int *d_flags = nullptr; // [8] device status words, see LvFlag
    int *d_tickets = nullptr; // [8] last-block tickets of fused kernels (0 matvec, 1 update_r, 2 update_xp); [7] = a peer wait timed out
    int *h_flags = nullptr; // pinned mirror
    int clip_level = 0;     // sticky polygon-capacity level (see lv_clip_run)
    int clip_last_level = 0; // level that produced the current mesh
    int64_t clip_anomalies = 0; // remeshes that had to be replayed by the edge-list kernel
    bool mesh_valid = false;
    // scratch for scans / label-order staging
    void *d_scratch = nullptr;
    int64_t cap_scratch = 0;
    void *d_stage_buf[2] = {nullptr, nullptr}; // label-order staging of the mesh for the device->host copy
    int64_t cap_stage_buf[2] = {0, 0};
    void *d_io_stage = nullptr;                // grow-only label-order staging of per-cell fields (host <-> slot order);
    int64_t cap_io_stage = 0;                  // persistent because cudaFree would wait for the background edge copies
    cudaStream_t copy_stream = nullptr; // background device->host copy of the edge view (lv_set_async_edges)
    cudaEvent_t ev_conv_done = nullptr, ev_stage_done[2] = {nullptr, nullptr};
    bool async_edges = false, async_all = false, stage_pending[2] = {false, false};
    int stage_cur = 0;
    bool flags_mapped = true; // status words via mapped pinned memory (default) or cudaMemcpyAsync (LV_FLAG_MODE=memcpy)
    // pressure (slot order)
    bool pr_valid = false;
    int64_t pr_cap = 0;
    double *d_mass = nullptr, *d_rho = nullptr, *d_c2 = nullptr, *d_P = nullptr;
    double2 *d_v = nullptr, *d_GP = nullptr;
    double *d_diag = nullptr, *d_w = nullptr; // operator: diagonal [nslot], weights [nnz]
    double *d_dinv = nullptr;                 // [nslot] 1/A_ii, stored as FLOAT (the buffer is sized for doubles): Jacobi preconditioner of LV_SOLVER_PCG
    double *d_lrr = nullptr;                  // [nnz] lr_ratio of the edge (polygon.jl:228)
    double2 *d_mx = nullptr, *d_mz = nullptr; // [nnz] m - p.x and m - z (pressure.jl:178,198)
    double *d_bvel = nullptr;                 // [nslot] P-independent part of the right-hand side
    bool bvel_valid = false;
    double asm_dt = 0.0;
    // per-boundary-edge data (pressure.jl:182 boundary_velocity(midpoint(e), e.label) evaluated by the host per edge):
    // boundary edges are numbered polygon by polygon in label order, inside a polygon in edge order
    int *d_bdry_ptr = nullptr;   // [n+1] first boundary-edge number of every label
    int64_t cap_bdry = 0, n_bedge = 0;
    bool bdry_valid = false;
    double2 *d_vbc_edge = nullptr; // [n_bedge] wall velocity per boundary edge, or unused
    unsigned char *d_bf_on = nullptr; // [n_bedge] bdary_friction! charfun(m) per boundary edge
    int64_t cap_bedge = 0;
    bool vbc_edge_on = false;
    double last_vbc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t cap_w = 0;
    double *d_b = nullptr;
    double *d_vec[8] = {nullptr}; // Krylov workspace vectors
    double *d_red = nullptr;      // reduction partials + scalars
    double *h_red = nullptr;      // pinned
    bool assembled = false;
    int cg_hint = 0; // iterations of the previous solve (sizes the first launch batch)
    // device-resident polygon fields in label order (lv_step.cu)
    double *st_field[16] = {nullptr};
    double *st_tmp = nullptr;
    bool st_x_alias = false; // strip mode: st_field[0] (x) is the strip's position array, not an allocation of its own
    int64_t st_n = 0, st_cap = 0;
    // multi-GPU: NCCL communicator + halo plan (lv_dist.cu)
    void *comm = nullptr; // ncclComm_t
    int rank = 0, nranks = 1;
    struct HaloPeer { int rank; int64_t nsend, nrecv; int64_t send_off, recv_off; };
    std::vector<HaloPeer> peers;
    int *d_send_slots = nullptr, *d_recv_slots = nullptr; // concatenated per peer
    double *d_send_buf = nullptr, *d_recv_buf = nullptr;  // 2 components per entry
    int64_t halo_send_total = 0, halo_recv_total = 0, cap_halo_send = 0, cap_halo_recv = 0;
    // strip exchange over peer memory (lv_strip.cu): the library owns the local generator arrays (owned first, then the
    // ghosts peer by peer) and one exported exchange area per rank that its strip neighbours map once (CUDA IPC).  Ghost
    // generators, their counts and the halo values of every slot-ordered vector are PULLED out of the owner's area over
    // NVLink, ordered by sequence flags in the area's header; nothing is reallocated after lv_strip_setup.
    struct StripPeer {
        int rank = -1, idx_there = 0; // my index in that peer's peer list
        int lo = 0, hi = 0;           // bucket rows [lo, hi) whose generators that peer needs (its rows +- H)
        char *area = nullptr;         // the peer's exchange area, mapped
        int64_t nsend = 0, nrecv = 0, soff = 0, roff = 0;
    };
    struct Strip {
        bool on = false;
        int npeers = 0;
        StripPeer peer[LV_SP_MAXP];
        int64_t capg = 0, cap_loc = 0, n_own = 0, n_loc = 0;
        char *area = nullptr;      // my exchange area (exported)
        size_t area_bytes = 0;
        double2 *loc_xy = nullptr; // [cap_loc]
        int *loc_key = nullptr;    // [cap_loc] global labels (bucket ordering key)
        int *sel = nullptr;        // [MAXP][capg] owned indices sent to each peer, in the order of the outbox
        int *cnt = nullptr;        // [8] device counters: [0..4) ghosts selected per peer, [4..6) send-slot bounds
        int *h_counts = nullptr;   // mapped pinned [16]
        int *send_pos0 = nullptr, *send_pos1 = nullptr; // [cap_slot] slot -> position in the halo outbox, or -1
        int64_t cap_pos = 0;
        int gseq = 0, hseq = 0, mseq = 0; // remeshes / halo exchanges / migrations issued so far (same on every rank)
        int *mig = nullptr, *holes = nullptr; // migration scratch: leavers per peer; holes + tails of the compaction
        unsigned char *leave = nullptr;       // [cap_loc]
        std::vector<int> world_rows;          // R[0..world]: bucket rows owned by every rank
        int last_mig_out = 0, last_mig_in = 0;
        bool mapped = false;
    } strip;
    int64_t owned_count = -1; // >= 0: labels below it are owned (strip mode), instead of owned_mask
    // peer-memory allreduce of the CG scalars: every rank owns a mailbox (2 parities x nranks x {2 doubles, flag})
    // that all other ranks write into over NVLink; sums are taken in rank order (deterministic)
    void *d_mailbox = nullptr;             // this rank's mailbox (exported)
    std::vector<void *> mailbox_of;        // [nranks] mapped mailboxes (own entry = d_mailbox)
    void **d_mailbox_ptrs = nullptr;       // device copy of mailbox_of
    unsigned char mailbox_handles[64 * 64];
    int ar_seq = 0;                        // reductions done so far (same on every rank)
    bool mailbox_ready = false;
    // instrumentation
    bool prof_on = false;
    LvProfSlot prof[LV_PROF_COUNT];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct ProfPending { int slot; cudaEvent_t a, b; };
    std::vector<ProfPending> prof_pending; // recorded, not yet resolved (no sync in the hot loop)
    std::vector<cudaEvent_t> prof_free;
    int64_t launches = 0;
    int64_t dev_bytes = 0;
    int num_sms = 148;
};

enum LvFlag { LVF_DESTROYED = 0, LVF_NAN = 1, LVF_OVERFLOW = 2, LVF_TICKET = 3, LVF_NNZ = 4, LVF_CONV = 5 };

// Device status words reach the host through mapped pinned memory (a 1-block kernel stores them, the host reads them
// after a stream synchronise) instead of a device->host memcpy: small reads must not queue behind a multi-GB
// background copy of the edge view on the same copy engine (lv_set_async_edges).
int lv_publish_flags(LvContext *c, const int *extra_dev_int /* nullable: lands in h_flags[8] */);

// ---- error helpers -------------------------------------------------------------------------
int lv_set_error(LvContext *c, int code, const char *fmt, ...);
#define LV_CUDA(c, expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return lv_set_error((c), LV_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                          \
    } while (0)
#define LV_TRY(expr)                       \
    do {                                   \
        int _s = (expr);                   \
        if (_s != LV_OK) return _s;        \
    } while (0)

// entry-point prologue: select the device and complete a deferred remesh of the pipelined mode
int lv_pipe_finish(LvContext *c);
#define LV_ENTER(c)                                             \
    do {                                                        \
        LV_CUDA((c), cudaSetDevice((c)->device));               \
        if ((c)->pipe_pending) LV_TRY(lv_pipe_finish(c));       \
    } while (0)
int lv_pipe_enable(LvContext *c);
int lv_pipe_remesh(LvContext *c, int64_t n, const double *xy, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid);
int lv_pipe_wait(LvContext *c);  // deferred remesh completed and every queued download delivered
int lv_pipe_drain(LvContext *c); // every queued download delivered
int lv_pipe_settle(LvContext *c); // queued downloads that declined (open chain) are delivered as full records, now
int lv_pipe_download(LvContext *c, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid);
bool lv_pipe_busy(const LvContext *c);
void lv_pipe_destroy(LvContext *c);
int lv_pipe_upload_begin(LvContext *c, const double *const src[5], const int nc[5], const double *dev[5]);
int lv_pipe_upload_join(LvContext *c);

int lv_ensure(LvContext *c, void **ptr, int64_t *cap, int64_t need, size_t elt); // grow-only device buffer
int lv_alloc(LvContext *c, void **ptr, size_t bytes);
void lv_free(LvContext *c, void *ptr, size_t bytes);
int lv_io_stage(LvContext *c, void **ptr, size_t bytes);
void *lv_mapped_alias(const void *host); // device alias of a pinned host buffer, nullptr for pageable memory

struct LvProfScope { // CUDA-event bracket on the handle's stream; resolved lazily by lv_prof_resolve
    LvContext *c;
    int slot;
    cudaEvent_t a = nullptr;
    LvProfScope(LvContext *c_, int slot_);
    ~LvProfScope();
};
void lv_prof_resolve(LvContext *c); // synchronises the pending event pairs and accumulates them

// ---- phases (each launches on c->stream) -----------------------------------------------------
int lv_cells_build(LvContext *c);                          // K1  lv_cells.cu
int lv_clip_run(LvContext *c);                             // K2  lv_clip.cu
int lv_exclusive_scan_i32(LvContext *c, const int *in, int *out, int64_t n); // out[n] = total
int lv_mesh_to_labels(LvContext *c, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area,
                      double *centroid); // lv_capi.cu
// pressure (lv_pressure.cu)
int lv_pr_ensure(LvContext *c);
int lv_pr_assemble(LvContext *c, double dt);
int lv_pr_matvec(LvContext *c, const double *x, double *y); // slot-order device vectors
int lv_pr_rhs(LvContext *c, double dt, int gp_step, const double *vbc_wall, bool fuse_init, bool *init_done, bool pcg);
int lv_pr_solve(LvContext *c, int solver, double rtol, double atol, int itmax, int *iters, double *relres, bool pre_init);
int lv_pr_find_pressure(LvContext *c, double dt, int niter, double rtol, double atol, int itmax, int solver,
                        const double *vbc_wall, int32_t *iters_out, double *relres_out);
int lv_minres_apply(LvContext *c, int n, const std::function<int(const double *, double *)> &apply, const double *b, double *x,
                    double rtol, double atol, int itmax, int *iters, int *solved);
int lv_gather_to_slots(LvContext *c, const double *src_label_dev, double *dst_slot, int ncomp, double fill);
int lv_scatter_to_labels(LvContext *c, const double *src_slot, double *dst_label_dev, int ncomp);

// allreduce mailbox slot: what one rank posts into another rank's mailbox (lv_dist.cu, lv_pressure.cu)
struct __align__(16) LvMailSlot { double v[2]; int flag; int pad[3]; };
#define LV_MB_MAX_RANKS 64

// multi-GPU (lv_dist.cu): both are no-ops when the handle has no communicator
int lv_halo_exchange(LvContext *c, double *vec_slot, int ncomp); // fill ghost slots from their owners
int lv_allreduce_sum(LvContext *c, double *dev_scalars, int count);
void lv_dist_destroy(LvContext *c);
int lv_halo_exchange_nccl(LvContext *c, double *vec_slot, int ncomp); // pack / ncclSend / ncclRecv / unpack (lv_dist.cu)
// strip exchange over peer memory (lv_strip.cu)
inline bool lv_strip_peer_mode(const LvContext *c) { return c->strip.on && c->strip.mapped; }
int lv_strip_halo_exchange(LvContext *c, double *vec_slot, int ncomp); // pack -> signal -> pull through the exchange areas
int lv_strip_halo_post(LvContext *c, const double *vec_slot, int ncomp);   // pack -> signal (the pull comes later)
int lv_strip_halo_pull(LvContext *c, double *vec_slot);                // pull only (the producer kernel packed and signalled)
void lv_strip_destroy(LvContext *c);
void lv_strip_unmap(LvContext *c);
struct LvHaloPack { // what a producer kernel needs to pack its freshly written values for the neighbours and signal them
    const int *send_pos0, *send_pos1, *bounds; // slot -> outbox position (-1: none); bounds[0..1]: slots outside [b0, b1) may be sent
    double *outbox;                             // this exchange's half of my halo outbox
    int *flag, *ticket;                         // header word the neighbours wait on; last-block ticket
    int seq;                                    // value to publish
};
int lv_strip_pack_args(LvContext *c, LvHaloPack *out); // advances the exchange sequence
int lv_strip_halo_state(LvContext *c, double *field_label_order, int ncomp); // ghosts of a label-ordered field (lv_step.cu state)
int lv_strip_migrate(LvContext *c, int nf, double *const *fields, const int *ncomp); // after a move: leavers travel with their state
int lv_remesh_common(LvContext *c, int64_t n);
int lv_bdry_index(LvContext *c); // numbers the boundary edges of the current mesh (d_bdry_ptr, n_bedge)

// ---- device helpers shared by kernels ---------------------------------------------------------
#define LV_WAIT_TIMEOUT_CYCLES 40000000000ll // ~20 s: only a dead peer gets there
__device__ __forceinline__ int lv_ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lv_st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded wait until *flag >= v (flag may live in a peer's memory).  A peer that never arrives sets *dead, which turns
// every later wait into a no-op; the host then reports an error instead of hanging the GPU.
__device__ __forceinline__ bool lv_wait_ge(const int *flag, int v, int *dead) {
    if (*(volatile int *)dead) return false;
    if (lv_ld_acquire_sys(flag) >= v) return true;
    const long long t0 = clock64();
    for (;;) {
        for (int k = 0; k < 64; k++)
            if (lv_ld_acquire_sys(flag) >= v) return true;
        if (*(volatile int *)dead) return false;
        if (clock64() - t0 > LV_WAIT_TIMEOUT_CYCLES) { *(volatile int *)dead = 1; __threadfence(); return false; }
    }
}
// true in exactly one block of the grid: the one that arrives last.  Every block must call it (all threads).
__device__ __forceinline__ bool lv_last_block(int *ticket) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

__device__ __forceinline__ double lv_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

// y = x + get_arrow(q, x)   voronoigrid.jl:116-125 with the caller's `x + arrow` (voronoigrid.jl:72)
__device__ __forceinline__ double2 lv_neighbor_pos(const LvGridParams &g, double2 x, double2 q) {
    double vx = q.x - x.x, vy = q.y - x.y;
    if (g.xper && (fabs(vx) > 0.5 * g.xperiod)) {
        double s = lv_sign(vx) * g.xperiod;
        vx = vx - s * 1.0;
        vy = vy - s * 0.0;
    }
    if (g.yper && (fabs(vy) > 0.5 * g.yperiod)) {
        double s = lv_sign(vy) * g.yperiod;
        vx = vx - s * 0.0;
        vy = vy - s * 1.0;
    }
    return make_double2(x.x + vx, x.y + vy);
}

// findkey  neighborlist.jl:47-52 (1-based bucket coordinates); false when not representable
__device__ __forceinline__ bool lv_findkey(const LvGridParams &g, double2 x, int &i1, int &i2) {
    double q1 = floor((x.x - g.ox) / g.h), q2 = floor((x.y - g.oy) / g.h);
    if (!isfinite(q1) || !isfinite(q2)) return false; // floor(Int, NaN/Inf) throws InexactError
    // far outside the cell list: any out-of-bounds value behaves the same (checkbounds fails)
    i1 = fabs(q1) < 1.0e9 ? (int)q1 + 1 : (q1 < 0.0 ? -(1 << 30) : (1 << 30));
    i2 = fabs(q2) < 1.0e9 ? (int)q2 + 1 : (q2 < 0.0 ? -(1 << 30) : (1 << 30));
    return true;
}
