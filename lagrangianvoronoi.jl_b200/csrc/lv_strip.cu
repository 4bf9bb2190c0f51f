// lv_strip.cu -- strip decomposition inside the library: ghost-generator exchange, halo plan and halo
// exchange over NVLink peer memory, without NCCL, torch passes or host negotiation.
//
// Per remesh (lv_strip_remesh), on the handle's stream:
//   1. k_strip_select   one pass over the owned generators: those whose primary or y-image bucket row
//                       (voronoigrid.jl:130-147, neighborlist.jl:47-52) falls into a neighbour's window are
//                       appended, per neighbour, to this rank's ghost outbox (x, y, global label);
//   2. k_strip_counts   publishes the counts + "outbox ready" sequence word, waits for the neighbours' words and
//                       reads their counts for me; both go to mapped host memory (ONE host sync per exchange);
//   3. k_strip_pull     copies the neighbours' outboxes (NVLink loads) behind my owned generators;
//   4. cell-list build + clipping on owned + ghosts (buckets ordered by global label => the mesh does not
//                       depend on the number of GPUs);
//   5. k_strip_plan     halo plan on the device: ghost slots to fill, and for every slot a neighbour needs, its
//                       position in my halo outbox.
// Halo values of any slot-ordered vector then travel as pack -> sequence word -> pull; the CG search direction is
// packed by the kernel that produces it (lv_pressure.cu).
//
// Why pulling needs no acknowledgements: every exchange is symmetric (each rank signals and then waits for each of its
// neighbours, also when a count is zero), and outboxes / count words are double-buffered by the parity of the
// sequence number.  A rank that starts exchange e+2 (overwriting parity e) has completed its wait of exchange e+1,
// so every neighbour has signalled e+1, which it does -- in stream order -- after its own pulls of exchange e.
//
// Every device-side wait is bounded (lv_wait_ge): a peer that never arrives raises a flag that turns all later
// waits into no-ops, and the host reports an error instead of hanging the GPU.
#include "lv_internal.cuh"
#include <cstring>

struct LvGhost { double x, y; long long key; }; // 24 B
struct LvStripHdr {
    int gseq;               // remeshes whose ghost outbox + counts are published
    int hseq;               // halo exchanges whose outbox is published
    int mseq;               // migrations whose outbox + counts are published
    int pad0;
    int gcount[2][LV_SP_MAXP]; // [parity][peer index]: ghosts I send to that peer
    int mcount[2][LV_SP_MAXP]; // [parity][peer index]: generators that migrate to that peer
    int pad1[44];
};
static_assert(sizeof(LvStripHdr) == 256, "header is 256 bytes");

static inline size_t sp_off_hbox(int64_t capg) { return sizeof(LvStripHdr) + (size_t)2 * LV_SP_MAXP * (size_t)capg * sizeof(LvGhost); }
#define SP_HB_NC 4   // components per halo value the outbox has room for (D of find_D! has four)
#define SP_MIG_REC 24 // doubles per migrating generator: x (2), global label (1), every resident field (21)
static inline size_t sp_hbox_half(int64_t capg) { return (size_t)LV_SP_MAXP * (size_t)capg * SP_HB_NC * sizeof(double); }
static inline int64_t sp_capm(int64_t capg) { return capg / 8 + 1024; } // migrants per peer and step: a small fraction of the ghost band
static inline size_t sp_off_mbox(int64_t capg) { return sp_off_hbox(capg) + 2 * sp_hbox_half(capg); }
static inline size_t sp_mbox_half(int64_t capg) { return (size_t)LV_SP_MAXP * (size_t)sp_capm(capg) * SP_MIG_REC * sizeof(double); }
static inline size_t sp_area_bytes(int64_t capg) { return sp_off_mbox(capg) + 2 * sp_mbox_half(capg); }

#define sp_ld_acquire lv_ld_acquire_sys
#define sp_st_release lv_st_release_sys

struct SpPeers { // by value to kernels
    char *area[LV_SP_MAXP];
    int idx_there[LV_SP_MAXP];
    int lo[LV_SP_MAXP], hi[LV_SP_MAXP];
    int nsend[LV_SP_MAXP], nrecv[LV_SP_MAXP], soff[LV_SP_MAXP], roff[LV_SP_MAXP];
    int rank[LV_SP_MAXP];
    int np;
};
static SpPeers sp_peers(const LvContext *c) {
    SpPeers P{};
    P.np = c->strip.npeers;
    for (int p = 0; p < P.np; p++) {
        const auto &q = c->strip.peer[p];
        P.area[p] = q.area; P.idx_there[p] = q.idx_there; P.lo[p] = q.lo; P.hi[p] = q.hi; P.rank[p] = q.rank;
        P.nsend[p] = (int)q.nsend; P.nrecv[p] = (int)q.nrecv; P.soff[p] = (int)q.soff; P.roff[p] = (int)q.roff;
    }
    return P;
}

// ---- 1. ghost selection -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_strip_select(SpPeers P, double oy, double h, double yperiod, int yper, int n_own,
                                                      const double2 *__restrict__ xy, const int *__restrict__ key, int capg,
                                                      int *__restrict__ cnt, int *__restrict__ sel, LvGhost *__restrict__ outbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n_own;
    const int lane = threadIdx.x & 31;
    double2 x = make_double2(0.0, 0.0);
    double r0 = -1.0e300, rp = -1.0e300, rm = -1.0e300;
    if (valid) {
        x = xy[i];
        r0 = floor((x.y - oy) / h);                        // findkey  neighborlist.jl:47-52 (0-based row)
        if (yper) {
            rp = floor(((x.y + yperiod * 1.0) - oy) / h);  // the +Y / -Y images of insert_periodic!  voronoigrid.jl:137-138
            rm = floor(((x.y - yperiod * 1.0) - oy) / h);  // (x images keep the row; corner images have these rows)
        }
    }
    for (int p = 0; p < P.np; p++) {
        const double lo = (double)P.lo[p], hi = (double)P.hi[p];
        const bool in = valid && ((r0 >= lo && r0 < hi) || (rp >= lo && rp < hi) || (rm >= lo && rm < hi));
        const unsigned b = __ballot_sync(0xffffffffu, in);
        if (!b) continue;
        const int lead = __ffs(b) - 1;
        int base = 0;
        if (lane == lead) base = atomicAdd(&cnt[p], __popc(b));
        base = __shfl_sync(0xffffffffu, base, lead);
        if (in) {
            const int pos = base + __popc(b & ((1u << lane) - 1u));
            if (pos < capg) {
                sel[(size_t)p * capg + pos] = i;
                LvGhost g; g.x = x.x; g.y = x.y; g.key = key[i];
                outbox[(size_t)p * capg + pos] = g;
            }
        }
    }
}

// ---- 2. counts: publish mine, read the neighbours' ----------------------------------------------------------------
// `mig` selects the migration words (mseq / mcount) instead of the ghost words (gseq / gcount)
__global__ void k_strip_counts(LvStripHdr *me, SpPeers P, int par, int seq, int mig, const int *__restrict__ cnt, int *dead, int *host_out) {
    const int t = threadIdx.x;
    if (t < P.np) (mig ? me->mcount : me->gcount)[par][t] = cnt[t];
    __threadfence_system(); // the outbox written before and the counts, before the sequence word
    __syncwarp();           // every lane's fence has completed before lane 0 publishes
    if (t == 0) sp_st_release(mig ? &me->mseq : &me->gseq, seq);
    if (t < P.np) {
        const LvStripHdr *ph = (const LvStripHdr *)P.area[t];
        const bool ok = lv_wait_ge(mig ? &ph->mseq : &ph->gseq, seq, dead);
        host_out[t] = cnt[t];
        host_out[4 + t] = ok ? __ldcv(&(mig ? ph->mcount : ph->gcount)[par][P.idx_there[t]]) : -1;
    }
    __syncwarp();
    __threadfence_system();
}

// ---- 3. pull the ghosts behind the owned generators ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_strip_pull(SpPeers P, int par, int capg, int n_own, double2 *__restrict__ loc_xy, int *__restrict__ loc_key) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nrecv[p]) return;
    const LvGhost *src = (const LvGhost *)(P.area[p] + sizeof(LvStripHdr)) + ((size_t)par * LV_SP_MAXP + P.idx_there[p]) * (size_t)capg + i;
    const double gx = __ldcv(&src->x), gy = __ldcv(&src->y); // .cv: the same addresses held other ghosts two remeshes ago
    const long long k = __ldcv(&src->key);
    loc_xy[(size_t)n_own + P.roff[p] + i] = make_double2(gx, gy);
    loc_key[(size_t)n_own + P.roff[p] + i] = (int)k;
}

// ---- 5. halo plan ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_strip_plan(SpPeers P, int capg, int n_own, int nslot, const int *__restrict__ sel,
                                                    const int *__restrict__ prim, int *__restrict__ recv_slots, int *__restrict__ send_slots,
                                                    int *send_pos0, int *send_pos1, int *bounds /*[2]*/, int *flags) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.nrecv[p]) {
        const int s = prim[n_own + P.roff[p] + i]; // the one slot per ghost the halo fills (lv_col_of)
        if (s < 0) atomicOr(&flags[LVF_OVERFLOW], 8);
        recv_slots[P.roff[p] + i] = s < 0 ? 0 : s;
    }
    if (i < P.nsend[p]) {
        const int s = prim[sel[(size_t)p * capg + i]];
        send_slots[P.soff[p] + i] = s < 0 ? 0 : s;
        if (s < 0) { atomicOr(&flags[LVF_OVERFLOW], 8); return; }
        const int pos = p * capg + i;
        if (atomicCAS(&send_pos0[s], -1, pos) != -1)
            if (atomicCAS(&send_pos1[s], -1, pos) != -1) atomicOr(&flags[LVF_OVERFLOW], 16); // three neighbours want one slot
        if (s < nslot / 2) atomicMax(&bounds[0], s + 1);
        else atomicMin(&bounds[1], s);
    }
}
__global__ void k_strip_bounds_init(int *bounds, int nslot) { bounds[0] = 0; bounds[1] = nslot; }

// ---- halo values: pack -> sequence word -> pull ----------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) k_hb_pack(SpPeers P, int capg, const int *__restrict__ send_slots, const double *__restrict__ vec,
                                                 double *__restrict__ outbox) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nsend[p]) return;
    const int s = send_slots[P.soff[p] + i];
#pragma unroll
    for (int k = 0; k < NC; k++) outbox[((size_t)p * capg + i) * NC + k] = vec[(size_t)NC * s + k];
}
__global__ void k_hb_signal(LvStripHdr *me, int seq) {
    __threadfence_system();
    sp_st_release(&me->hseq, seq);
}
template <int NC>
__global__ void __launch_bounds__(256) k_hb_pull(SpPeers P, int capg, size_t hbox_off, int seq, const int *__restrict__ recv_slots,
                                                 double *__restrict__ vec, int *dead) {
    const int p = blockIdx.y;
    __shared__ int ok;
    if (threadIdx.x == 0) ok = lv_wait_ge(&((const LvStripHdr *)P.area[p])->hseq, seq, dead) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nrecv[p]) return;
    const double *src = (const double *)(P.area[p] + hbox_off) + ((size_t)P.idx_there[p] * capg + i) * NC;
    const int s = recv_slots[P.roff[p] + i];
#pragma unroll
    for (int k = 0; k < NC; k++) vec[(size_t)NC * s + k] = __ldcv(src + k);
}

static inline int sp_grid(int64_t n) { return (int)((n > 0 ? n : 1) + 255) / 256; }
static int64_t sp_max_send(const LvContext *c) { int64_t m = 0; for (int p = 0; p < c->strip.npeers; p++) m = c->strip.peer[p].nsend > m ? c->strip.peer[p].nsend : m; return m; }
static int64_t sp_max_recv(const LvContext *c) { int64_t m = 0; for (int p = 0; p < c->strip.npeers; p++) m = c->strip.peer[p].nrecv > m ? c->strip.peer[p].nrecv : m; return m; }

int lv_strip_halo_pull(LvContext *c, double *vec) { // NC = 1, exchange number c->strip.hseq (already signalled by the producer)
    auto &S = c->strip;
    if (S.npeers == 0) return LV_OK;
    const int seq = S.hseq;
    const size_t off = sp_off_hbox(S.capg) + (size_t)(seq & 1) * sp_hbox_half(S.capg);
    dim3 grid(sp_grid(sp_max_recv(c)), S.npeers);
    k_hb_pull<1><<<grid, 256, 0, c->stream>>>(sp_peers(c), (int)S.capg, off, seq, c->d_recv_slots, vec, c->d_tickets + 7);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// pack + signal only: the matching lv_strip_halo_pull comes later (ncomp = 1)
int lv_strip_halo_post(LvContext *c, const double *vec, int ncomp) {
    auto &S = c->strip;
    if (S.npeers == 0) return LV_OK;
    const int seq = ++S.hseq;
    const size_t off = sp_off_hbox(S.capg) + (size_t)(seq & 1) * sp_hbox_half(S.capg);
    dim3 gs(sp_grid(sp_max_send(c)), S.npeers);
    double *out = (double *)(S.area + off);
    if (ncomp == 1) k_hb_pack<1><<<gs, 256, 0, c->stream>>>(sp_peers(c), (int)S.capg, c->d_send_slots, vec, out);
    else k_hb_pack<2><<<gs, 256, 0, c->stream>>>(sp_peers(c), (int)S.capg, c->d_send_slots, vec, out);
    k_hb_signal<<<1, 1, 0, c->stream>>>((LvStripHdr *)S.area, seq);
    c->launches += 2;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int lv_strip_halo_exchange(LvContext *c, double *vec, int ncomp) {
    auto &S = c->strip;
    if (S.npeers == 0) return LV_OK;
    LV_TRY(lv_strip_halo_post(c, vec, ncomp));
    const int seq = S.hseq;
    const size_t off = sp_off_hbox(S.capg) + (size_t)(seq & 1) * sp_hbox_half(S.capg);
    const SpPeers P = sp_peers(c);
    dim3 gr(sp_grid(sp_max_recv(c)), S.npeers);
    if (ncomp == 1) k_hb_pull<1><<<gr, 256, 0, c->stream>>>(P, (int)S.capg, off, seq, c->d_recv_slots, vec, c->d_tickets + 7);
    else k_hb_pull<2><<<gr, 256, 0, c->stream>>>(P, (int)S.capg, off, seq, c->d_recv_slots, vec, c->d_tickets + 7);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// arguments for a producer kernel that packs + signals exchange number hseq + 1 itself (lv_pressure.cu: k_cg_update_xp)
int lv_strip_pack_args(LvContext *c, LvHaloPack *out) {
    auto &S = c->strip;
    const int seq = ++S.hseq;
    out->send_pos0 = S.send_pos0; out->send_pos1 = S.send_pos1; out->bounds = S.cnt + 4;
    out->outbox = (double *)(S.area + sp_off_hbox(S.capg) + (size_t)(seq & 1) * sp_hbox_half(S.capg));
    out->flag = &((LvStripHdr *)S.area)->hseq;
    out->ticket = c->d_tickets + 2;
    out->seq = seq;
    return LV_OK;
}

// ---- halo of a LABEL-ordered field (device-resident state of lv_step.cu): owned generators sit at [0, n_own), the ghosts
// received from peer p at n_own + roff[p] + i in the order of that peer's send list -> no slot lists are needed
__global__ void __launch_bounds__(256) k_hs_pack(SpPeers P, int capg, int nc, const int *__restrict__ sel, const double *__restrict__ field,
                                                 double *__restrict__ outbox) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nsend[p]) return;
    const int src = sel[(size_t)p * capg + i];
    for (int k = 0; k < nc; k++) outbox[((size_t)p * capg + i) * nc + k] = field[(size_t)nc * src + k];
}
__global__ void __launch_bounds__(256) k_hs_pull(SpPeers P, int capg, int nc, size_t hbox_off, int seq, int n_own, double *__restrict__ field, int *dead) {
    const int p = blockIdx.y;
    __shared__ int ok;
    if (threadIdx.x == 0) ok = lv_wait_ge(&((const LvStripHdr *)P.area[p])->hseq, seq, dead) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nrecv[p]) return;
    const double *src = (const double *)(P.area[p] + hbox_off) + ((size_t)P.idx_there[p] * capg + i) * nc;
    double *dst = field + ((size_t)n_own + P.roff[p] + i) * nc;
    for (int k = 0; k < nc; k++) dst[k] = __ldcv(src + k);
}
int lv_strip_halo_state(LvContext *c, double *field, int nc) {
    auto &S = c->strip;
    if (!lv_strip_peer_mode(c) || S.npeers == 0) return LV_OK;
    if (nc < 1 || nc > SP_HB_NC) return lv_set_error(c, LV_EINVAL, "halo of %d components", nc);
    const int seq = ++S.hseq;
    const size_t off = sp_off_hbox(S.capg) + (size_t)(seq & 1) * sp_hbox_half(S.capg);
    const SpPeers P = sp_peers(c);
    dim3 gs(sp_grid(sp_max_send(c)), S.npeers), gr(sp_grid(sp_max_recv(c)), S.npeers);
    k_hs_pack<<<gs, 256, 0, c->stream>>>(P, (int)S.capg, nc, S.sel, field, (double *)(S.area + off));
    k_hb_signal<<<1, 1, 0, c->stream>>>((LvStripHdr *)S.area, seq);
    k_hs_pull<<<gr, 256, 0, c->stream>>>(P, (int)S.capg, nc, off, seq, (int)S.n_own, field, c->d_tickets + 7);
    c->launches += 3;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// ---- migration: generators that left the strip take their state to the new owner ---------------------------------------
struct SpRows { int R[LV_MB_MAX_RANKS + 1]; int world, rank; double oy, h; }; // bucket rows [R[r], R[r+1]) belong to rank r
struct SpFields { double *ptr[LV_STATE_MAX]; int nc[LV_STATE_MAX]; int nf; };

__global__ void __launch_bounds__(256) k_mig_select(SpPeers P, SpRows W, int n_own, const double2 *__restrict__ xy, int capm,
                                                    int *__restrict__ cnt, int *__restrict__ mig, unsigned char *__restrict__ leave, int *flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const double q = floor((xy[i].y - W.oy) / W.h); // findkey  neighborlist.jl:47-52
    int owner = 0;
    while (owner + 1 < W.world && q >= (double)W.R[owner + 1]) owner++;
    if (owner == W.rank || !(q == q)) { leave[i] = 0; return; }
    int p = -1;
    for (int k = 0; k < P.np; k++)
        if (P.rank[k] == owner) p = k;
    if (p < 0) { atomicOr(&flags[LVF_OVERFLOW], 32); leave[i] = 0; return; } // jumped over a whole strip in one step
    leave[i] = 1;
    const int pos = atomicAdd(&cnt[p], 1);
    if (pos < capm) mig[(size_t)p * capm + pos] = i;
}
__global__ void __launch_bounds__(128) k_mig_pack(SpPeers P, int capm, SpFields F, const double2 *__restrict__ xy, const int *__restrict__ key,
                                                  const int *__restrict__ mig, const int *__restrict__ cnt, double *__restrict__ outbox) {
    const int p = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = cnt[p] < capm ? cnt[p] : capm;
    if (k >= m) return;
    const int i = mig[(size_t)p * capm + k];
    double *rec = outbox + ((size_t)p * capm + k) * SP_MIG_REC;
    rec[0] = xy[i].x; rec[1] = xy[i].y; rec[2] = (double)key[i];
    int o = 3;
    for (int f = 0; f < F.nf; f++)
        for (int cpt = 0; cpt < F.nc[f]; cpt++) rec[o++] = F.ptr[f][(size_t)F.nc[f] * i + cpt];
}
// holes = leavers below n_stay, tails = stayers at or above n_stay; there are equally many of both
__global__ void __launch_bounds__(256) k_mig_lists(int n_stay, int n_own, const unsigned char *__restrict__ leave, int *__restrict__ holes,
                                                   int *__restrict__ tails, int *__restrict__ cnt2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    if (i < n_stay && leave[i]) holes[atomicAdd(&cnt2[0], 1)] = i;
    if (i >= n_stay && !leave[i]) tails[atomicAdd(&cnt2[1], 1)] = i;
}
__global__ void __launch_bounds__(128) k_mig_fill_bounded(int nmax, const int *__restrict__ cnt2, const int *__restrict__ holes, const int *__restrict__ tails,
                                                          SpFields F, double2 *__restrict__ xy, int *__restrict__ key) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nmax || j >= cnt2[0]) return; // cnt2[0] = number of holes (= cnt2[1], the number of tails)
    const int dst = holes[j], src = tails[j];
    xy[dst] = xy[src];
    key[dst] = key[src];
    for (int f = 0; f < F.nf; f++)
        for (int cpt = 0; cpt < F.nc[f]; cpt++) F.ptr[f][(size_t)F.nc[f] * dst + cpt] = F.ptr[f][(size_t)F.nc[f] * src + cpt];
}
struct SpMigIn { int n[LV_SP_MAXP], off[LV_SP_MAXP]; };
__global__ void __launch_bounds__(128) k_mig_pull(SpPeers P, SpMigIn in, int par, int capm, size_t mbox_off, size_t mbox_half, int n_stay, SpFields F,
                                                  double2 *__restrict__ xy, int *__restrict__ key) {
    const int p = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= in.n[p]) return;
    const double *rec = (const double *)(P.area[p] + mbox_off + (size_t)par * mbox_half) + ((size_t)P.idx_there[p] * capm + k) * SP_MIG_REC;
    const int dst = n_stay + in.off[p] + k;
    xy[dst] = make_double2(__ldcv(rec), __ldcv(rec + 1));
    key[dst] = (int)__ldcv(rec + 2);
    int o = 3;
    for (int f = 0; f < F.nf; f++)
        for (int cpt = 0; cpt < F.nc[f]; cpt++) F.ptr[f][(size_t)F.nc[f] * dst + cpt] = __ldcv(rec + o++);
}

// After positions moved: every owned generator whose bucket row now belongs to another rank travels there with all its
// resident fields (fields: label-ordered device arrays with `nc` components each); the owned set stays contiguous at the
// front of the local arrays (order is irrelevant: buckets are ordered by global label).  Collective over the strip
// neighbours; one host synchronisation.  The mesh is invalid afterwards (lv_strip_remesh follows).
int lv_strip_migrate(LvContext *c, int nf, double *const *fields, const int *ncomp) {
    auto &S = c->strip;
    if (!S.on) return lv_set_error(c, LV_EINVAL, "lv_strip_migrate: not in strip mode");
    if (S.npeers == 0) return LV_OK;
    if (!S.mapped) return lv_set_error(c, LV_EINVAL, "lv_strip_migrate: peers not mapped");
    if (S.world_rows.empty()) return lv_set_error(c, LV_EINVAL, "lv_strip_migrate: strip rows unknown (lv_strip_set_rows)");
    SpFields F{};
    int rec = 3;
    if (nf > LV_STATE_MAX) return lv_set_error(c, LV_EINVAL, "too many fields");
    F.nf = nf;
    for (int f = 0; f < nf; f++) { F.ptr[f] = fields[f]; F.nc[f] = ncomp[f]; rec += ncomp[f]; }
    if (rec > SP_MIG_REC) return lv_set_error(c, LV_EINVAL, "migration record of %d doubles exceeds %d", rec, SP_MIG_REC);
    cudaStream_t st = c->stream;
    const int capm = (int)sp_capm(S.capg);
    if (!S.mig) {
        LV_CUDA(c, cudaMalloc((void **)&S.mig, sizeof(int) * (size_t)LV_SP_MAXP * capm));
        LV_CUDA(c, cudaMalloc((void **)&S.holes, sizeof(int) * (size_t)LV_SP_MAXP * capm * 2));
        LV_CUDA(c, cudaMalloc((void **)&S.leave, (size_t)S.cap_loc));
    }
    SpRows W{};
    W.world = (int)S.world_rows.size() - 1; W.rank = c->rank; W.oy = c->gp.oy; W.h = c->gp.h;
    for (size_t k = 0; k < S.world_rows.size(); k++) W.R[k] = S.world_rows[k];
    const int seq = ++S.mseq, par = seq & 1;
    SpPeers P = sp_peers(c);
    double *outbox = (double *)(S.area + sp_off_mbox(S.capg) + (size_t)par * sp_mbox_half(S.capg));
    LV_CUDA(c, cudaMemsetAsync(S.cnt, 0, sizeof(int) * 8, st));
    LV_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 8, st));
    const int n_own = (int)S.n_own;
    if (n_own > 0) k_mig_select<<<sp_grid(n_own), 256, 0, st>>>(P, W, n_own, S.loc_xy, capm, S.cnt, S.mig, S.leave, c->d_flags);
    {
        dim3 g((capm + 127) / 128, S.npeers);
        k_mig_pack<<<g, 128, 0, st>>>(P, capm, F, S.loc_xy, S.loc_key, S.mig, S.cnt, outbox);
    }
    k_strip_counts<<<1, 32, 0, st>>>((LvStripHdr *)S.area, P, par, seq, 1, S.cnt, c->d_tickets + 7, S.h_counts);
    c->launches += 3;
    LV_TRY(lv_publish_flags(c, nullptr)); // synchronises
    if (c->h_flags[LVF_OVERFLOW] & 32) return lv_set_error(c, LV_EINVAL, "strip: a generator moved past a neighbouring strip in one step");
    SpMigIn in{};
    int out_total = 0, in_total = 0;
    for (int p = 0; p < S.npeers; p++) {
        const int ns = S.h_counts[p], nr = S.h_counts[4 + p];
        if (nr < 0) return lv_set_error(c, LV_ECUDA, "strip: rank %d did not publish its migrants (peer timeout)", S.peer[p].rank);
        if (ns > capm || nr > capm) return lv_set_error(c, LV_ECAPACITY, "strip: %d / %d migrants exceed the capacity %d", ns, nr, capm);
        in.n[p] = nr; in.off[p] = in_total;
        out_total += ns; in_total += nr;
    }
    const int n_stay = n_own - out_total;
    if ((int64_t)n_stay + in_total > S.cap_loc) return lv_set_error(c, LV_ECAPACITY, "strip: owned generators exceed the local capacity after migration");
    if (out_total > 0) {
        LV_CUDA(c, cudaMemsetAsync(S.cnt, 0, sizeof(int) * 8, st));
        k_mig_lists<<<sp_grid(n_own), 256, 0, st>>>(n_stay, n_own, S.leave, S.holes, S.holes + (size_t)LV_SP_MAXP * capm, S.cnt);
        // #holes = #leavers below n_stay <= out_total; the lists are complete when the kernel is: launch with the upper bound
        k_mig_fill_bounded<<<(out_total + 127) / 128, 128, 0, st>>>(out_total, S.cnt, S.holes, S.holes + (size_t)LV_SP_MAXP * capm, F, S.loc_xy, S.loc_key);
        c->launches += 2;
    }
    if (in_total > 0) {
        int mx = 0;
        for (int p = 0; p < S.npeers; p++) mx = in.n[p] > mx ? in.n[p] : mx;
        dim3 g((mx + 127) / 128, S.npeers);
        k_mig_pull<<<g, 128, 0, st>>>(P, in, par, capm, sp_off_mbox(S.capg), sp_mbox_half(S.capg), n_stay, F, S.loc_xy, S.loc_key);
        c->launches++;
    }
    LV_CUDA(c, cudaGetLastError());
    S.n_own = n_stay + in_total;
    S.last_mig_out = out_total; S.last_mig_in = in_total;
    c->mesh_valid = false;
    return LV_OK;
}

void lv_strip_unmap(LvContext *c) {
    auto &S = c->strip;
    for (int p = 0; p < S.npeers; p++)
        if (S.peer[p].area) { cudaIpcCloseMemHandle(S.peer[p].area); S.peer[p].area = nullptr; }
    S.mapped = false;
}
void lv_strip_destroy(LvContext *c) {
    auto &S = c->strip;
    lv_strip_unmap(c);
    cudaFree(S.area); cudaFree(S.loc_xy); cudaFree(S.loc_key); cudaFree(S.sel); cudaFree(S.cnt);
    cudaFree(S.send_pos0); cudaFree(S.send_pos1); cudaFree(S.mig); cudaFree(S.holes); cudaFree(S.leave);
    if (S.h_counts) cudaFreeHost(S.h_counts);
    S = LvContext::Strip();
}

extern "C" {

// Allocate the strip state: peers (rank, my index in its peer list, the bucket-row window [lo, hi) it needs), the ghost
// capacity per peer (the same number on every rank: it fixes the layout of the exchange areas) and the capacity of the
// local generator arrays.  out64 receives the CUDA IPC handle of this rank's exchange area.
int32_t lv_strip_setup(LvHandle c, int32_t npeers, const int32_t *peer_rank, const int32_t *idx_there, const int32_t *lo,
                       const int32_t *hi, int64_t capg, int64_t cap_loc, uint8_t *out64) {
    if (!c || npeers < 0 || npeers > LV_SP_MAXP || capg < 1 || cap_loc < 1 || !out64) return lv_set_error(c, LV_EINVAL, "lv_strip_setup: bad arguments");
    if (capg * LV_SP_MAXP >= ((int64_t)1 << 30)) return lv_set_error(c, LV_EINVAL, "lv_strip_setup: ghost capacity too large");
    LV_ENTER(c);
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    lv_strip_destroy(c);
    auto &S = c->strip;
    S.npeers = npeers;
    for (int p = 0; p < npeers; p++) {
        S.peer[p] = LvContext::StripPeer();
        S.peer[p].rank = peer_rank[p]; S.peer[p].idx_there = idx_there[p]; S.peer[p].lo = lo[p]; S.peer[p].hi = hi[p];
    }
    S.capg = capg; S.cap_loc = cap_loc;
    S.area_bytes = sp_area_bytes(capg);
    LV_CUDA(c, cudaMalloc((void **)&S.area, S.area_bytes));
    LV_CUDA(c, cudaMemset(S.area, 0, sizeof(LvStripHdr)));
    LV_CUDA(c, cudaMalloc((void **)&S.loc_xy, sizeof(double2) * (size_t)cap_loc));
    LV_CUDA(c, cudaMalloc((void **)&S.loc_key, sizeof(int) * (size_t)cap_loc));
    LV_CUDA(c, cudaMalloc((void **)&S.sel, sizeof(int) * (size_t)LV_SP_MAXP * (size_t)capg));
    LV_CUDA(c, cudaMalloc((void **)&S.cnt, sizeof(int) * 8));
    LV_CUDA(c, cudaMemset(S.cnt, 0, sizeof(int) * 8));
    LV_CUDA(c, cudaHostAlloc((void **)&S.h_counts, sizeof(int) * 16, cudaHostAllocMapped | cudaHostAllocPortable));
    c->dev_bytes += (int64_t)(S.area_bytes + (sizeof(double2) + sizeof(int)) * (size_t)cap_loc + sizeof(int) * LV_SP_MAXP * (size_t)capg);
    cudaIpcMemHandle_t hm;
    LV_CUDA(c, cudaIpcGetMemHandle(&hm, S.area));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &hm, 64);
    S.on = true;
    S.gseq = S.hseq = 0;
    return LV_OK;
}

// Map the exchange areas of the peers (handles in the order of lv_strip_setup's peer list).
int32_t lv_strip_map(LvHandle c, const uint8_t *handles /* npeers x 64 */) {
    if (!c || !c->strip.on || (c->strip.npeers > 0 && !handles)) return lv_set_error(c, LV_EINVAL, "lv_strip_map: call lv_strip_setup first");
    LV_ENTER(c);
    auto &S = c->strip;
    lv_strip_unmap(c);
    for (int p = 0; p < S.npeers; p++) {
        // two peer entries may name the same rank only if the caller merged them; each rank appears once
        cudaIpcMemHandle_t hm;
        memcpy(&hm, handles + 64 * p, 64);
        void *base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, hm, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            lv_strip_unmap(c);
            return lv_set_error(c, LV_ECUDA, "cudaIpcOpenMemHandle(exchange area of rank %d) failed: %s", S.peer[p].rank, cudaGetErrorString(e));
        }
        S.peer[p].area = (char *)base;
    }
    S.mapped = true;
    return LV_OK;
}

// bucket-row ownership of all ranks, R[0..world]: rank r owns rows [R[r], R[r+1]) -- what migration needs to find new owners
int32_t lv_strip_set_rows(LvHandle c, int32_t world, const int32_t *R) {
    if (!c || !c->strip.on || world < 1 || world > LV_MB_MAX_RANKS || !R) return lv_set_error(c, LV_EINVAL, "lv_strip_set_rows: bad arguments");
    c->strip.world_rows.assign(R, R + world + 1);
    return LV_OK;
}

// Owned generators of this rank: positions (host or device memory) and their global labels (device int32, may be NULL
// to keep the previous ones), copied to the front of the library's local arrays.
int32_t lv_strip_set_owned(LvHandle c, int64_t n_own, const double *xy, int32_t xy_on_host, const int32_t *key_dev) {
    if (!c || !c->strip.on || n_own < 0 || (n_own > 0 && !xy)) return lv_set_error(c, LV_EINVAL, "lv_strip_set_owned: bad arguments");
    auto &S = c->strip;
    if (n_own + (int64_t)S.npeers * 0 > S.cap_loc) return lv_set_error(c, LV_ECAPACITY, "strip: %lld owned generators exceed the capacity %lld", (long long)n_own, (long long)S.cap_loc);
    LV_ENTER(c);
    if (n_own > 0) {
        LV_CUDA(c, cudaMemcpyAsync(S.loc_xy, xy, sizeof(double2) * (size_t)n_own, xy_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
        if (key_dev) LV_CUDA(c, cudaMemcpyAsync(S.loc_key, key_dev, sizeof(int) * (size_t)n_own, cudaMemcpyDeviceToDevice, c->stream));
    }
    S.n_own = n_own;
    return LV_OK;
}

// remesh!(grid) on a strip: ghost exchange, cell list + clipping of the owned polygons, halo plan.  counts_out (nullable):
// [n_loc, nsend[4], nrecv[4]].
int32_t lv_strip_remesh(LvHandle c, int64_t *counts_out) {
    if (!c || !c->strip.on) return lv_set_error(c, LV_EINVAL, "lv_strip_remesh: call lv_strip_setup first");
    auto &S = c->strip;
    if (S.npeers > 0 && !S.mapped) return lv_set_error(c, LV_EINVAL, "lv_strip_remesh: peers not mapped");
    LV_ENTER(c);
    cudaStream_t st = c->stream;
    LvStripHdr *me = (LvStripHdr *)S.area;
    int64_t nrecv_total = 0;
    if (S.npeers > 0) {
        LvProfScope prof(c, LV_PROF_CELLS);
        const int seq = ++S.gseq, par = seq & 1;
        LV_CUDA(c, cudaMemsetAsync(S.cnt, 0, sizeof(int) * 4, st));
        SpPeers P = sp_peers(c);
        LvGhost *outbox = (LvGhost *)(S.area + sizeof(LvStripHdr)) + (size_t)par * LV_SP_MAXP * (size_t)S.capg;
        k_strip_select<<<sp_grid(S.n_own), 256, 0, st>>>(P, c->gp.oy, c->gp.h, c->gp.yperiod, c->gp.yper, (int)S.n_own, S.loc_xy, S.loc_key,
                                                        (int)S.capg, S.cnt, S.sel, outbox);
        k_strip_counts<<<1, 32, 0, st>>>(me, P, par, seq, 0, S.cnt, c->d_tickets + 7, S.h_counts);
        c->launches += 2;
        LV_CUDA(c, cudaGetLastError());
        LV_CUDA(c, cudaStreamSynchronize(st));
        int64_t so = 0, ro = 0;
        for (int p = 0; p < S.npeers; p++) {
            const int ns = S.h_counts[p], nr = S.h_counts[4 + p];
            if (nr < 0) return lv_set_error(c, LV_ECUDA, "strip: rank %d did not publish its ghosts (peer timeout)", S.peer[p].rank);
            if (ns > S.capg || nr > S.capg)
                return lv_set_error(c, LV_ECAPACITY, "strip: %d ghosts to / %d from rank %d exceed the ghost capacity %lld", ns, nr, S.peer[p].rank, (long long)S.capg);
            S.peer[p].nsend = ns; S.peer[p].nrecv = nr; S.peer[p].soff = so; S.peer[p].roff = ro;
            so += ns; ro += nr;
        }
        nrecv_total = ro;
        if (S.n_own + ro > S.cap_loc)
            return lv_set_error(c, LV_ECAPACITY, "strip: %lld owned + %lld ghost generators exceed the capacity %lld", (long long)S.n_own, (long long)ro, (long long)S.cap_loc);
        // halo bookkeeping shared with the NCCL fallback (c->peers, send / recv slot lists)
        c->peers.clear();
        for (int p = 0; p < S.npeers; p++) c->peers.push_back({S.peer[p].rank, S.peer[p].nsend, S.peer[p].nrecv, S.peer[p].soff, S.peer[p].roff});
        c->halo_send_total = so; c->halo_recv_total = ro;
        if (so > c->cap_halo_send || !c->d_send_slots) {
            cudaFree(c->d_send_slots); cudaFree(c->d_send_buf);
            c->cap_halo_send = (int64_t)S.npeers * S.capg + 16;
            LV_CUDA(c, cudaMalloc((void **)&c->d_send_slots, sizeof(int) * (size_t)c->cap_halo_send));
            LV_CUDA(c, cudaMalloc((void **)&c->d_send_buf, sizeof(double) * 2 * (size_t)c->cap_halo_send));
        }
        if (ro > c->cap_halo_recv || !c->d_recv_slots) {
            cudaFree(c->d_recv_slots); cudaFree(c->d_recv_buf);
            c->cap_halo_recv = (int64_t)S.npeers * S.capg + 16;
            LV_CUDA(c, cudaMalloc((void **)&c->d_recv_slots, sizeof(int) * (size_t)c->cap_halo_recv));
            LV_CUDA(c, cudaMalloc((void **)&c->d_recv_buf, sizeof(double) * 2 * (size_t)c->cap_halo_recv));
        }
        if (ro > 0) {
            P = sp_peers(c);
            dim3 grid(sp_grid(sp_max_recv(c)), S.npeers);
            k_strip_pull<<<grid, 256, 0, st>>>(P, par, (int)S.capg, (int)S.n_own, S.loc_xy, S.loc_key);
            c->launches++;
        }
    }
    S.n_loc = S.n_own + nrecv_total;
    // cell list + clipping of owned + ghosts; buckets ordered by global label, rows only for the owned generators
    {
        const int64_t n = S.n_loc;
        if (n >= ((int64_t)1 << 30)) return lv_set_error(c, LV_EINVAL, "n = %lld out of range (limit 2^30)", (long long)n);
        if (!c->d_prim_of_label || c->cap_n < S.cap_loc) {
            if (c->d_xy) { LV_CUDA(c, cudaStreamSynchronize(st)); lv_free(c, c->d_xy, sizeof(double2) * (size_t)c->cap_n); c->d_xy = nullptr; }
            int64_t cap = c->d_prim_of_label ? c->cap_n : 0;
            LV_TRY(lv_ensure(c, (void **)&c->d_prim_of_label, &cap, S.cap_loc, sizeof(int)));
            c->cap_n = S.cap_loc;
        }
        c->xy = S.loc_xy;
        c->owned_mask = nullptr;
        c->order_key = S.loc_key;
        c->owned_count = S.n_own;
        int stt = lv_remesh_common(c, n);
        c->order_key = nullptr;
        c->owned_count = -1;
        LV_TRY(stt);
    }
    if (S.npeers > 0) {
        LvProfScope prof(c, LV_PROF_CELLS);
        if (c->cap_slot > S.cap_pos || !S.send_pos0) {
            LV_CUDA(c, cudaStreamSynchronize(st));
            cudaFree(S.send_pos0); cudaFree(S.send_pos1);
            S.cap_pos = c->cap_slot;
            LV_CUDA(c, cudaMalloc((void **)&S.send_pos0, sizeof(int) * (size_t)S.cap_pos));
            LV_CUDA(c, cudaMalloc((void **)&S.send_pos1, sizeof(int) * (size_t)S.cap_pos));
        }
        LV_CUDA(c, cudaMemsetAsync(S.send_pos0, 0xff, sizeof(int) * (size_t)c->nslot, st));
        LV_CUDA(c, cudaMemsetAsync(S.send_pos1, 0xff, sizeof(int) * (size_t)c->nslot, st));
        k_strip_bounds_init<<<1, 1, 0, st>>>(S.cnt + 4, (int)c->nslot);
        const int64_t m = sp_max_send(c) > sp_max_recv(c) ? sp_max_send(c) : sp_max_recv(c);
        dim3 grid(sp_grid(m), S.npeers);
        k_strip_plan<<<grid, 256, 0, st>>>(sp_peers(c), (int)S.capg, (int)S.n_own, (int)c->nslot, S.sel, c->d_prim_of_label, c->d_recv_slots,
                                           c->d_send_slots, S.send_pos0, S.send_pos1, S.cnt + 4, c->d_flags);
        c->launches += 2;
        LV_TRY(lv_publish_flags(c, nullptr));
        if (c->h_flags[LVF_OVERFLOW] & 8) return lv_set_error(c, LV_EINVAL, "strip: a generator on the halo plan has no slot in the local cell list");
        if (c->h_flags[LVF_OVERFLOW] & 16) return lv_set_error(c, LV_EINVAL, "strip: a slot is needed by more than two neighbours (strips too thin)");
    }
    if (counts_out) {
        counts_out[0] = S.n_loc;
        for (int p = 0; p < LV_SP_MAXP; p++) { counts_out[1 + p] = p < S.npeers ? S.peer[p].nsend : 0; counts_out[5 + p] = p < S.npeers ? S.peer[p].nrecv : 0; }
    }
    return LV_OK;
}

} // extern "C"
