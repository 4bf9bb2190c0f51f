// lv_clip_fast.cu -- K2, production kernel: linked-slot half-plane clipping.
//
// Same results, bit for bit, as the edge-list kernel in lv_clip.cu (and therefore as the
// reference, polygon.jl:51-97 / voronoigrid.jl:53-81 / IO.jl:35-48), at a fraction of the
// instructions:
//
//  * a polygon is a ring of "slots" in shared memory, [slot][thread] layout, 20 B per slot:
//    the start vertex v1 (double2) and the label (int).  The end vertex of an edge is the start
//    vertex of its successor, so every vertex is stored and tested ONCE per cut (the reference
//    evaluates f at both ends of every edge);
//  * all bookkeeping lives in registers, 4 bits per slot: `nxt`/`prv` (ring links), `ord`
//    (the reference's FastVector storage order, which decides which edge starts the sorted
//    chain and which intersection the reference calls "X") and its inverse `pos`, plus a
//    16-bit occupancy mask and two 16-bit sign masks;
//  * a cut first classifies every vertex (f > eps, |f| <= eps, f < -eps); if no vertex is
//    outside nothing else happens.  Otherwise the outside vertices must form one arc of the
//    ring; the edge entering the arc and the edge leaving it are the two edges the reference
//    cuts, everything in between is what it deletes.  Its swap-remove deletions
//    (fastvector.jl:44-50) are replayed on the 4-bit storage order only when something is
//    deleted; the two intersection points use the reference's expression verbatim;
//  * the neighbour walk (voronoigrid.jl:57-79) is split in three warp-synchronous phases:
//    (A) every lane scans a fixed number of candidates / path nodes ahead with the current
//    influence radius and queues the ones that pass the distance filter, (B) lanes pop queued
//    candidates -- re-applying, with the radius of that moment, the node-entry test and the
//    distance filter exactly where the reference applies them -- until one has a vertex
//    outside, (C) the cut.  Scanning ahead is exact because the influence radius never grows
//    (checked; a violation is reported as an anomaly);
//  * tiles are warps and rows are placed with one atomicAdd per warp: no barrier and no
//    ordering between warps.  The CSR is (row start, degree) per slot; rows of one warp are
//    contiguous, warps land in completion order.
//
// The kernel proves, per cut, that it is in the generic situation (one contiguous outside arc,
// the reference's orientation test agrees with the ring, distinct finite intersection points,
// non-increasing influence radius, no zero-length edge in the final ring).  Anything else --
// only degenerate inputs get there -- raises OVF_ANOMALY and lv_clip_run repeats the remesh
// with the edge-list kernel, which replays the reference literally.  Results are therefore
// always the reference's.
#include "lv_clip.cuh"

typedef unsigned long long u64;

#define QCAP 4 // queued candidates per lane
#define EV 4   // scan events per lane and round
#define QLOW 1 // a scan pass starts (and goes on) only while some lane has at most this many candidates queued
#define BMAX 2 // candidates a lane may pop per round before the warp moves on to the cut

template <int MAXE, int BLOCK>
struct Ring {
    double2 *sv;
    int *sl;
    u64 ord, pos, nxt, prv;
    unsigned used;
    int m;
    __device__ __forceinline__ double2 &V(int s) { return sv[s * BLOCK + threadIdx.x]; }
    __device__ __forceinline__ int &L(int s) { return sl[s * BLOCK + threadIdx.x]; }
};

__device__ __forceinline__ int nib(u64 w, int i) { return (int)((w >> (4 * i)) & 15ull); }
__device__ __forceinline__ u64 setnib(u64 w, int i, int v) { return (w & ~(15ull << (4 * i))) | ((u64)v << (4 * i)); }

template <int MAXE, int BLOCK>
__device__ __forceinline__ double ring_influence_rr(Ring<MAXE, BLOCK> &p, double2 x) { // polygon.jl:101-107
    double rr = 0.0;
    for (unsigned u = p.used; u; u &= u - 1) {
        const int s = __ffs(u) - 1;
        const double2 a = p.V(s);
        const double ex = a.x - x.x, ey = a.y - x.y;
        const double t = 4.0 * (ex * ex + ey * ey);
        rr = (t > rr || isnan(t)) ? t : rr; // NaN propagates like Julia's max (NaN then fails every test -> anomaly)
    }
    return rr;
}

// LV_CLIP_STATS=1 (diagnostics): per-phase warp iterations and active lanes of the tile kernel, printed to stderr
#define LV_STAT(...) do { if (STATS) { __VA_ARGS__ } } while (0)
template <int MAXE, int BLOCK, int MINB, bool STATS = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_clip_fast(ClipArgs a, int ntiles, unsigned long long *stats = nullptr) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *sv = (double2 *)smem;
    int *sl = (int *)(sv + MAXE * BLOCK);
    int *sq = sl + MAXE * BLOCK;      // [QCAP][BLOCK] queued candidate slots
    int *sqt = sq + QCAP * BLOCK;     // [QCAP][BLOCK] their path-node indices
    LvPathNode *spath = (LvPathNode *)(sqt + QCAP * BLOCK);
    const LvGridParams g = a.g;
    for (int k = threadIdx.x; k < g.npath; k += BLOCK) spath[k] = a.path[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    unsigned long long st_a_it = 0, st_a_ln = 0, st_b_it = 0, st_b_ln = 0, st_c_it = 0, st_c_ln = 0, st_rounds = 0, st_alive = 0;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(&a.flags[LVF_TICKET], 1);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= ntiles) break;
        const int slot = tile * 32 + lane;

        Ring<MAXE, BLOCK> p;
        p.sv = sv; p.sl = sl; p.m = 0; p.used = 0; p.ord = p.pos = p.nxt = p.prv = 0;
        bool alive = false;
        double2 x = make_double2(0.0, 0.0);
        if (slot < a.nslot) {
            alive = a.own[slot] != 0;
            x = a.ent_xy[slot];
        }
        const bool active = alive;
        bool bad = false; // anomaly or capacity overflow for this polygon
        double prr = 0.0;
        int k1 = 0, k2 = 0;
        int t_scan = -1, s = 0, s_end = 0, t_end = g.npath, t_chk = -1; // scan cursor / last node whose entry test ran
        bool scan_done = !active;
        int qh = 0, qn = 0; // queue = positions [qh, qn) modulo QCAP
        const double hpx = 0.5 * g.xperiod, hpy = 0.5 * g.yperiod;
        const double slack = 2e-15 * (fabs(x.x) + fabs(x.y)); // 2 * rounding bound of x - (x + v), see phase A
        // The periodic shift of phase A can only fire for a candidate more than half a period away, i.e. for a polygon
        // within the reach of the path table (wreach = (max |i| + 1) h) of the domain edge: warp-uniform switch.
        const bool wrapx = g.xper && __any_sync(FULL, active && !(x.x - a.wlo.x > a.wreach && a.whi.x - x.x > a.wreach && a.wreach < hpx));
        const bool wrapy = g.yper && __any_sync(FULL, active && !(x.y - a.wlo.y > a.wreach && a.whi.y - x.y > a.wreach && a.wreach < hpy));
#include "lv_clip_init.inc"
        // ---- voronoicut!(grid, poly)  voronoigrid.jl:53-81
        while (__any_sync(FULL, alive)) {
#include "lv_clip_round.inc"
        }
        __syncwarp();
        if (a.force_anomaly && active && (slot % 1009) == 7) bad = true;
        if (bad) atomicOr(&a.flags[LVF_OVERFLOW], OVF_ANOMALY);

        // ---- row placement: one atomicAdd per warp
        const int deg = (active && !bad) ? p.m : 0;
        int inc = deg;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int v = __shfl_up_sync(FULL, inc, o2);
            if (lane >= o2) inc += v;
        }
        const int tile_total = __shfl_sync(FULL, inc, 31);
        int base = 0;
        if (lane == 0 && tile_total > 0) base = atomicAdd(&a.flags[LVF_NNZ], tile_total);
        base = __shfl_sync(FULL, base, 0);
        const long long off = (long long)base + inc - deg;

        // ---- emit: ring walk from storage position 0 == sort_edges!  (IO.jl:35-48), area, centroid
        if (slot < a.nslot) {
            double area = 0.0, cx = 0.0, cy = 0.0;
            const bool fits = (long long)base + tile_total <= a.cap_nnz;
            if (!fits && deg > 0) atomicOr(&a.flags[LVF_OVERFLOW], OVF_NNZ);
            if (deg > 0) {
                int cur = nib(p.ord, 0);
                const int start = cur;
                double2 u = p.V(cur);
                bool zero_len = false;
                for (int k = 0; k < deg; k++) {
                    const int nx = nib(p.nxt, cur);
                    const double2 w = p.V(nx);
                    zero_len |= (u.x == w.x && u.y == w.y);
                    const double ax = u.x - x.x, ay = u.y - x.y, bx = w.x - x.x, by = w.y - x.y;
                    const double dA = 0.5 * fabs(ax * by - ay * bx); // polygon.jl:114-122, 201-203
                    area += dA;
                    cx += (dA * ((x.x + u.x) + w.x)) / 3.0;        // polygon.jl:210-219
                    cy += (dA * ((x.y + u.y) + w.y)) / 3.0;
                    if (fits) {
                        const int l = p.L(cur);
                        int cc = l;
                        if (l >= 0) cc = lv_col_of(a, l);
                        a.col[off + k] = cc;
                        a.v1[off + k] = u;
                        a.v2[off + k] = w;
                    }
                    cur = nx;
                    u = w;
                }
                // a ring that does not close, or a zero-length edge (ambiguous for the reference's
                // equality-based chaining), is not the generic case
                if (cur != start || zero_len) atomicOr(&a.flags[LVF_OVERFLOW], OVF_ANOMALY);
            }
            a.rowptr[slot] = (int)off;
            a.rdeg[slot] = (unsigned char)deg;
            a.area[slot] = area;
            a.cen[slot] = deg > 0 ? make_double2(cx / area, cy / area) : make_double2(0.0, 0.0);
        }
        __syncwarp();
    }
    if (STATS && lane == 0) {
        const unsigned long long v[8] = {st_a_it, st_a_ln, st_b_it, st_b_ln, st_c_it, st_c_ln, st_rounds, st_alive};
        for (int k = 0; k < 8; k++) atomicAdd(&stats[k], v[k]);
    }
}

static void clip_stats_print(const char *what, const unsigned long long *h, int ntiles) {
    fprintf(stderr, "[clip stats %s] tiles %d  rounds/tile %.2f alive/round %.2f | A events/tile %.1f lanes %.2f | B iters/tile %.1f lanes %.2f | "
                    "C cuts/tile %.1f lanes %.2f\n", what, ntiles, (double)h[6] / ntiles, (double)h[7] / (h[6] ? h[6] : 1), (double)h[0] / ntiles,
            (double)h[1] / (h[0] ? h[0] : 1), (double)h[2] / ntiles, (double)h[3] / (h[2] ? h[2] : 1), (double)h[4] / ntiles,
            (double)h[5] / (h[4] ? h[4] : 1));
}

template <int MAXE, int BLOCK, int MINB>
static int launch_fast(LvContext *c, const ClipArgs &a) {
    const size_t smem = (size_t)MAXE * BLOCK * (sizeof(double2) + sizeof(int)) + (size_t)2 * QCAP * BLOCK * sizeof(int) +
                        sizeof(LvPathNode) * (size_t)c->gp.npath;
    LV_CUDA(c, cudaFuncSetAttribute(k_clip_fast<MAXE, BLOCK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (int)((c->nslot + 31) / 32);
    if (ntiles == 0) return LV_OK;
    int per_sm = 1;
    LV_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_clip_fast<MAXE, BLOCK, MINB>, BLOCK, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->num_sms * per_sm;
    const long long need = (ntiles + (BLOCK / 32) - 1) / (BLOCK / 32);
    if (grid > need) grid = need;
    static const bool stats = [] { const char *m = getenv("LV_CLIP_STATS"); return m && m[0] == '1'; }();
    if (stats) { // diagnostics: one instrumented launch, synchronous, printed
        unsigned long long *d = nullptr, h[8];
        LV_CUDA(c, cudaMalloc((void **)&d, sizeof(h)));
        LV_CUDA(c, cudaMemsetAsync(d, 0, sizeof(h), c->stream));
        LV_CUDA(c, cudaFuncSetAttribute(k_clip_fast<MAXE, BLOCK, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_clip_fast<MAXE, BLOCK, MINB, true><<<(int)grid, BLOCK, smem, c->stream>>>(a, ntiles, d);
        LV_CUDA(c, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(d);
        clip_stats_print("tile", h, ntiles);
        c->launches++;
        return LV_OK;
    }
    k_clip_fast<MAXE, BLOCK, MINB><<<(int)grid, BLOCK, smem, c->stream>>>(a, ntiles);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}


// ---- lane-refill variant ------------------------------------------------------------------------------------------------
// The tile kernel above keeps a warp on its 32 polygons until the slowest one is done: with 8-14 cuts per polygon about a
// third of the lanes idle in every round.  Here a lane whose polygon is finished PARKS it -- the raw ring: start vertices,
// labels, links -- in global memory and takes the next unassigned slot of the warp's chunk, so the rounds run with nearly
// all lanes busy.  The CSR rows, areas and centroids are produced afterwards by k_clip_emit, one thread per slot: a fully
// convergent walk of the parked rings with the same warp-contiguous row placement (and the same checks) as the tile kernel.
// Rounds, cuts and emission arithmetic are the same code (lv_clip_round.inc), hence the same bytes.
#define RF_CHUNK 256 // slots per ticket
#define RF_THR 6     // lanes without a live polygon that trigger a park + refill pass

template <int MAXE, int BLOCK, int MINB, bool STATS = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_clip_refill(ClipArgs a, int nchunks, unsigned long long *stats = nullptr) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *sv = (double2 *)smem;
    int *sl = (int *)(sv + MAXE * BLOCK);
    int *sq = sl + MAXE * BLOCK;
    int *sqt = sq + QCAP * BLOCK;
    LvPathNode *spath = (LvPathNode *)(sqt + QCAP * BLOCK);
    const LvGridParams g = a.g;
    for (int k = threadIdx.x; k < g.npath; k += BLOCK) spath[k] = a.path[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const double hpx = 0.5 * g.xperiod, hpy = 0.5 * g.yperiod;
    const size_t nsl = (size_t)a.nslot;

    Ring<MAXE, BLOCK> p;
    p.sv = sv; p.sl = sl; p.m = 0; p.used = 0; p.ord = p.pos = p.nxt = p.prv = 0;
    bool running = false, alive = false, active = false, bad = false, scan_done = true, nowork = false;
    double2 x = make_double2(0.0, 0.0);
    double prr = 0.0, slack = 0.0;
    int slot = -1, k1 = 0, k2 = 0, t_scan = -1, s = 0, s_end = 0, t_end = g.npath, t_chk = -1, qh = 0, qn = 0;
    int chunk_next = 0, chunk_end = 0; // the warp's current chunk (same values in every lane)
    const bool wrapx = g.xper != 0, wrapy = g.yper != 0;
    unsigned long long st_a_it = 0, st_a_ln = 0, st_b_it = 0, st_b_ln = 0, st_c_it = 0, st_c_ln = 0, st_rounds = 0, st_alive = 0;

    for (;;) {
        // ---- park finished polygons
        if (running && !alive) {
            if (a.force_anomaly && (slot % 1009) == 7) bad = true;
            if (bad) atomicOr(&a.flags[LVF_OVERFLOW], OVF_ANOMALY);
            for (unsigned u = p.used; u; u &= u - 1) {
                const int sk = __ffs(u) - 1;
                a.park_v[(size_t)sk * nsl + slot] = p.V(sk);
                a.park_l[(size_t)sk * nsl + slot] = p.L(sk);
            }
            a.park_nxt[slot] = p.nxt;
            a.park_hdr[slot] = (bad ? 0 : p.m) | (nib(p.ord, 0) << 8); // degree, first edge of the sorted chain (IO.jl:35-48)
            running = false;
        }
        __syncwarp();
        // ---- hand idle lanes the next slots of the chunk
        unsigned need = __ballot_sync(FULL, !running);
        while (need && !nowork) {
            const int avail = chunk_end - chunk_next;
            if (avail <= 0) {
                int t = 0;
                if (lane == 0) t = atomicAdd(&a.flags[LVF_TICKET], 1);
                t = __shfl_sync(FULL, t, 0);
                if (t >= nchunks) { nowork = true; break; }
                chunk_next = t * RF_CHUNK;
                chunk_end = chunk_next + RF_CHUNK < a.nslot ? chunk_next + RF_CHUNK : a.nslot;
                continue;
            }
            const int rank = __popc(need & lt);
            if (!running && rank < avail) {
                slot = chunk_next + rank;
                active = a.own[slot] != 0; // image / ghost slots have no polygon: their header stays 0 (memset)
                if (active) {
                    x = a.ent_xy[slot];
                    alive = true; running = true; bad = false;
                    prr = 0.0; k1 = k2 = 0;
                    t_scan = -1; s = 0; s_end = 0; t_end = g.npath; t_chk = -1;
                    scan_done = false; qh = 0; qn = 0;
                    slack = 2e-15 * (fabs(x.x) + fabs(x.y));
                    p.m = 0; p.used = 0; p.ord = p.pos = p.nxt = p.prv = 0;
#include "lv_clip_init.inc"
                }
            }
            const int asked = __popc(need);
            chunk_next += asked < avail ? asked : avail;
            need = __ballot_sync(FULL, !running);
        }
        if (!__any_sync(FULL, running)) break;
        // ---- rounds (voronoicut!(grid, poly)  voronoigrid.jl:53-81) until enough lanes are free again
        for (;;) {
#include "lv_clip_round.inc"
            const unsigned al = __ballot_sync(FULL, alive);
            if (!al) break;
            if (!nowork && __popc(~al) >= RF_THR) break;
        }
    }
    if (STATS && lane == 0) {
        const unsigned long long v[8] = {st_a_it, st_a_ln, st_b_it, st_b_ln, st_c_it, st_c_ln, st_rounds, st_alive};
        for (int k = 0; k < 8; k++) atomicAdd(&stats[k], v[k]);
    }
}
#undef LV_STAT
#define LV_STAT(...)

// One thread per slot: CSR row + area + centroid from the parked ring.  Same walk, arithmetic and checks as the emission
// part of k_clip_fast; rows of 32 consecutive slots are contiguous (one atomicAdd per warp).
template <int MAXE>
__global__ void __launch_bounds__(256) k_clip_emit(ClipArgs a) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const bool valid = slot < a.nslot;
    const size_t nsl = (size_t)a.nslot;
    const int hdr = valid ? a.park_hdr[slot] : 0;
    const int deg = hdr & 0xff;
    int inc = deg;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int v = __shfl_up_sync(FULL, inc, o2);
        if (lane >= o2) inc += v;
    }
    const int tile_total = __shfl_sync(FULL, inc, 31);
    int base = 0;
    if (lane == 0 && tile_total > 0) base = atomicAdd(&a.flags[LVF_NNZ], tile_total);
    base = __shfl_sync(FULL, base, 0);
    const long long off = (long long)base + inc - deg;
    if (!valid) return;
    double area = 0.0, cx = 0.0, cy = 0.0;
    const bool fits = (long long)base + tile_total <= a.cap_nnz;
    if (!fits && deg > 0) atomicOr(&a.flags[LVF_OVERFLOW], OVF_NNZ);
    if (deg > 0) {
        const double2 x = a.ent_xy[slot];
        const u64 nxt = a.park_nxt[slot];
        int cur = (hdr >> 8) & 15;
        const int start = cur;
        double2 u = a.park_v[(size_t)cur * nsl + slot];
        bool zero_len = false;
        for (int k = 0; k < deg; k++) {
            const int nx = nib(nxt, cur);
            const double2 w = a.park_v[(size_t)nx * nsl + slot];
            zero_len |= (u.x == w.x && u.y == w.y);
            const double ax = u.x - x.x, ay = u.y - x.y, bx = w.x - x.x, by = w.y - x.y;
            const double dA = 0.5 * fabs(ax * by - ay * bx); // polygon.jl:114-122, 201-203
            area += dA;
            cx += (dA * ((x.x + u.x) + w.x)) / 3.0;        // polygon.jl:210-219
            cy += (dA * ((x.y + u.y) + w.y)) / 3.0;
            if (fits) {
                const int l = a.park_l[(size_t)cur * nsl + slot];
                int cc = l;
                if (l >= 0) cc = lv_col_of(a, l);
                a.col[off + k] = cc;
                a.v1[off + k] = u;
                a.v2[off + k] = w;
            }
            cur = nx;
            u = w;
        }
        if (cur != start || zero_len) atomicOr(&a.flags[LVF_OVERFLOW], OVF_ANOMALY);
    }
    a.rowptr[slot] = (int)off;
    a.rdeg[slot] = (unsigned char)deg;
    a.area[slot] = area;
    a.cen[slot] = deg > 0 ? make_double2(cx / area, cy / area) : make_double2(0.0, 0.0);
}

template <int MAXE, int BLOCK, int MINB>
static int launch_refill(LvContext *c, ClipArgs a) {
    const int64_t nslot = c->nslot;
    if (nslot == 0) return LV_OK;
    // parking space: MAXE start vertices + labels per slot, links, header
    LV_TRY(lv_ensure(c, (void **)&c->d_park_v, &c->cap_park_v, (int64_t)MAXE * nslot, sizeof(double2)));
    LV_TRY(lv_ensure(c, (void **)&c->d_park_l, &c->cap_park_l, (int64_t)MAXE * nslot, sizeof(int)));
    LV_TRY(lv_ensure(c, (void **)&c->d_park_nxt, &c->cap_park_n, nslot, sizeof(unsigned long long)));
    LV_TRY(lv_ensure(c, (void **)&c->d_park_hdr, &c->cap_park_h, nslot, sizeof(int)));
    a.park_v = (double2 *)c->d_park_v; a.park_l = (int *)c->d_park_l; a.park_nxt = (unsigned long long *)c->d_park_nxt; a.park_hdr = (int *)c->d_park_hdr;
    LV_CUDA(c, cudaMemsetAsync(c->d_park_hdr, 0, sizeof(int) * (size_t)nslot, c->stream));
    const size_t smem = (size_t)MAXE * BLOCK * (sizeof(double2) + sizeof(int)) + (size_t)2 * QCAP * BLOCK * sizeof(int) +
                        sizeof(LvPathNode) * (size_t)c->gp.npath;
    LV_CUDA(c, cudaFuncSetAttribute(k_clip_refill<MAXE, BLOCK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nchunks = (int)((nslot + RF_CHUNK - 1) / RF_CHUNK);
    int per_sm = 1;
    LV_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_clip_refill<MAXE, BLOCK, MINB>, BLOCK, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->num_sms * per_sm;
    const long long need = ((long long)nchunks * (RF_CHUNK / 32) + (BLOCK / 32) - 1) / (BLOCK / 32);
    if (grid > need) grid = need;
    static const bool stats = [] { const char *m = getenv("LV_CLIP_STATS"); return m && m[0] == '1'; }();
    if (stats) {
        unsigned long long *d = nullptr, h[8];
        LV_CUDA(c, cudaMalloc((void **)&d, sizeof(h)));
        LV_CUDA(c, cudaMemsetAsync(d, 0, sizeof(h), c->stream));
        LV_CUDA(c, cudaFuncSetAttribute(k_clip_refill<MAXE, BLOCK, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_clip_refill<MAXE, BLOCK, MINB, true><<<(int)grid, BLOCK, smem, c->stream>>>(a, nchunks, d);
        LV_CUDA(c, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(d);
        clip_stats_print("refill", h, (int)((nslot + 31) / 32));
    } else
    k_clip_refill<MAXE, BLOCK, MINB><<<(int)grid, BLOCK, smem, c->stream>>>(a, nchunks);
    k_clip_emit<MAXE><<<(int)((nslot + 255) / 256), 256, 0, c->stream>>>(a);
    c->launches += 2;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// The tile kernel (one warp stays on its 32 polygons) is the production kernel.  LV_CLIP_MODE=refill selects the lane-refill
// variant: measured SLOWER on B200 (16.8M cells: 25.8 + 2.8 ms against 21.7 ms; 11.7 instead of 13.3 active lanes per
// instruction, profiles/r2_clip_refill_16M_full.txt) -- refilled lanes are at different "ages" (young polygons scan, old ones
// cut), so every round runs all three phases at partial occupancy, whereas a tile's 32 polygons age in lockstep.
int lv_clip_launch_fast(LvContext *c, const ClipArgs &a, int level) {
    static const bool tile = [] { const char *m = getenv("LV_CLIP_MODE"); return !(m && !strcmp(m, "refill")); }();
    if (tile) {
        if (level == 0) return launch_fast<12, 128, 6>(c, a);
        return launch_fast<16, 128, 4>(c, a);
    }
    if (level == 0) return launch_refill<12, 128, 6>(c, a);
    return launch_refill<16, 128, 4>(c, a);
}
