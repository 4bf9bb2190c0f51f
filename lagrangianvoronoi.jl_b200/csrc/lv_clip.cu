// lv_clip.cu -- K2: per-generator half-plane clipping -> CSR mesh.
//
// Replaces voronoicut!(grid, poly) + sort_edges! of the reference (voronoigrid.jl:53-81,
// 102-105; polygon.jl:37-107; IO.jl:35-48) and emits, in slot order, the neighbour
// connectivity in CSR form with both end points of every edge, cell areas
// (polygon.jl:114-122) and centroids (polygon.jl:210-219).
//
// One thread owns one polygon.  The polygon's edge list lives in shared memory in a
// [edge][thread] layout (bank-conflict free for any per-thread edge index), cut in exactly the
// reference's sequence: same candidate order (magic_path walk x ascending labels in a
// bucket), same swap-remove edge-list mutation (fastvector.jl:44-50), same floating-point
// expressions without FMA contraction (the file is compiled with -fmad=false), so vertices,
// edge order and therefore connectivity are bit-identical to `julia -t 1`.
//
// CSR offsets are produced in the same launch by a decoupled look-back scan over tiles
// (dynamic tile ids from a ticket counter guarantee forward progress), so edges are written
// straight to their final position: no second pass over the mesh.
#include "lv_clip.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstdlib>
#include <cstring>

__device__ __forceinline__ int signum(double x) { // polygon.jl:9-16
    if (x < -SIGNUM_EPS) return -1;
    else if (x > SIGNUM_EPS) return 1;
    return 0;
}

// shared-memory polygon: edge k of thread t at [k * BLOCK + t]
template <int MAXE, int BLOCK>
struct Poly {
    double2 *v1, *v2;
    int *lab;
    int m;
    bool ovf;
    __device__ __forceinline__ double2 &V1(int k) { return v1[k * BLOCK + threadIdx.x]; }
    __device__ __forceinline__ double2 &V2(int k) { return v2[k * BLOCK + threadIdx.x]; }
    __device__ __forceinline__ int &L(int k) { return lab[k * BLOCK + threadIdx.x]; }
    __device__ __forceinline__ void push(double2 a, double2 b, int l) { // fastvector.jl:14-22
        if (m >= MAXE) { ovf = true; return; }
        V1(m) = a; V2(m) = b; L(m) = l;
        m++;
    }
    __device__ __forceinline__ void deleteat(int i) { // fastvector.jl:44-50 (0-based i)
        if (m - 1 > i) { V1(i) = V1(m - 1); V2(i) = V2(m - 1); L(i) = L(m - 1); }
        m--;
    }
};

// polygon.jl:51-97
template <int MAXE, int BLOCK>
__device__ __forceinline__ bool voronoicut(Poly<MAXE, BLOCK> &p, double2 x, double2 y, int label) {
    const double dx = y.x - x.x, dy = y.y - x.y;
    const double mx = 0.5 * (y.x + x.x), my = 0.5 * (y.y + x.y);
    const double c = dx * mx + dy * my;
    int i = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double2 X = make_double2(nan, nan), Y = make_double2(nan, nan);
    while (i < p.m) {
        const double2 a = p.V1(i), b = p.V2(i);
        const double f1 = (dx * a.x + dy * a.y) - c;
        const double f2 = (dx * b.x + dy * b.y) - c;
        const int s1 = signum(f1), s2 = signum(f2);
        const int s12 = s1 + s2;
        if ((0 <= s12 && s12 <= 1) && ((s1 | s2) != 0)) {
            Y = X;
            const double r = 1.0 / (f1 - f2);
            X = make_double2(r * (f1 * b.x - f2 * a.x), r * (f1 * b.y - f2 * a.y));
            if (s1 == 0) X = a;
            if (s2 == 0) X = b;
            if (s1 == 1) p.V1(i) = X; else p.V2(i) = X;
        }
        if (1 <= s12) p.deleteat(i);
        else i++;
    }
    const bool ynull = isnan(Y.x) && isnan(Y.y);
    if (!ynull && !(X.x == Y.x && X.y == Y.y)) {
        // reorient so that the generator lies to the right of v1->v2 (clockwise)
        const double cr = (Y.x - X.x) * (x.y - X.y) - (Y.y - X.y) * (x.x - X.x);
        if (cr > 0.0) p.push(Y, X, label);
        else p.push(X, Y, label);
        return true;
    }
    return false;
}

template <int MAXE, int BLOCK>
__device__ __forceinline__ double influence_rr(Poly<MAXE, BLOCK> &p, double2 x) { // polygon.jl:101-107
    double rr = 0.0;
    for (int k = 0; k < p.m; k++) {
        const double2 a = p.V1(k);
        const double ex = a.x - x.x, ey = a.y - x.y;
        const double t = 4.0 * (ex * ex + ey * ey);
        if (isnan(t) || isnan(rr)) rr = t + rr; // Julia max propagates NaN
        else rr = rr > t ? rr : t;
    }
    return rr;
}

template <int MAXE, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_clip(ClipArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *sv1 = (double2 *)smem;
    double2 *sv2 = sv1 + MAXE * BLOCK;
    int *slab = (int *)(sv2 + MAXE * BLOCK);
    LvPathNode *spath = (LvPathNode *)(slab + MAXE * BLOCK);
    __shared__ int s_scan[34];
    __shared__ int s_tile;
    __shared__ long long s_prefix;

    const LvGridParams g = a.g;
    for (int k = threadIdx.x; k < g.npath; k += BLOCK) spath[k] = a.path[k];
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.flags[LVF_TICKET], 1);
    __syncthreads();
    const int tile = s_tile;
    const int slot = tile * BLOCK + threadIdx.x;

    Poly<MAXE, BLOCK> p;
    p.v1 = sv1; p.v2 = sv2; p.lab = slab; p.m = 0; p.ovf = false;
    bool active = false;
    double2 x = make_double2(0.0, 0.0);
    if (slot < a.nslot) {
        active = a.own[slot] != 0;
        x = a.ent_xy[slot];
    }
    double area = 0.0;
    double2 cen = make_double2(0.0, 0.0);
    if (active) {
        // reset!  polygon.jl:37-47
        const double2 A = make_double2(g.cminx, g.cminy), C = make_double2(g.cmaxx, g.cmaxy);
        const double2 B = make_double2(C.x, A.y), D = make_double2(A.x, C.y);
        p.push(B, A, BD_DOWN);
        p.push(A, D, BD_LEFT);
        p.push(D, C, BD_UP);
        p.push(C, B, BD_RIGHT);
        // voronoicut!(grid, poly)  voronoigrid.jl:53-81
        double prr = influence_rr(p, x);
        int k1, k2;
        if (!lv_findkey(g, x, k1, k2)) { atomicOr(&a.flags[LVF_NAN], 1); k1 = k2 = -(1 << 30); }
        for (int t = 0; t < g.npath; t++) {
            const LvPathNode nd = spath[t];
            if (nd.rr > prr) break;
            if (nd.rr > g.rr_max) { atomicOr(&a.flags[LVF_DESTROYED], 1); break; }
            const int c1 = k1 + nd.i1, c2 = k2 + nd.i2;
            if (!(c1 >= 1 && c1 <= g.n1 && c2 >= 1 && c2 <= g.n2)) continue;
            const int lin = (c1 - 1) + g.n1 * (c2 - 1);
            const int s0 = a.cell_start[lin], s1 = a.cell_start[lin + 1];
            for (int s = s0; s < s1; s++) {
                const double2 q = a.ent_xy[s];
                const double2 y = lv_neighbor_pos(g, x, q);
                const double ex = x.x - y.x, ey = x.y - y.y;
                if ((x.x == y.x && x.y == y.y) || (ex * ex + ey * ey > prr)) continue;
                if (voronoicut(p, x, y, s)) prr = influence_rr(p, x);
            }
            if (p.ovf) break;
        }
        if (p.ovf) { atomicOr(&a.flags[LVF_OVERFLOW], OVF_POLY); p.m = 0; }
        // sort_edges!  IO.jl:35-48
        for (int i = 0; i < p.m; i++) {
            const double2 last = p.V2(i);
            for (int j = i + 1; j < p.m; j++) {
                const double2 w = p.V1(j);
                if (last.x == w.x && last.y == w.y) {
                    const double2 t1 = p.V1(i + 1), t2 = p.V2(i + 1);
                    const int tl = p.L(i + 1);
                    p.V1(i + 1) = p.V1(j); p.V2(i + 1) = p.V2(j); p.L(i + 1) = p.L(j);
                    p.V1(j) = t1; p.V2(j) = t2; p.L(j) = tl;
                    break;
                }
            }
        }
        // area (polygon.jl:114-122) and centroid (polygon.jl:210-219)
        double cx = 0.0, cy = 0.0;
        for (int k = 0; k < p.m; k++) {
            const double2 u = p.V1(k), w = p.V2(k);
            const double ax = u.x - x.x, ay = u.y - x.y, bx = w.x - x.x, by = w.y - x.y;
            const double dA = 0.5 * fabs(ax * by - ay * bx);
            area += dA;
            cx += (dA * ((x.x + u.x) + w.x)) / 3.0;
            cy += (dA * ((x.y + u.y) + w.y)) / 3.0;
        }
        cen = make_double2(cx / area, cy / area);
    }

    // ---- CSR offsets: block scan + decoupled look-back over tiles
    const int deg = active ? p.m : 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = deg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_scan[w] = inc;
    __syncthreads();
    if (w == 0) {
        constexpr int NW = BLOCK / 32;
        int s = lane < NW ? s_scan[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        s_scan[lane] = si - s;
        if (lane == 31) s_scan[32] = si;
    }
    __syncthreads();
    const int block_total = s_scan[32];
    const int local_off = s_scan[w] + inc - deg;
    if (threadIdx.x == 0) {
        long long prefix = 0;
        volatile unsigned long long *st = a.tile_state;
        if (tile > 0) {
            st[tile] = TS_AGG | (unsigned long long)block_total;
            __threadfence();
            int j = tile - 1;
            for (;;) {
                unsigned long long v = st[j];
                if ((v & ~TS_MASK) == TS_INC) { prefix += (long long)(v & TS_MASK); break; }
                if ((v & ~TS_MASK) == TS_AGG) { prefix += (long long)(v & TS_MASK); j--; }
            }
        }
        st[tile] = TS_INC | (unsigned long long)(prefix + block_total);
        __threadfence();
        s_prefix = prefix;
        if ((long long)(tile + 1) * BLOCK >= a.nslot) { // last tile: totals
            a.rowptr[a.nslot] = (int)(prefix + block_total);
            a.flags[LVF_NNZ] = (int)(prefix + block_total);
        }
    }
    __syncthreads();
    const long long off = s_prefix + local_off;
    if (slot < a.nslot) {
        a.rowptr[slot] = (int)off;
        a.rdeg[slot] = (unsigned char)deg;
        a.area[slot] = area;
        a.cen[slot] = cen;
        if (off + deg > a.cap_nnz) {
            if (deg > 0) atomicOr(&a.flags[LVF_OVERFLOW], OVF_NNZ);
        } else {
            for (int k = 0; k < deg; k++) {
                const int l = p.L(k);
                int cc = l;
                if (l >= 0) cc = lv_col_of(a, l);
                a.col[off + k] = cc;
                a.v1[off + k] = p.V1(k);
                a.v2[off + k] = p.V2(k);
            }
        }
    }
}

template <int MAXE, int BLOCK>
static int launch_clip(LvContext *c, const ClipArgs &a) {
    const size_t smem = (size_t)MAXE * BLOCK * (sizeof(double2) * 2 + sizeof(int)) + sizeof(LvPathNode) * (size_t)c->gp.npath;
    LV_CUDA(c, cudaFuncSetAttribute(k_clip<MAXE, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (int)((c->nslot + BLOCK - 1) / BLOCK);
    if (ntiles == 0) return LV_OK;
    k_clip<MAXE, BLOCK><<<ntiles, BLOCK, smem, c->stream>>>(a);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// Capacity / exactness ladder.  Levels 0-1: linked-slot kernel (12 / 16 slots per polygon);
// levels 2-4: edge-list kernel (16 / 32 / 128 edges).  A polygon that outgrows a level moves the
// (sticky) capacity level up; an anomaly reported by the linked-slot kernel reruns this remesh
// with the edge-list kernel, which replays the reference literally.
// LV_CLIP_MODE=plain forces the edge-list kernel (used by the tests to cross-check both).
// One attempt: buffers for need_nnz edges, flag reset, launch at `level` (nothing is synchronised).
static int clip_attempt(LvContext *c, int level, int64_t need_nnz) {
    const int64_t nslot = c->nslot;
    if (need_nnz > c->cap_nnz) {
        int64_t c1 = c->cap_nnz, c2 = c->cap_nnz, c3 = c->cap_nnz;
        LV_TRY(lv_ensure(c, (void **)&c->d_col, &c1, need_nnz, sizeof(int)));
        LV_TRY(lv_ensure(c, (void **)&c->d_v1, &c2, need_nnz, sizeof(double2)));
        LV_TRY(lv_ensure(c, (void **)&c->d_v2, &c3, need_nnz, sizeof(double2)));
        c->cap_nnz = need_nnz;
    }
    const int block = level <= 1 ? 32 : (level == 2 ? 128 : (level == 3 ? 64 : 32)); // tile size
    const int64_t ntiles = (nslot + block - 1) / block;
    LV_TRY(lv_ensure(c, (void **)&c->d_tile_state, &c->cap_tiles, ntiles + 1, sizeof(unsigned long long)));
    LV_CUDA(c, cudaMemsetAsync(c->d_tile_state, 0, sizeof(unsigned long long) * (size_t)(ntiles + 1), c->stream));
    LV_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 8, c->stream));
    ClipArgs a;
    a.g = c->gp;
    a.path = c->d_path;
    a.cell_start = c->d_cell_start;
    a.ent_label = c->d_ent_label;
    a.ent_xy = c->d_ent_xy;
    a.prim_of_label = c->d_prim_of_label;
    a.own = c->d_own;
    a.nslot = (int)nslot;
    a.rowptr = c->d_rowptr;
    a.rdeg = c->d_deg;
    a.col = c->d_col;
    a.v1 = c->d_v1;
    a.v2 = c->d_v2;
    a.area = c->d_area;
    a.cen = c->d_cen;
    a.tile_state = c->d_tile_state;
    a.flags = c->d_flags;
    a.cap_nnz = c->cap_nnz;
    a.park_v = nullptr; a.park_l = nullptr; a.park_nxt = nullptr; a.park_hdr = nullptr;
    { const char *fa = getenv("LV_CLIP_FORCE_ANOMALY"); a.force_anomaly = fa && fa[0] == '1'; }
    {
        int reach = 0;
        for (int k = 0; k < c->gp.npath; k++) reach = std::max(reach, std::max(std::abs(c->h_path[k].i1), std::abs(c->h_path[k].i2)));
        a.wreach = (double)(reach + 2) * c->gp.h; // a candidate's bucket is at most `reach` buckets away: |q - x| < (reach + 1) h
        a.wlo = make_double2(c->bmin[0], c->bmin[1]);
        a.whi = make_double2(c->bmax[0], c->bmax[1]);
    }
    {
        LvProfScope prof(c, LV_PROF_CLIP);
        if (level <= 1) LV_TRY(lv_clip_launch_fast(c, a, level));
        else if (level == 2) LV_TRY((launch_clip<16, 128>(c, a)));
        else if (level == 3) LV_TRY((launch_clip<32, 64>(c, a)));
        else LV_TRY((launch_clip<128, 32>(c, a)));
    }
    c->clip_last_level = level;
    return LV_OK;
}

// Reads the status words of an attempt (hf: host copy).  done: the mesh stands (c->nnz set); otherwise level / need_nnz
// say how to try again.
static int clip_decide(LvContext *c, const int *hf, int &level, int64_t &need_nnz, bool &done) {
    done = false;
    if (c->nslot == 0) { c->nnz = 0; done = true; return LV_OK; }
    if (hf[LVF_NAN]) return lv_set_error(c, LV_ENAN, "generator position is NaN or Inf");
    if (hf[LVF_DESTROYED]) return lv_set_error(c, LV_EDESTROYED, "The Voronoi Mesh has been destroyed.");
    const int ov = hf[LVF_OVERFLOW];
    if (ov == 0) { c->nnz = hf[LVF_NNZ]; done = true; return LV_OK; }
    if (ov & OVF_NNZ) need_nnz = (int64_t)hf[LVF_NNZ] + 1024;
    if (ov & OVF_POLY) { // capacity: sticky
        if (level >= 4) return lv_set_error(c, LV_ECAPACITY, "polygon with more than 128 edges during clipping");
        level = level == 0 ? 1 : (level < 3 ? 3 : 4);
        c->clip_level = level;
    } else if (ov & OVF_ANOMALY) { // exactness: this remesh only
        c->clip_anomalies++;
        level = level == 0 ? 2 : 3;
    }
    return LV_OK;
}

static int clip_first_level(const LvContext *c) {
    const char *mode = getenv("LV_CLIP_MODE");
    const bool force_plain = mode && !strcmp(mode, "plain");
    int level = c->clip_level;
    if (force_plain && level < 2) level = 2;
    return level;
}

static int clip_loop(LvContext *c, int level, int64_t need_nnz, int attempts) {
    for (int attempt = 0; attempt < attempts; attempt++) {
        LV_TRY(clip_attempt(c, level, need_nnz));
        LV_TRY(lv_publish_flags(c, nullptr));
        bool done = false;
        LV_TRY(clip_decide(c, c->h_flags, level, need_nnz, done));
        if (done) return LV_OK;
    }
    return lv_set_error(c, LV_ECAPACITY, "clip kernel did not fit after retries");
}

int lv_clip_run(LvContext *c) {
    // edge buffers: 6n on a torus (Euler), fewer with walls plus the wall edges; grow on demand
    return clip_loop(c, clip_first_level(c), 7 * c->nslot + 1024, 8);
}

// Pipelined host-buffer mode (lv_pipeline.cu): the first attempt is only queued -- its status words are stored into a
// snapshot in mapped pinned memory by a kernel behind it -- and lv_clip_resume looks at them once the host gets there.
int lv_clip_attempt_first(LvContext *c) {
    return clip_attempt(c, clip_first_level(c), 7 * c->nslot + 1024);
}
int lv_clip_resume(LvContext *c, const int *snap, bool *replayed) {
    int level = c->clip_last_level;
    int64_t need_nnz = c->cap_nnz;
    bool done = false;
    *replayed = false;
    LV_TRY(clip_decide(c, snap, level, need_nnz, done));
    if (done) return LV_OK;
    *replayed = true;
    return clip_loop(c, level, need_nnz, 7);
}
