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
#define EV 8   // scan events per lane and round

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

template <int MAXE, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_clip_fast(ClipArgs a, int ntiles) {
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
        if (active) {
            // reset!  polygon.jl:37-47: B->A DOWN, A->D LEFT, D->C UP, C->B RIGHT (storage order 0..3)
            p.V(0) = make_double2(g.cmaxx, g.cminy); p.L(0) = BD_DOWN;
            p.V(1) = make_double2(g.cminx, g.cminy); p.L(1) = BD_LEFT;
            p.V(2) = make_double2(g.cminx, g.cmaxy); p.L(2) = BD_UP;
            p.V(3) = make_double2(g.cmaxx, g.cmaxy); p.L(3) = BD_RIGHT;
            p.ord = 0x3210ull; p.pos = 0x3210ull;
            p.nxt = 0x0321ull; // 0->1->2->3->0
            p.prv = 0x2103ull;
            p.used = 0xfu;
            p.m = 4;
            prr = ring_influence_rr(p, x);
            if (!lv_findkey(g, x, k1, k2)) { atomicOr(&a.flags[LVF_NAN], 1); k1 = k2 = -(1 << 30); }
        }

        // ---- voronoicut!(grid, poly)  voronoigrid.jl:53-81
        while (__any_sync(FULL, alive)) {
            // phase A: scan ahead, queue candidates that pass the distance filter with the current radius.  One event = (advance to
            // the next path node if the current bucket is exhausted) + (test one candidate): lanes that just entered a bucket test
            // its first entry in the same event, so the candidate part runs with more lanes and a cell needs fewer events.
            // The filter here is CONSERVATIVE (phase B re-applies the reference's exact test before anything is decided): the
            // candidate is dropped only if it is outside the influence radius by more than the rounding error that separates
            // |q - x|^2 from the reference's |x - (x + arrow)|^2 -- at most ulp(x + v) per component, bounded by
            // 1e-15 (|x.x| + |x.y|) + 1e-15 |v| (see lv_neighbor_pos: the periodic shift itself is the same operation).
#pragma unroll 1
            for (int ev = 0; ev < EV; ev++) {
                const bool scanning = alive && !scan_done && (qn - qh) < QCAP;
                if (!__any_sync(FULL, scanning)) break;
                if (scanning) {
                    if (s >= s_end) {
                        if (++t_scan >= g.npath) {
                            scan_done = true; t_end = g.npath;
                        } else {
                            const LvPathNode nd = spath[t_scan];
                            if (nd.rr > prr || nd.rr > g.rr_max) { scan_done = true; t_end = t_scan; } // the walk ends here at the latest
                            else {
                                const int c1 = k1 + nd.i1, c2 = k2 + nd.i2;
                                s = s_end = 0;
                                if (c1 >= 1 && c1 <= g.n1 && c2 >= 1 && c2 <= g.n2) {
                                    const int lin = (c1 - 1) + g.n1 * (c2 - 1);
                                    s = a.cell_start[lin];
                                    s_end = a.cell_start[lin + 1];
                                }
                            }
                        }
                    }
                    if (!scan_done && s < s_end) {
                        const double2 q = a.ent_xy[s];
                        double vx = q.x - x.x, vy = q.y - x.y;
                        if (g.xper) { if (vx > hpx) vx -= g.xperiod; else if (vx < -hpx) vx += g.xperiod; }
                        if (g.yper) { if (vy > hpy) vy -= g.yperiod; else if (vy < -hpy) vy += g.yperiod; }
                        const double d2 = vx * vx + vy * vy;
                        if (!(d2 * (1.0 - 2e-15) - (fabs(vx) + fabs(vy)) * slack > prr)) { // NaN ends up queued: phase B decides
                            sq[(qn & (QCAP - 1)) * BLOCK + threadIdx.x] = s;
                            sqt[(qn & (QCAP - 1)) * BLOCK + threadIdx.x] = t_scan;
                            qn++;
                        }
                        s++;
                    }
                }
            }
            __syncwarp();
            // phase B: pop candidates in order until one has a vertex outside its half plane
            unsigned plus = 0, zero = 0;
            double dx = 0, dy = 0, c = 0;
            int cand = 0;
            while (alive && qh < qn && !plus) {
                cand = sq[(qh & (QCAP - 1)) * BLOCK + threadIdx.x];
                const int tq = sqt[(qh & (QCAP - 1)) * BLOCK + threadIdx.x];
                qh++;
                if (tq > t_chk) { // node entry test of the reference, with the radius of this moment (voronoigrid.jl:60-62)
                    t_chk = tq;
                    if (spath[tq].rr > prr) { alive = false; break; }
                }
                const double2 q = a.ent_xy[cand];
                const double2 y = lv_neighbor_pos(g, x, q);
                const double ex = x.x - y.x, ey = x.y - y.y;
                if ((x.x == y.x && x.y == y.y) || (ex * ex + ey * ey > prr)) continue; // voronoigrid.jl:73-75
                dx = y.x - x.x; dy = y.y - x.y;                                        // polygon.jl:52-54
                const double mx = 0.5 * (y.x + x.x), my = 0.5 * (y.y + x.y);
                c = dx * mx + dy * my;
                zero = 0;
                for (unsigned u = p.used; u; u &= u - 1) {
                    const int sk = __ffs(u) - 1;
                    const double2 v = p.V(sk);
                    const double f = (dx * v.x + dy * v.y) - c;
                    if (f > SIGNUM_EPS) plus |= 1u << sk;
                    else if (!(f < -SIGNUM_EPS)) zero |= 1u << sk;
                }
            }
            if (alive && !plus && scan_done && qh >= qn) {
                // the scan stopped at node t_end; replay that node's entry test with the final radius
                if (t_end < g.npath) {
                    const LvPathNode nd = spath[t_end];
                    if (!(nd.rr > prr) && nd.rr > g.rr_max) atomicOr(&a.flags[LVF_DESTROYED], 1); // voronoigrid.jl:63-65
                }
                alive = false;
            }
            __syncwarp();
            // phase C: cut (polygon.jl:59-96)
            if (plus) {
                const int np = __popc(plus);
                unsigned starts = 0; // outside vertices whose predecessor is not outside
                for (unsigned u = plus; u; u &= u - 1) {
                    const int sk = __ffs(u) - 1;
                    if (!((plus >> nib(p.prv, sk)) & 1)) starts |= 1u << sk;
                }
                if (__popc(starts) != 1) bad = true;
                else {
                    const int pfirst = __ffs(starts) - 1;
                    const int a_slot = nib(p.prv, pfirst); // edge (-|0, +): the polygon is left here
                    int b_slot = pfirst, cnt = 1;          // edge (+, -|0): the polygon is re-entered here
                    unsigned del = 0;                      // edges with s1 + s2 >= 1 (polygon.jl:81-85)
                    while (cnt < np && ((plus >> nib(p.nxt, b_slot)) & 1)) { del |= 1u << b_slot; b_slot = nib(p.nxt, b_slot); cnt++; }
                    const int nb = nib(p.nxt, b_slot);
                    if (cnt != np || ((plus >> nb) & 1)) bad = true; // outside vertices not contiguous
                    else {
                        const bool za = (zero >> a_slot) & 1, zb = (zero >> nb) & 1;
                        if (za) del |= 1u << a_slot;
                        if (zb) del |= 1u << b_slot;
                        const double2 va1 = p.V(a_slot), va2 = p.V(pfirst), vb1 = p.V(b_slot), vb2 = p.V(nb);
                        double2 XA, XB;
                        {
                            const double f1 = (dx * va1.x + dy * va1.y) - c, f2 = (dx * va2.x + dy * va2.y) - c;
                            const double r = 1.0 / (f1 - f2);
                            XA = make_double2(r * (f1 * va2.x - f2 * va1.x), r * (f1 * va2.y - f2 * va1.y));
                            if (za) XA = va1;
                        }
                        {
                            const double f1 = (dx * vb1.x + dy * vb1.y) - c, f2 = (dx * vb2.x + dy * vb2.y) - c;
                            const double r = 1.0 / (f1 - f2);
                            XB = make_double2(r * (f1 * vb2.x - f2 * vb1.x), r * (f1 * vb2.y - f2 * vb1.y));
                            if (zb) XB = vb2;
                        }
                        // which of the two edges the reference's scan over the storage order meets last
                        int last_kind, mm = p.m;
                        u64 o = p.ord;
                        if (del == 0) last_kind = nib(p.pos, b_slot) > nib(p.pos, a_slot) ? 2 : 1;
                        else { // replay deleteat! (swap-remove) on the storage order
                            int i = 0;
                            last_kind = 0;
                            while (i < mm) {
                                const int sk = nib(o, i);
                                if (sk == a_slot) last_kind = 1;
                                else if (sk == b_slot) last_kind = 2;
                                if ((del >> sk) & 1) { o = setnib(o, i, nib(o, mm - 1)); mm--; }
                                else { p.pos = setnib(p.pos, sk, i); i++; }
                            }
                        }
                        // the reference's X is the point met last, Y the one before (polygon.jl:67-69, 87-95)
                        const double2 X = last_kind == 2 ? XB : XA, Y = last_kind == 2 ? XA : XB;
                        const double cr = (Y.x - X.x) * (x.y - X.y) - (Y.y - X.y) * (x.x - X.x);
                        const bool invert = cr > 0.0;
                        const bool nanx = (isnan(XA.x) && isnan(XA.y)) || (isnan(XB.x) && isnan(XB.y)) || isnan(cr);
                        const bool same = (XA.x == XB.x && XA.y == XB.y);
                        const unsigned freemask = ~(p.used & ~del) & ((1u << MAXE) - 1u);
                        if (nanx || same || (invert != (last_kind == 2))) bad = true; // not the generic case
                        else if (!freemask) { bad = true; atomicOr(&a.flags[LVF_OVERFLOW], OVF_POLY); }
                        else {
                            p.used &= ~del;
                            const int ns = __ffs(freemask) - 1;
                            p.used |= 1u << ns;
                            p.V(ns) = XA; // new edge XA -> XB keeps the generator on its right
                            p.L(ns) = cand;
                            const int pa = za ? nib(p.prv, a_slot) : a_slot; // last surviving edge before the new one
                            p.nxt = setnib(p.nxt, pa, ns);
                            p.prv = setnib(p.prv, ns, pa);
                            const int sb = zb ? nb : b_slot;                 // first surviving edge after the new one
                            if (!zb) p.V(b_slot) = XB;                       // edge b keeps its end, starts at XB
                            p.nxt = setnib(p.nxt, ns, sb);
                            p.prv = setnib(p.prv, sb, ns);
                            p.ord = setnib(o, mm, ns); // push!
                            p.pos = setnib(p.pos, ns, mm);
                            p.m = mm + 1;
                            const double prr_new = ring_influence_rr(p, x); // voronoigrid.jl:76-78
                            if (!(prr_new <= prr)) bad = true;              // scanning ahead relied on a non-increasing radius
                            prr = prr_new;
                        }
                    }
                }
                if (bad) alive = false;
            }
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
    k_clip_fast<MAXE, BLOCK, MINB><<<(int)grid, BLOCK, smem, c->stream>>>(a, ntiles);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int lv_clip_launch_fast(LvContext *c, const ClipArgs &a, int level) {
    if (level == 0) return launch_fast<12, 128, 6>(c, a);
    return launch_fast<16, 128, 4>(c, a);
}
