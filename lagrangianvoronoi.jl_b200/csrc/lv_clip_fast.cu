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
//  * all bookkeeping lives in registers: `nxt` (4-bit successor per slot), `ord` (the
//    reference's FastVector storage order, 4 bits per position, needed because the storage
//    order decides which edge starts the sorted chain and which intersection is "X"), a
//    16-bit occupancy mask and two 16-bit sign masks;
//  * a cut first classifies every vertex (f > eps, |f| <= eps, f < -eps); if no vertex is
//    outside nothing else happens (the reference would scan the whole edge list to find that
//    out).  Otherwise the reference's scan over the storage order, including its swap-remove
//    deletions (fastvector.jl:44-50), is replayed on the 2-bit signs only, which yields the
//    surviving storage order, the two cut edges and the order in which the reference met
//    them; then the two intersection points are computed with the reference's expression.
//  * warps run a three-phase loop -- (A) advance to the next candidate that passes the
//    distance filter, (B) classify, (C) cut -- with a __syncwarp between phases, so that the
//    expensive phases execute converged instead of once per divergent lane.
//  * tiles are warps: no block-wide barrier in the main loop; CSR offsets come from a
//    warp-level decoupled look-back scan (ticket-ordered tiles).
//
// The kernel proves, per cut, that it is in the generic situation (exactly one entering and
// one leaving edge, the reference's orientation test agrees with the ring, distinct
// intersection points, no zero-length edge in the final ring).  Anything else -- only
// degenerate inputs get there -- raises OVF_ANOMALY and lv_clip_run repeats the remesh with
// the edge-list kernel, which replays the reference literally.  Results are therefore always
// the reference's.
#include "lv_clip.cuh"

typedef unsigned long long u64;

template <int MAXE, int BLOCK>
struct Ring {
    double2 *sv;
    int *sl;
    u64 ord, nxt;
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
        if (isnan(t) || isnan(rr)) rr = t + rr;
        else rr = rr > t ? rr : t;
    }
    return rr;
}

template <int MAXE, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_clip_fast(ClipArgs a, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *sv = (double2 *)smem;
    int *sl = (int *)(sv + MAXE * BLOCK);
    LvPathNode *spath = (LvPathNode *)(sl + MAXE * BLOCK);
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
        p.sv = sv; p.sl = sl; p.m = 0; p.used = 0; p.ord = 0; p.nxt = 0;
        bool alive = false;
        double2 x = make_double2(0.0, 0.0);
        if (slot < a.nslot) {
            alive = !(a.ent_label[slot] & LV_IMAGE_BIT);
            x = a.ent_xy[slot];
        }
        const bool active = alive;
        bool bad = false; // anomaly or capacity overflow for this polygon
        double prr = 0.0;
        int k1 = 0, k2 = 0, t = -1, s = 0, s_end = 0;
        if (active) {
            // reset!  polygon.jl:37-47: B->A DOWN, A->D LEFT, D->C UP, C->B RIGHT (storage order 0..3)
            p.V(0) = make_double2(g.cmaxx, g.cminy); p.L(0) = BD_DOWN;
            p.V(1) = make_double2(g.cminx, g.cminy); p.L(1) = BD_LEFT;
            p.V(2) = make_double2(g.cminx, g.cmaxy); p.L(2) = BD_UP;
            p.V(3) = make_double2(g.cmaxx, g.cmaxy); p.L(3) = BD_RIGHT;
            p.ord = 0x3210ull;
            p.nxt = 0x0321ull; // 0->1->2->3->0
            p.used = 0xfu;
            p.m = 4;
            prr = ring_influence_rr(p, x);
            if (!lv_findkey(g, x, k1, k2)) { atomicOr(&a.flags[LVF_NAN], 1); k1 = k2 = -(1 << 30); }
        }

        // ---- voronoicut!(grid, poly)  voronoigrid.jl:53-81 as a three-phase warp loop
        while (__any_sync(FULL, alive)) {
            // phase A: next candidate that passes the distance filter (voronoigrid.jl:57-75)
            bool have = false;
            double2 y = make_double2(0.0, 0.0);
            int cand = 0;
            while (alive) {
                if (s < s_end) {
                    const double2 q = a.ent_xy[s];
                    cand = s++;
                    y = lv_neighbor_pos(g, x, q);
                    const double ex = x.x - y.x, ey = x.y - y.y;
                    if ((x.x == y.x && x.y == y.y) || (ex * ex + ey * ey > prr)) continue;
                    have = true;
                    break;
                }
                if (++t >= g.npath) { alive = false; break; }
                const LvPathNode nd = spath[t];
                if (nd.rr > prr) { alive = false; break; }
                if (nd.rr > g.rr_max) { atomicOr(&a.flags[LVF_DESTROYED], 1); alive = false; break; }
                const int c1 = k1 + nd.i1, c2 = k2 + nd.i2;
                if (!(c1 >= 1 && c1 <= g.n1 && c2 >= 1 && c2 <= g.n2)) continue;
                const int lin = (c1 - 1) + g.n1 * (c2 - 1);
                s = a.cell_start[lin];
                s_end = a.cell_start[lin + 1];
            }
            __syncwarp();
            // phase B: classify every vertex against the half plane (polygon.jl:52-66)
            unsigned plus = 0, zero = 0;
            double dx = 0, dy = 0, c = 0;
            if (have) {
                dx = y.x - x.x; dy = y.y - x.y;
                const double mx = 0.5 * (y.x + x.x), my = 0.5 * (y.y + x.y);
                c = dx * mx + dy * my;
                for (unsigned u = p.used; u; u &= u - 1) {
                    const int sk = __ffs(u) - 1;
                    const double2 v = p.V(sk);
                    const double f = (dx * v.x + dy * v.y) - c;
                    if (f > SIGNUM_EPS) plus |= 1u << sk;
                    else if (!(f < -SIGNUM_EPS)) zero |= 1u << sk;
                }
            }
            __syncwarp();
            // phase C: cut (polygon.jl:59-96)
            if (plus) {
                // replay the reference's scan over the storage order on the signs only
                int mm = p.m, i = 0, a_slot = -1, b_slot = -1, nA = 0, nB = 0, last_kind = 0;
                u64 o = p.ord;
                unsigned del = 0;
                while (i < mm) {
                    const int sk = nib(o, i);
                    const int nx = nib(p.nxt, sk);
                    const bool p1 = (plus >> sk) & 1, z1 = (zero >> sk) & 1, p2 = (plus >> nx) & 1, z2 = (zero >> nx) & 1;
                    if (p2 && !p1) { a_slot = sk; nA++; last_kind = 1; } // (-,+) or (0,+): polygon is left here
                    if (p1 && !p2) { b_slot = sk; nB++; last_kind = 2; } // (+,-) or (+,0): polygon is re-entered here
                    if ((p1 && (p2 || z2)) || (z1 && p2)) {              // s1 + s2 >= 1: deleteat!  (swap-remove)
                        del |= 1u << sk;
                        o = setnib(o, i, nib(o, mm - 1));
                        mm--;
                    } else i++;
                }
                if (nA != 1 || nB != 1) bad = true;
                else {
                    const int na = nib(p.nxt, a_slot), nb = nib(p.nxt, b_slot);
                    const bool za = (zero >> a_slot) & 1, zb = (zero >> nb) & 1;
                    const double2 va1 = p.V(a_slot), va2 = p.V(na), vb1 = p.V(b_slot), vb2 = p.V(nb);
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
                        p.V(ns) = XA;   // new edge XA -> XB keeps the generator on its right
                        p.L(ns) = cand;
                        if (!za) p.nxt = setnib(p.nxt, a_slot, ns); // edge a keeps v1, its v2 becomes XA
                        else {                                      // edge a deleted: its predecessor ends at XA = va1
                            for (unsigned u = p.used & ~(1u << ns); u; u &= u - 1) {
                                const int q = __ffs(u) - 1;
                                if (nib(p.nxt, q) == a_slot) { p.nxt = setnib(p.nxt, q, ns); break; }
                            }
                        }
                        if (!zb) { p.V(b_slot) = XB; p.nxt = setnib(p.nxt, ns, b_slot); } // edge b: v1 becomes XB
                        else p.nxt = setnib(p.nxt, ns, nb);                              // edge b deleted
                        p.ord = setnib(o, mm, ns); // push!
                        p.m = mm + 1;
                        prr = ring_influence_rr(p, x); // voronoigrid.jl:76-78
                    }
                }
                if (bad) alive = false;
            }
        }
        __syncwarp();
        if (bad) atomicOr(&a.flags[LVF_OVERFLOW], OVF_ANOMALY);

        // ---- CSR offsets: warp scan + decoupled look-back over warp tiles
        const int deg = (active && !bad) ? p.m : 0;
        int inc = deg;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int v = __shfl_up_sync(FULL, inc, o2);
            if (lane >= o2) inc += v;
        }
        const int tile_total = __shfl_sync(FULL, inc, 31);
        long long prefix = 0;
        {
            volatile u64 *st = a.tile_state;
            if (lane == 0 && tile > 0) { st[tile] = TS_AGG | (u64)tile_total; }
            if (tile > 0) {
                __threadfence();
                int j = tile - 1; // lane l inspects tile j - l; every earlier tile publishes eventually
                for (;;) {
                    const int jj = j - lane;
                    u64 v = TS_INC; // "tiles" before 0: inclusive prefix 0
                    if (jj >= 0) {
                        do { v = st[jj]; } while ((v & ~TS_MASK) == 0);
                    }
                    const unsigned incm = __ballot_sync(FULL, (v & ~TS_MASK) == TS_INC);
                    const int first_inc = incm ? __ffs(incm) - 1 : 32;
                    long long contrib = (lane <= first_inc) ? (long long)(v & TS_MASK) : 0ll;
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) contrib += __shfl_down_sync(FULL, contrib, o2);
                    prefix += __shfl_sync(FULL, contrib, 0);
                    if (incm) break;
                    j -= 32;
                }
            }
            if (lane == 0) { st[tile] = TS_INC | (u64)(prefix + tile_total); __threadfence(); }
            if (lane == 0 && tile == ntiles - 1) {
                a.rowptr[a.nslot] = (int)(prefix + tile_total);
                a.flags[LVF_NNZ] = (int)(prefix + tile_total);
            }
        }
        const long long off = prefix + inc - deg;

        // ---- emit: ring walk from storage position 0 == sort_edges!  (IO.jl:35-48), area, centroid
        if (slot < a.nslot) {
            double area = 0.0, cx = 0.0, cy = 0.0;
            const bool fits = off + deg <= a.cap_nnz;
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
                        if (l >= 0) {
                            const unsigned e = a.ent_label[l];
                            cc = (e & LV_IMAGE_BIT) ? a.prim_of_label[e & ~LV_IMAGE_BIT] : l;
                        }
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
            a.area[slot] = area;
            a.cen[slot] = deg > 0 ? make_double2(cx / area, cy / area) : make_double2(0.0, 0.0);
        }
        __syncwarp();
    }
}

template <int MAXE, int BLOCK, int MINB>
static int launch_fast(LvContext *c, const ClipArgs &a) {
    const size_t smem = (size_t)MAXE * BLOCK * (sizeof(double2) + sizeof(int)) + sizeof(LvPathNode) * (size_t)c->gp.npath;
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
    return launch_fast<16, 128, 5>(c, a);
}
