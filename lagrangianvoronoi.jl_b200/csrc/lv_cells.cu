// lv_cells.cu -- K1: GPU cell-list build.
//
// Replaces the locked per-bucket pushes of the reference (voronoigrid.jl:91-100,
// neighborlist.jl:47-67) by a one-digit radix (counting) sort of (bucket key, label) pairs:
//   count -> exclusive scan -> fill -> per-bucket ascending-label order.
// Bucket keys are dense integers in [0, n1*n2) so a single counting pass is the whole radix
// sort; the per-bucket ordering pass makes the result deterministic and equal to the order
// `julia -t 1` produces (labels ascend inside a bucket).  Periodic images are separate
// (key, label) pairs computed with the reference's own arithmetic (voronoigrid.jl:130-147):
// the period is in general not a multiple of h, so image buckets are NOT translated copies
// of the primary buckets and must be keyed individually to reproduce the candidate order.
#include "lv_internal.cuh"
#include <cmath>

// position of periodic image k of x, k in the reference's insertion order
// 0: x, 1: +X, 2: -X, 3: +Y, 4: -Y, 5: +X+Y, 6: +X-Y, 7: -X+Y, 8: -X-Y   (voronoigrid.jl:130-147)
__device__ __forceinline__ bool lv_image(const LvGridParams &g, double2 x, int k, double2 &out) {
    const double px = g.xperiod, py = g.yperiod;
    switch (k) {
    case 0: out = x; return true;
    case 1: out = make_double2(x.x + px * 1.0, x.y + px * 0.0); return g.xper;
    case 2: out = make_double2(x.x - px * 1.0, x.y - px * 0.0); return g.xper;
    case 3: out = make_double2(x.x + py * 0.0, x.y + py * 1.0); return g.yper;
    case 4: out = make_double2(x.x - py * 0.0, x.y - py * 1.0); return g.yper;
    case 5: out = make_double2((x.x + px * 1.0) + py * 0.0, (x.y + px * 0.0) + py * 1.0); return g.xper && g.yper;
    case 6: out = make_double2((x.x + px * 1.0) - py * 0.0, (x.y + px * 0.0) - py * 1.0); return g.xper && g.yper;
    case 7: out = make_double2((x.x - px * 1.0) + py * 0.0, (x.y - px * 0.0) + py * 1.0); return g.xper && g.yper;
    default: out = make_double2((x.x - px * 1.0) - py * 0.0, (x.y - px * 0.0) - py * 1.0); return g.xper && g.yper;
    }
}

// linear bucket index of image k of generator x, or -1 when the reference drops it
// (neighborlist.jl:57,65-66).  The primary image (k = 0) of a generator that lies outside the
// cell list goes to the extra bucket `ncell` so that the generator still gets a slot: the
// reference still clips such a polygon (voronoigrid.jl:102-105), nobody sees it as neighbour.
template <bool FILL>
__global__ void __launch_bounds__(256) k_cells_count_fill(LvGridParams g, int64_t n, const double2 *__restrict__ xy,
                                                          int *__restrict__ cnt, const int *__restrict__ start,
                                                          unsigned *__restrict__ ent_label, int *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 x = xy[i];
    const int ncell = g.n1 * g.n2;
    const int nimg = (g.xper || g.yper) ? 9 : 1;
    for (int k = 0; k < nimg; k++) {
        double2 p;
        if (!lv_image(g, x, k, p)) continue;
        int i1, i2;
        if (!lv_findkey(g, p, i1, i2)) { // NaN / Inf coordinate: floor(Int, .) throws in the reference
            if (!FILL) atomicOr(&flags[LVF_NAN], 1);
            if (k == 0) i1 = i2 = -1; else continue;
        }
        int lin;
        if (i1 >= 1 && i1 <= g.n1 && i2 >= 1 && i2 <= g.n2) lin = (i1 - 1) + g.n1 * (i2 - 1);
        else if (k == 0) lin = ncell;
        else continue;
        if (!FILL) atomicAdd(&cnt[lin], 1);
        else {
            int pos = start[lin] + atomicAdd(&cnt[lin], 1);
            ent_label[pos] = (unsigned)i | (k ? LV_IMAGE_BIT : 0u);
        }
    }
}

// One thread per bucket: ascending label order (the `julia -t 1` order), then the slot-ordered
// copy of the generator positions and the label -> primary-slot map.
__global__ void __launch_bounds__(128) k_cells_order(int ncell_ext, const int *__restrict__ start,
                                                     unsigned *__restrict__ ent_label, double2 *__restrict__ ent_xy,
                                                     const double2 *__restrict__ xy, int *__restrict__ prim_of_label,
                                                     const unsigned char *__restrict__ owned_mask, unsigned char *__restrict__ own,
                                                     const int *__restrict__ order_key, int owned_count) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= ncell_ext) return;
    const int s0 = start[b], s1 = start[b + 1];
    const int k = s1 - s0;
    if (k <= 0) return;
    const unsigned M = ~LV_IMAGE_BIT;
    // sort key: (label, image bit); in the multi-GPU decomposition the label is replaced by the generator's
    // GLOBAL label (order_key) so that every rank orders a bucket exactly like the single-GPU run does
    auto key_of = [order_key](unsigned e) -> unsigned {
        const unsigned lab = e & ~LV_IMAGE_BIT;
        const unsigned base = order_key ? (unsigned)order_key[lab] : lab;
        return (base << 1) | (e >> 31);
    };
    if (k <= 16) {
        unsigned v[16], kk[16];
#pragma unroll
        for (int a = 0; a < 16; a++) {
            v[a] = a < k ? ent_label[s0 + a] : 0xffffffffu;
            kk[a] = a < k ? key_of(v[a]) : 0xffffffffu;
        }
        // odd-even transposition network: branch-free, registers only
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int a = (r & 1); a + 1 < 16; a += 2) {
                const bool sw = kk[a + 1] < kk[a];
                const unsigned lo = sw ? v[a + 1] : v[a], hi = sw ? v[a] : v[a + 1];
                const unsigned klo = sw ? kk[a + 1] : kk[a], khi = sw ? kk[a] : kk[a + 1];
                v[a] = lo; v[a + 1] = hi;
                kk[a] = klo; kk[a + 1] = khi;
            }
        }
#pragma unroll
        for (int a = 0; a < 16; a++)
            if (a < k) ent_label[s0 + a] = v[a];
    } else { // heap sort in place (crowded bucket)
        unsigned *A = ent_label + s0;
        auto key = [&](unsigned v) { return key_of(v); };
        for (int st = k / 2 - 1; st >= 0; st--) {
            int r = st;
            for (;;) {
                int ch = 2 * r + 1;
                if (ch >= k) break;
                if (ch + 1 < k && key(A[ch + 1]) > key(A[ch])) ch++;
                if (key(A[ch]) > key(A[r])) { unsigned t = A[ch]; A[ch] = A[r]; A[r] = t; r = ch; } else break;
            }
        }
        for (int end = k - 1; end > 0; end--) {
            unsigned t = A[0]; A[0] = A[end]; A[end] = t;
            int r = 0;
            for (;;) {
                int ch = 2 * r + 1;
                if (ch >= end) break;
                if (ch + 1 < end && key(A[ch + 1]) > key(A[ch])) ch++;
                if (key(A[ch]) > key(A[r])) { unsigned t2 = A[ch]; A[ch] = A[r]; A[r] = t2; r = ch; } else break;
            }
        }
    }
    for (int a = 0; a < k; a++) {
        unsigned e = ent_label[s0 + a];
        unsigned lab = e & M;
        ent_xy[s0 + a] = xy[lab];
        const bool prim = !(e & LV_IMAGE_BIT);
        if (prim) prim_of_label[lab] = s0 + a;
        own[s0 + a] = prim && (owned_mask ? owned_mask[lab] != 0 : (owned_count < 0 || (int)lab < owned_count));
    }
}

// ---- exclusive scan (int32), three kernels: block sums, scan of block sums, final ----------
#define SCAN_BLOCK 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)

__device__ __forceinline__ int block_exclusive_scan(int v, int &total, int *sm /*>= 32 ints*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int s = lane < nw ? sm[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        sm[lane] = si - s; // exclusive warp offsets
        if (lane == 31) sm[32] = si;
    }
    __syncthreads();
    total = sm[32];
    int r = sm[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(const int *__restrict__ in, int64_t n, int *__restrict__ sums) {
    __shared__ int sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + (int64_t)k * SCAN_BLOCK + threadIdx.x;
        if (i < n) s += in[i];
    }
    int total;
    block_exclusive_scan(s, total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_top(int *sums, int nblk, int *total_out) {
    __shared__ int sm[33];
    int carry = 0;
    for (int b0 = 0; b0 < nblk; b0 += 1024) {
        int i = b0 + threadIdx.x;
        int v = i < nblk ? sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, total, sm);
        if (i < nblk) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_final(const int *__restrict__ in, int *__restrict__ out, int64_t n,
                                                           const int *__restrict__ sums, const int *__restrict__ total) {
    __shared__ int sm[33];
    // thread t owns SCAN_ITEMS consecutive items so the per-thread scan is sequential
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    int tot;
    int ex = block_exclusive_scan(s, tot, sm) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

int lv_exclusive_scan_i32(LvContext *c, const int *in, int *out, int64_t n) {
    int nblk = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (nblk < 1) nblk = 1;
    LV_TRY(lv_ensure(c, &c->d_scratch, &c->cap_scratch, (int64_t)(nblk + 4) * sizeof(int), 1));
    int *sums = (int *)c->d_scratch;
    int *total = sums + nblk;
    k_scan_sums<<<nblk, SCAN_BLOCK, 0, c->stream>>>(in, n, sums);
    k_scan_top<<<1, 1024, 0, c->stream>>>(sums, nblk, total);
    k_scan_final<<<nblk, SCAN_BLOCK, 0, c->stream>>>(in, out, n, sums, total);
    c->launches += 3;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// ---- driver -----------------------------------------------------------------------------------
int lv_cells_build(LvContext *c) {
    LvProfScope prof(c, LV_PROF_CELLS);
    const int64_t n = c->n;
    const int64_t ncell_ext = c->ncell + 1; // + the "outside" bucket
    cudaStream_t st = c->stream;
    LV_CUDA(c, cudaMemsetAsync(c->d_cell_cnt, 0, sizeof(int) * (size_t)(ncell_ext + 1), st));
    LV_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 8, st));
    const int nb = (int)((n + 255) / 256);
    if (n > 0) {
        k_cells_count_fill<false><<<nb, 256, 0, st>>>(c->gp, n, c->xy, c->d_cell_cnt, nullptr, nullptr, c->d_flags);
        c->launches++;
    }
    LV_TRY(lv_exclusive_scan_i32(c, c->d_cell_cnt, c->d_cell_start, ncell_ext));
    // slot count is needed on the host to size the slot-indexed buffers
    int nslot = 0;
    LV_TRY(lv_publish_flags(c, c->d_cell_start + ncell_ext));
    nslot = c->h_flags[8];
    if (c->h_flags[LVF_NAN]) {
        // error path only: name the first offending generator (in strip mode: owned or ghost) -- positions come to the host
        std::vector<double2> h((size_t)n);
        long long bad = -1;
        if (cudaMemcpy(h.data(), c->xy, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int64_t i = 0; i < n && bad < 0; i++)
                if (!std::isfinite(h[(size_t)i].x) || !std::isfinite(h[(size_t)i].y)) bad = i;
        if (c->strip.on)
            return lv_set_error(c, LV_ENAN, "generator position is NaN or Inf (local generator %lld of %lld; %lld owned -> %s; last migration out %d in %d)",
                                bad, (long long)n, (long long)c->strip.n_own, bad >= 0 && bad < c->strip.n_own ? "owned" : "ghost",
                                c->strip.last_mig_out, c->strip.last_mig_in);
        return lv_set_error(c, LV_ENAN, "generator position is NaN or Inf (generator %lld of %lld)", bad + 1, (long long)n);
    }
    c->nslot = nslot;
    {
        int64_t cap = c->cap_slot;
        int64_t need = nslot + 1;
        if (need > cap) {
            int64_t ncap = need + need / 16 + 1024;
            int64_t c1 = cap, c2 = cap, c3 = cap, c4 = cap, c5 = cap, c6 = cap, c7 = cap;
            LV_TRY(lv_ensure(c, (void **)&c->d_ent_label, &c1, ncap, sizeof(unsigned)));
            LV_TRY(lv_ensure(c, (void **)&c->d_ent_xy, &c2, ncap, sizeof(double2)));
            LV_TRY(lv_ensure(c, (void **)&c->d_rowptr, &c3, ncap + 1, sizeof(int)));
            LV_TRY(lv_ensure(c, (void **)&c->d_deg, &c6, ncap + 1, sizeof(unsigned char)));
            LV_TRY(lv_ensure(c, (void **)&c->d_own, &c7, ncap + 1, sizeof(unsigned char)));
            LV_TRY(lv_ensure(c, (void **)&c->d_area, &c4, ncap, sizeof(double)));
            LV_TRY(lv_ensure(c, (void **)&c->d_cen, &c5, ncap, sizeof(double2)));
            c->cap_slot = ncap;
        }
    }
    LV_CUDA(c, cudaMemsetAsync(c->d_cell_cnt, 0, sizeof(int) * (size_t)(ncell_ext + 1), st));
    if (n > 0) LV_CUDA(c, cudaMemsetAsync(c->d_prim_of_label, 0xff, sizeof(int) * (size_t)n, st)); // -1: no primary slot here
    if (n > 0) {
        k_cells_count_fill<true><<<nb, 256, 0, st>>>(c->gp, n, c->xy, c->d_cell_cnt, c->d_cell_start, c->d_ent_label,
                                                     c->d_flags);
        k_cells_order<<<(int)((ncell_ext + 127) / 128), 128, 0, st>>>((int)ncell_ext, c->d_cell_start, c->d_ent_label,
                                                                      c->d_ent_xy, c->xy, c->d_prim_of_label, c->owned_mask, c->d_own, c->order_key,
                                                                      (int)c->owned_count);
        c->launches += 2;
    }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}
