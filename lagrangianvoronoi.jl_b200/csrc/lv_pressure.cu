// lv_pressure.cu -- K3/K4/K5: the semi-implicit pressure system on the Voronoi stencil.
//
// Replaces pressure.jl:89-225 of the reference: operator assembly (refresh!(A,...),
// :104-117), the matrix-free matvec mul! (:119-130), the right-hand side (refresh!(solver,...),
// :162-203) and the find_pressure! fixed-point loop (:215-225) whose inner Krylov.jl MINRES call
// is replaced by an FP64 conjugate-gradient iteration that never leaves the device.
// All vectors live in slot order (see lv_internal.cuh); image slots are empty rows (diag = 0,
// b = 0, x = 0) and take no part in the system.
#include "lv_internal.cuh"

#define PR_BLOCK 256

// scalars kept on the device between kernels (doubles in c->d_red[0..15])
enum { SC_RR = 0, SC_ALPHA = 1, SC_BETA = 2, SC_TOL = 3, SC_PAP = 4, SC_RR0 = 5, SC_BNORM2 = 6, SC_RES2 = 7, SC_ITER = 8, SC_CONV = 9, SC_TMP0 = 10, SC_TMP1 = 11, SC_DEAD = 13 /* a peer wait timed out */, SC_RZ = 14 /* r.z of the preconditioned iteration (= r.r without preconditioner) */, SC_SOLVED = 12 /* MINRES: stats.solved (tolerance met; not the ill-conditioned / NaN exits) */,
       // MINRES (Paige-Saunders, in the formulation of Krylov.jl 0.9.8 minres!) scalar state
       MR_BETA = 16, MR_OLDB = 17, MR_DBAR = 18, MR_EPS = 19, MR_CS = 20, MR_SN = 21, MR_PHIBAR = 22, MR_GAMMA = 23, MR_PHI = 24,
       MR_DELTA = 25, MR_ANORM2 = 26, MR_GMAX = 27, MR_GMIN = 28, MR_XENORM2 = 29, MR_ROOT = 30, MR_BETA1 = 31, MR_ERRV = 32 /* ..36 */,
       SC_COUNT = 64 };

// ---- label <-> slot permutation --------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(PR_BLOCK) k_gather_slots(int nslot, const unsigned *__restrict__ ent_label,
                                                           const double *__restrict__ src, double *__restrict__ dst, double fill) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslot) return;
    const unsigned e = ent_label[s];
    if (e & LV_IMAGE_BIT) {
#pragma unroll
        for (int k = 0; k < NC; k++) dst[(size_t)NC * s + k] = fill;
    } else {
#pragma unroll
        for (int k = 0; k < NC; k++) dst[(size_t)NC * s + k] = src[(size_t)NC * e + k];
    }
}
template <int NC>
__global__ void __launch_bounds__(PR_BLOCK) k_scatter_labels(int64_t n, const int *__restrict__ prim, const double *__restrict__ src,
                                                             double *__restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = prim[i];
#pragma unroll
    for (int k = 0; k < NC; k++) dst[(size_t)NC * i + k] = s >= 0 ? src[(size_t)NC * s + k] : 0.0;
}

int lv_gather_to_slots(LvContext *c, const double *src, double *dst, int ncomp, double fill) {
    const int ns = (int)c->nslot;
    if (ns == 0) return LV_OK;
    const int nb = (ns + PR_BLOCK - 1) / PR_BLOCK;
    if (ncomp == 1) k_gather_slots<1><<<nb, PR_BLOCK, 0, c->stream>>>(ns, c->d_ent_label, src, dst, fill);
    else k_gather_slots<2><<<nb, PR_BLOCK, 0, c->stream>>>(ns, c->d_ent_label, src, dst, fill);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}
int lv_scatter_to_labels(LvContext *c, const double *src, double *dst, int ncomp) {
    if (c->n == 0) return LV_OK;
    const int nb = (int)((c->n + PR_BLOCK - 1) / PR_BLOCK);
    if (ncomp == 1) k_scatter_labels<1><<<nb, PR_BLOCK, 0, c->stream>>>(c->n, c->d_prim_of_label, src, dst);
    else k_scatter_labels<2><<<nb, PR_BLOCK, 0, c->stream>>>(c->n, c->d_prim_of_label, src, dst);
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// ---- workspace ---------------------------------------------------------------------------------
int lv_pr_ensure(LvContext *c) {
    if (!c->mesh_valid) return lv_set_error(c, LV_EINVAL, "no valid mesh: call lv_remesh first");
    const int64_t need = c->cap_slot;
    if (c->pr_cap < need || !c->d_mass) {
        double **one[] = {&c->d_mass, &c->d_rho, &c->d_c2, &c->d_P, &c->d_diag, &c->d_dinv, &c->d_b, &c->d_bvel, &c->d_vec[0], &c->d_vec[1],
                          &c->d_vec[2], &c->d_vec[3], &c->d_vec[4], &c->d_vec[5], &c->d_vec[6], &c->d_vec[7]};
        for (double **p : one) {
            int64_t cap = *p ? c->pr_cap : 0;
            LV_TRY(lv_ensure(c, (void **)p, &cap, need, sizeof(double)));
        }
        int64_t cap = c->d_v ? c->pr_cap : 0;
        LV_TRY(lv_ensure(c, (void **)&c->d_v, &cap, need, sizeof(double2)));
        cap = c->d_GP ? c->pr_cap : 0;
        LV_TRY(lv_ensure(c, (void **)&c->d_GP, &cap, need, sizeof(double2)));
        c->pr_cap = need;
        // neutral defaults so that image slots never produce NaN
        LV_CUDA(c, cudaMemsetAsync(c->d_mass, 0, sizeof(double) * (size_t)need, c->stream));
        LV_CUDA(c, cudaMemsetAsync(c->d_P, 0, sizeof(double) * (size_t)need, c->stream));
        LV_CUDA(c, cudaMemsetAsync(c->d_v, 0, sizeof(double2) * (size_t)need, c->stream));
        c->pr_valid = false;
    }
    if (c->cap_w < c->cap_nnz || !c->d_w) {
        int64_t c1 = c->cap_w, c2 = c->cap_w, c3 = c->cap_w, c4 = c->cap_w;
        LV_TRY(lv_ensure(c, (void **)&c->d_w, &c1, c->cap_nnz, sizeof(double)));
        LV_TRY(lv_ensure(c, (void **)&c->d_lrr, &c2, c->cap_nnz, sizeof(double)));
        LV_TRY(lv_ensure(c, (void **)&c->d_mx, &c3, c->cap_nnz, sizeof(double2)));
        LV_TRY(lv_ensure(c, (void **)&c->d_mz, &c4, c->cap_nnz, sizeof(double2)));
        c->cap_w = c->cap_nnz;
        c->assembled = false;
    }
    if (!c->d_red) LV_TRY(lv_alloc(c, (void **)&c->d_red, sizeof(double) * (SC_COUNT + 2 * 4096)));
    return LV_OK;
}

// ---- K3: operator assembly  pressure.jl:104-117 ------------------------------------------------
// Besides A.diagonal and the weights w = lrr*(0.5/rho_i + 0.5/rho_j) the kernel keeps, per edge, the three
// geometric factors every later sweep of find_pressure! needs -- lrr = lr_ratio(p.x - y, e) (polygon.jl:228),
// m - p.x and m - z (pressure.jl:176,178,196,198) -- so that the 10 fixed-point passes neither repeat the
// FP64 divide/sqrt nor re-gather the neighbour positions.  Values are the reference's expressions.
__global__ void __launch_bounds__(PR_BLOCK) k_assemble(LvGridParams g, int nslot, double dt, const unsigned char *__restrict__ own,
                                                       const double2 *__restrict__ ent_xy, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                       const int *__restrict__ col, const double2 *__restrict__ v1,
                                                       const double2 *__restrict__ v2, const double *__restrict__ mass,
                                                       const double *__restrict__ rho, const double *__restrict__ c2,
                                                       double *__restrict__ diag, double *__restrict__ w, double *__restrict__ lrr_out,
                                                       double2 *__restrict__ mx_out, double2 *__restrict__ mz_out, float *__restrict__ dinv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslot) return;
    if (!own[i]) { diag[i] = 0.0; dinv[i] = 0.0f; return; }
    const double ri = rho[i];
    const double dg = mass[i] / (((ri * ri) * c2[i]) * (dt * dt)); // pressure.jl:110
    diag[i] = dg;
    double aii = dg; // A_ii = diagonal + sum of the row's weights: the Jacobi preconditioner of LV_SOLVER_PCG
    const double2 x = ent_xy[i];
    const int r0 = rowptr[i], r1 = r0 + rdeg[i];
    for (int k = r0; k < r1; k++) {
        const int j = col[k];
        const double2 a = v1[k], b = v2[k];
        const double mx = 0.5 * (a.x + b.x), my = 0.5 * (a.y + b.y); // midpoint(e)  geometry.jl:145-147
        mx_out[k] = make_double2(mx - x.x, my - x.y);
        if (j < 0) { w[k] = 0.0; lrr_out[k] = 0.0; mz_out[k] = make_double2(0.0, 0.0); continue; } // wall edge: not in neighbors(p, grid)
        const double2 y = lv_neighbor_pos(g, x, ent_xy[j]);
        const double ex = a.x - b.x, ey = a.y - b.y, dx = x.x - y.x, dy = x.y - y.y;
        const double lrr = sqrt((ex * ex + ey * ey) / (dx * dx + dy * dy)); // lr_ratio  polygon.jl:228-232
        lrr_out[k] = lrr;
        const double wk = lrr * (0.5 / ri + 0.5 / rho[j]);                  // pressure.jl:113
        w[k] = wk;
        aii += wk;
        const double zx = 0.5 * (x.x + y.x), zy = 0.5 * (x.y + y.y);        // midpoint(p.x, y)  pressure.jl:196
        mz_out[k] = make_double2(mx - zx, my - zy);
    }
    // The preconditioner M^-1 = diag(1/A_ii) is kept in single precision: any positive diagonal is a valid (SPD) Jacobi
    // preconditioner, the Krylov arithmetic stays FP64, and the two update kernels read 4 instead of 8 bytes per cell for it.
    dinv[i] = aii > 0.0 ? (float)(1.0 / aii) : 0.0f;
}

// Edge-parallel form of k_assemble (same values, bit for bit).  The rows of an aligned group of 32 slots are contiguous in
// the edge arrays, in slot order (every clipping kernel places the rows of a warp / block with one prefix sum), so a warp
// walks the group's edges with consecutive lanes on consecutive edges: every load of col / v1 / v2 and every store of
// w / lrr / mx / mz is coalesced (the row-parallel kernel writes 6 consecutive entries per lane, i.e. 96-byte strides).
// The owner row of an edge comes from a 5-step search over the group's 32 row offsets in shared memory; the row's own
// position and density travel by shuffle; the row sums A_ii are taken from staged weights in edge order, as before.
#define AS_MAXT 512 // staged weights per group; a group with more edges (polygons of the edge-list kernel) takes the row loop
__global__ void __launch_bounds__(PR_BLOCK) k_assemble_tile(LvGridParams g, int nslot, double dt, const unsigned char *__restrict__ own,
                                                            const double2 *__restrict__ ent_xy, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                            const int *__restrict__ col, const double2 *__restrict__ v1,
                                                            const double2 *__restrict__ v2, const double *__restrict__ mass,
                                                            const double *__restrict__ rho, const double *__restrict__ c2,
                                                            double *__restrict__ diag, double *__restrict__ w, double *__restrict__ lrr_out,
                                                            double2 *__restrict__ mx_out, double2 *__restrict__ mz_out, float *__restrict__ dinv, int maxt) {
    __shared__ int s_off[PR_BLOCK / 32][32];
    __shared__ double s_w[PR_BLOCK / 32][AS_MAXT];
    const unsigned FULL = 0xffffffffu;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const bool mine = i < nslot && own[i];
    const int d = mine ? rdeg[i] : 0;
    const int r0 = mine ? rowptr[i] : 0;
    double ri = 1.0, dg = 0.0;
    double2 x = make_double2(0.0, 0.0);
    if (mine) {
        ri = rho[i];
        dg = mass[i] / (((ri * ri) * c2[i]) * (dt * dt)); // pressure.jl:110
        x = ent_xy[i];
    }
    int off = d;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int t = __shfl_up_sync(FULL, off, o2);
        if (lane >= o2) off += t;
    }
    const int T = __shfl_sync(FULL, off, 31);
    off -= d;
    const unsigned has = __ballot_sync(FULL, d > 0);
    if (i < nslot && !mine) { diag[i] = 0.0; dinv[i] = 0.0f; }
    if (T == 0) {
        if (mine) { diag[i] = dg; dinv[i] = dg > 0.0 ? (float)(1.0 / dg) : 0.0f; }
        return;
    }
    const int E0 = __shfl_sync(FULL, r0 - off, __ffs(has) - 1); // first edge of the group
    // rows of the group must be contiguous in slot order (see above); anything else takes the row loop
    const bool contiguous = __all_sync(FULL, d == 0 || r0 == E0 + off);
    double aii = dg;
    if (contiguous && T <= maxt) {
        s_off[wp][lane] = off;
        __syncwarp();
        for (int e = lane; e < ((T + 31) & ~31); e += 32) {
            int lo = 0;
#pragma unroll
            for (int st = 16; st; st >>= 1)
                if (s_off[wp][lo + st] <= (e < T ? e : T - 1)) lo += st;
            const double xx = __shfl_sync(FULL, x.x, lo), xy_ = __shfl_sync(FULL, x.y, lo), rio = __shfl_sync(FULL, ri, lo);
            if (e >= T) continue;
            const int k = E0 + e;
            const int j = col[k];
            const double2 a = v1[k], b = v2[k];
            const double mx = 0.5 * (a.x + b.x), my = 0.5 * (a.y + b.y); // midpoint(e)  geometry.jl:145-147
            mx_out[k] = make_double2(mx - xx, my - xy_);
            double wk = 0.0;
            if (j < 0) { w[k] = 0.0; lrr_out[k] = 0.0; mz_out[k] = make_double2(0.0, 0.0); } // wall edge: not in neighbors(p, grid)
            else {
                const double2 xo = make_double2(xx, xy_);
                const double2 y = lv_neighbor_pos(g, xo, ent_xy[j]);
                const double ex = a.x - b.x, ey = a.y - b.y, dx = xx - y.x, dy = xy_ - y.y;
                const double lrr = sqrt((ex * ex + ey * ey) / (dx * dx + dy * dy)); // lr_ratio  polygon.jl:228-232
                lrr_out[k] = lrr;
                wk = lrr * (0.5 / rio + 0.5 / rho[j]);                              // pressure.jl:113
                w[k] = wk;
                const double zx = 0.5 * (xx + y.x), zy = 0.5 * (xy_ + y.y);         // midpoint(p.x, y)  pressure.jl:196
                mz_out[k] = make_double2(mx - zx, my - zy);
            }
            s_w[wp][e] = wk;
        }
        __syncwarp();
        for (int k = 0; k < d; k++) aii += s_w[wp][off + k]; // wall edges contribute 0.0, exactly like the row loop skipping them
    } else if (mine) {
        for (int k = r0; k < r0 + d; k++) {
            const int j = col[k];
            const double2 a = v1[k], b = v2[k];
            const double mx = 0.5 * (a.x + b.x), my = 0.5 * (a.y + b.y);
            mx_out[k] = make_double2(mx - x.x, my - x.y);
            if (j < 0) { w[k] = 0.0; lrr_out[k] = 0.0; mz_out[k] = make_double2(0.0, 0.0); continue; }
            const double2 y = lv_neighbor_pos(g, x, ent_xy[j]);
            const double ex = a.x - b.x, ey = a.y - b.y, dx = x.x - y.x, dy = x.y - y.y;
            const double lrr = sqrt((ex * ex + ey * ey) / (dx * dx + dy * dy));
            lrr_out[k] = lrr;
            const double wk = lrr * (0.5 / ri + 0.5 / rho[j]);
            w[k] = wk;
            aii += wk;
            const double zx = 0.5 * (x.x + y.x), zy = 0.5 * (x.y + y.y);
            mz_out[k] = make_double2(mx - zx, my - zy);
        }
    }
    if (mine) {
        diag[i] = dg;
        dinv[i] = aii > 0.0 ? (float)(1.0 / aii) : 0.0f;
    }
}

int lv_pr_assemble(LvContext *c, double dt) {
    LV_TRY(lv_pr_ensure(c));
    if (!c->pr_valid) return lv_set_error(c, LV_EINVAL, "fields not uploaded: call lv_fields_upload first");
    LvProfScope prof(c, LV_PROF_ASSEMBLE);
    const int ns = (int)c->nslot;
    if (ns > 0) {
        static const bool rows = [] { const char *e = getenv("LV_ASSEMBLE"); return e && !strcmp(e, "rows"); }(); // A/B switch
        // test hook: groups with more than LV_ASSEMBLE_MAXT edges take the kernel's row loop (default: the staging capacity)
        int maxt = AS_MAXT;
        if (const char *e = getenv("LV_ASSEMBLE_MAXT")) { const int v = atoi(e); if (v >= 0 && v < AS_MAXT) maxt = v; }
        if (rows)
            k_assemble<<<(ns + PR_BLOCK - 1) / PR_BLOCK, PR_BLOCK, 0, c->stream>>>(c->gp, ns, dt, c->d_own, c->d_ent_xy, c->d_rowptr, c->d_deg,
                                                                                 c->d_col, c->d_v1, c->d_v2, c->d_mass, c->d_rho, c->d_c2,
                                                                                 c->d_diag, c->d_w, c->d_lrr, c->d_mx, c->d_mz, (float *)c->d_dinv);
        else
            k_assemble_tile<<<(ns + PR_BLOCK - 1) / PR_BLOCK, PR_BLOCK, 0, c->stream>>>(c->gp, ns, dt, c->d_own, c->d_ent_xy, c->d_rowptr, c->d_deg,
                                                                                      c->d_col, c->d_v1, c->d_v2, c->d_mass, c->d_rho, c->d_c2,
                                                                                      c->d_diag, c->d_w, c->d_lrr, c->d_mx, c->d_mz, (float *)c->d_dinv, maxt);
        c->launches++;
        LV_CUDA(c, cudaGetLastError());
    }
    c->assembled = true;
    c->bvel_valid = false;
    c->asm_dt = dt;
    return LV_OK;
}

// ---- block reduction helper ---------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *sm /*>=32*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r; // valid in thread 0
}

// one block: finish reductions and update the scalars.  mode 0: init, 1: alpha, 2: beta, 3: residual check
// stage 0: reduce the partials and update (single GPU); stage 1: reduce only, leaving the two local sums in
// scal[SC_TMP0..1] for the NCCL allreduce; stage 2: update from the (now global) sums in scal[SC_TMP0..1].
// stage 3 (multi-GPU with mapped mailboxes): reduce, exchange the two local sums with every rank through the
// NVLink mailboxes (rank-ordered, deterministic sum) and update -- one launch, no NCCL in the Krylov loop.
// The mailbox exchange always runs, also after convergence, so that all ranks keep the same sequence numbers.
struct MailArgs { LvMailSlot *const *boxes; int nranks, rank, seq; int *dead; };
// Called by all 256 threads of ONE block: the stand-alone kernel k_cg_scalars, or the block of a producer kernel that
// arrives last (lv_last_block) -- then no separate launch sits between the producer and the consumer of the scalars.
__device__ __noinline__ void cg_finish(int mode, int stage, int nblk, int nblk_max, const double *partial, double *scal, double rtol,
                                       double atol, MailArgs mail) {
    __shared__ double sm[32];
    __shared__ double sh[2], m0[LV_MB_MAX_RANKS], m1[LV_MB_MAX_RANKS];
    if (stage == 3) {
        double a = 0.0, b2 = 0.0;
        for (int k = threadIdx.x; k < nblk; k += blockDim.x) { a += __ldcg(&partial[k]); if (mode == 0 || mode == 3 || mode >= 8) b2 += __ldcg(&partial[nblk_max + k]); }
        const double t1 = block_sum(a, sm);
        const double t2 = block_sum(b2, sm);
        if (threadIdx.x == 0) { sh[0] = t1; sh[1] = t2; }
        __syncthreads();
        const int q = threadIdx.x, par = mail.seq & 1;
        if (q < mail.nranks) {
            LvMailSlot *dst = mail.boxes[q] + par * LV_MB_MAX_RANKS + mail.rank;
            dst->v[0] = sh[0];
            dst->v[1] = sh[1];
            __threadfence_system();
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(&dst->flag), "r"(mail.seq) : "memory");
            LvMailSlot *src = mail.boxes[mail.rank] + par * LV_MB_MAX_RANKS + q;
            lv_wait_ge(&src->flag, mail.seq, mail.dead); // bounded: a dead peer must not hang the GPU
            m0[q] = __ldcv(&src->v[0]);
            m1[q] = __ldcv(&src->v[1]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double g1 = 0.0, g2 = 0.0;
            for (int k = 0; k < mail.nranks; k++) { g1 += m0[k]; g2 += m1[k]; }
            scal[SC_TMP0] = g1;
            scal[SC_TMP1] = g2;
        }
        __syncthreads();
        stage = 2; // fall through to the update from scal[SC_TMP0..1]
    }
    if (stage != 1 && mode != 0 && mode != 3 && mode != 4 && mode != 8 && scal[SC_CONV] != 0.0) return;
    double s1, s2;
    if (stage != 2) {
        double a = 0.0, b2 = 0.0;
        for (int k = threadIdx.x; k < nblk; k += blockDim.x) { a += __ldcg(&partial[k]); if (mode == 0 || mode == 3 || mode >= 8) b2 += __ldcg(&partial[nblk_max + k]); }
        s1 = block_sum(a, sm);
        s2 = block_sum(b2, sm);
        if (stage == 1) {
            if (threadIdx.x == 0) { scal[SC_TMP0] = s1; scal[SC_TMP1] = s2; }
            return;
        }
    } else { s1 = scal[SC_TMP0]; s2 = scal[SC_TMP1]; }
    if (threadIdx.x != 0) return;
    if (mode == 8 || mode == 9) { // Jacobi-preconditioned CG: s1 = r.r (stopping test, unpreconditioned), s2 = r.z with z = D^-1 r
        if (mode == 8) {
            scal[SC_RR] = s1; scal[SC_RR0] = s1; scal[SC_RZ] = s2;
            const double tol = atol + rtol * sqrt(s1);
            scal[SC_TOL] = tol;
            scal[SC_ITER] = 0.0;
            scal[SC_CONV] = (sqrt(s1) <= tol) ? 1.0 : 0.0;
        } else {
            scal[SC_BETA] = s2 / scal[SC_RZ];
            scal[SC_RZ] = s2;
            scal[SC_RR] = s1;
            scal[SC_ITER] += 1.0;
            if (sqrt(s1) <= scal[SC_TOL] || !(s1 == s1)) scal[SC_CONV] = 1.0;
        }
    } else if (mode == 0) {
        scal[SC_RZ] = s1;
        scal[SC_RR] = s1; scal[SC_RR0] = s1; scal[SC_BNORM2] = s2;
        const double tol = atol + rtol * sqrt(s1); // Krylov: eps = atol + rtol*||r0||
        scal[SC_TOL] = tol;
        scal[SC_ITER] = 0.0;
        scal[SC_CONV] = (sqrt(s1) <= tol) ? 1.0 : 0.0;
    } else if (mode == 1) {
        scal[SC_PAP] = s1;
        scal[SC_ALPHA] = scal[SC_RZ] / s1;
    } else if (mode == 2) {
        const double rr_old = scal[SC_RR];
        scal[SC_BETA] = s1 / rr_old;
        scal[SC_RR] = s1;
        scal[SC_RZ] = s1;
        scal[SC_ITER] += 1.0;
        if (sqrt(s1) <= scal[SC_TOL] || !(s1 == s1)) scal[SC_CONV] = 1.0;
    } else if (mode == 3) {
        scal[SC_RES2] = s1; scal[SC_BNORM2] = s2;
    } else if (mode == 4) { // MINRES init: beta1 = ||r0||
        const double beta1 = sqrt(s1);
        scal[MR_BETA1] = beta1; scal[MR_BETA] = beta1; scal[MR_OLDB] = 0.0; scal[MR_DBAR] = 0.0; scal[MR_EPS] = 0.0;
        scal[MR_CS] = -1.0; scal[MR_SN] = 0.0; scal[MR_PHIBAR] = beta1; scal[MR_ANORM2] = 0.0; scal[MR_GMAX] = 0.0;
        scal[MR_GMIN] = __longlong_as_double(0x7ff0000000000000ll); scal[MR_XENORM2] = 0.0;
        for (int k = 0; k < 5; k++) scal[MR_ERRV + k] = 0.0;
        scal[SC_TOL] = atol + rtol * beta1;
        scal[SC_ITER] = 0.0;
        scal[SC_CONV] = (beta1 == 0.0 || beta1 <= scal[SC_TOL]) ? 1.0 : 0.0;
        scal[SC_SOLVED] = scal[SC_CONV];
    } else if (mode == 5) { // alpha = v.y / beta ; delta = cs*dbar + sn*alpha
        const double alpha = s1 / scal[MR_BETA];
        scal[SC_ALPHA] = alpha;
        scal[MR_DELTA] = scal[MR_CS] * scal[MR_DBAR] + scal[MR_SN] * alpha;
    } else if (mode == 6) { // new beta, plane rotation
        const double alpha = scal[SC_ALPHA], oldb = scal[MR_BETA], cs = scal[MR_CS], sn = scal[MR_SN], dbar = scal[MR_DBAR];
        const double beta = sqrt(s1);
        scal[MR_OLDB] = oldb; scal[MR_BETA] = beta;
        scal[MR_ANORM2] = scal[MR_ANORM2] + alpha * alpha + oldb * oldb + beta * beta;
        const double gbar = sn * dbar - cs * alpha;
        scal[MR_EPS] = sn * beta;
        scal[MR_DBAR] = -cs * beta;
        scal[MR_ROOT] = sqrt(gbar * gbar + scal[MR_DBAR] * scal[MR_DBAR]);
        double gamma = sqrt(gbar * gbar + beta * beta);
        gamma = gamma > 2.220446049250313e-16 ? gamma : 2.220446049250313e-16;
        scal[MR_GAMMA] = gamma;
        scal[MR_CS] = gbar / gamma; scal[MR_SN] = beta / gamma;
        scal[MR_PHI] = scal[MR_CS] * scal[MR_PHIBAR];
        scal[MR_PHIBAR] = scal[MR_SN] * scal[MR_PHIBAR];
    } else if (mode == 7) { // stopping tests (s1 = ||x||^2)
        const double epsM = 2.220446049250313e-16, etol = sqrt(epsM), ctol = sqrt(epsM);
        const int iter = (int)scal[SC_ITER] + 1;
        scal[SC_ITER] = (double)iter;
        const double gamma = scal[MR_GAMMA], phi = scal[MR_PHI];
        scal[MR_ERRV + (iter % 5)] = phi;
        double err_lbnd = 0.0;
        if (iter >= 5) { double e2 = 0.0; for (int k = 0; k < 5; k++) e2 += scal[MR_ERRV + k] * scal[MR_ERRV + k]; err_lbnd = sqrt(e2); }
        scal[MR_GMAX] = scal[MR_GMAX] > gamma ? scal[MR_GMAX] : gamma;
        scal[MR_GMIN] = scal[MR_GMIN] < gamma ? scal[MR_GMIN] : gamma;
        const double ANorm = sqrt(scal[MR_ANORM2]), xNorm = sqrt(s1), Acond = scal[MR_GMAX] / scal[MR_GMIN], rNorm = scal[MR_PHIBAR];
        const double test1 = rNorm / (ANorm * xNorm), test2 = scal[MR_ROOT] / ANorm, tol = scal[SC_TOL];
        scal[MR_XENORM2] = scal[MR_XENORM2] + phi * phi;
        const bool ill = (1.0 + 1.0 / Acond <= 1.0) || (1.0 / Acond <= ctol);
        const bool solved = (1.0 + test2 <= 1.0) || (test2 <= tol) || (1.0 + test1 <= 1.0) || (test1 <= tol) ||
                            (iter >= 5 && err_lbnd <= etol * sqrt(scal[MR_XENORM2])) || (rNorm + 1.0 <= 1.0) || (rNorm <= tol);
        scal[SC_RR] = rNorm * rNorm;
        if (solved || ill || !(rNorm == rNorm)) scal[SC_CONV] = 1.0;
        if (solved && rNorm == rNorm) scal[SC_SOLVED] = 1.0; // Krylov's stats.solved: false for the ill-conditioned and NaN exits
    }
}
__global__ void __launch_bounds__(256) k_cg_scalars(int mode, int stage, int nblk, int nblk_max, const double *__restrict__ partial,
                                                    double *__restrict__ scal, double rtol, double atol, MailArgs mail) {
    cg_finish(mode, stage, nblk, nblk_max, partial, scal, rtol, atol, mail);
}
// what a producer kernel needs to finish its reduction in its last block
struct FuseArgs { int *ticket; int stage, nblk_max; double rtol, atol; MailArgs mail; };

// ---- K4: matvec  pressure.jl:119-130, with optional fused dot(x, y) partial ----------------------
// One thread per row, grid-stride.  Rows are in bucket (slot) order and the rows of 32 consecutive
// slots are contiguous in the edge arrays, so a warp's col / w reads fall into a handful of lines
// that stay in L1 over the ~6 iterations of the row loop and x[j] gathers hit L1/L2.  The per-row sum
// keeps the reference's order: diagonal first, then the neighbours in edge order.
// (A variant that staged each warp tile's col / w through shared memory measured 1.7x slower on
// B200 -- three dependent memory round trips per tile instead of one -- and was dropped.)
#define MV_U 8
template <bool DOT, int MINB, bool FUSE>
__global__ void __launch_bounds__(PR_BLOCK, MINB) k_matvec(int nslot, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                     const int *__restrict__ col, const double *__restrict__ w,
                                                     const double *__restrict__ diag, const double *__restrict__ x,
                                                     double *__restrict__ y, double *__restrict__ partial,
                                                     double *scal, FuseArgs fz) {
    __shared__ double sm[32];
    const bool idle = DOT && scal[SC_CONV] != 0.0; // converged: queued launches are no-ops (the fused finish still runs
    if (idle && !FUSE) return;                     // so that all ranks keep the same mailbox sequence numbers)
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (idle ? 0 : nslot); i += gridDim.x * blockDim.x) {
        const double xi = x[i];
        double yi = diag[i] * xi;
        const int r0 = rowptr[i], d = rdeg[i];
        // all loads of the first MV_U edges are issued before the first use (memory-level parallelism:
        // the kernel is latency bound, not bandwidth bound, when the loads of a row are chained)
        int cj[MV_U];
        double wk[MV_U], xj[MV_U];
#pragma unroll
        for (int k = 0; k < MV_U; k++) {
            cj[k] = k < d ? col[r0 + k] : -1;
            wk[k] = k < d ? w[r0 + k] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < MV_U; k++) xj[k] = cj[k] >= 0 ? x[cj[k]] : xi;
#pragma unroll
        for (int k = 0; k < MV_U; k++)
            if (k < d) yi += wk[k] * (xi - xj[k]);
        for (int k = MV_U; k < d; k++) {
            const int j = col[r0 + k];
            const double xv = j >= 0 ? x[j] : xi;
            yi += w[r0 + k] * (xi - xv);
        }
        y[i] = yi;
        if (DOT) acc += xi * yi;
    }
    if (DOT) {
        const double s = block_sum(acc, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
        if (FUSE && lv_last_block(fz.ticket)) { // mode 1: alpha = rr / p.Ap, finished here instead of in a one-block launch
            cg_finish(1, fz.stage, gridDim.x, fz.nblk_max, partial, scal, fz.rtol, fz.atol, fz.mail);
            if (threadIdx.x == 0) *fz.ticket = 0;
        }
    }
}

static inline int pr_grid(const LvContext *c, int64_t n) {
    int64_t nb = (n + PR_BLOCK - 1) / PR_BLOCK;
    int64_t cap = (int64_t)c->num_sms * 8;
    if (cap > 4096) cap = 4096;
    return (int)(nb < cap ? (nb < 1 ? 1 : nb) : cap);
}

// Grid of the matvec: exactly one resident wave.  At 48 registers only 5 blocks of 256 threads fit an SM, so the
// 8-per-SM grid above ran 1.6 waves and the second one left 40 % of the machine idle (the kernel is latency bound,
// so throughput follows the number of resident warps).  LV_MV_GRID=legacy restores the old sizing.
// MINB = 6 caps the kernel at 40 registers (no spills; 48 without the cap): 6 instead of 5 resident blocks per SM.
// LV_MV_MINB=1 selects the uncapped build, LV_MV_GRID=legacy the old grid (A/B switches for the bench).
static int mv_minb() {
    static int v = 0;
    if (v == 0) { const char *e = getenv("LV_MV_MINB"); v = (e && e[0] == '1') ? 1 : 6; }
    return v;
}
template <bool DOT> static int mv_grid(const LvContext *c, int64_t n) {
    static int per_sm = 0;
    if (per_sm == 0) {
        const char *e = getenv("LV_MV_GRID");
        int occ = 0;
        cudaError_t st = mv_minb() == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_matvec<DOT, 1, DOT>, PR_BLOCK, 0)
                                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_matvec<DOT, 6, DOT>, PR_BLOCK, 0);
        if ((e && !strcmp(e, "legacy")) || st != cudaSuccess || occ < 1) occ = 8;
        per_sm = occ;
    }
    int64_t nb = (n + PR_BLOCK - 1) / PR_BLOCK, cap = (int64_t)c->num_sms * per_sm;
    if (cap > 4096) cap = 4096;
    return (int)(nb < cap ? (nb < 1 ? 1 : nb) : cap);
}
template <bool DOT, bool FUSE = false>
static void mv_launch(LvContext *c, int grid, cudaStream_t st, int ns, const double *x, double *y, double *partial, double *scal,
                      const FuseArgs &fz = FuseArgs()) {
    if (mv_minb() == 1) k_matvec<DOT, 1, FUSE><<<grid, PR_BLOCK, 0, st>>>(ns, c->d_rowptr, c->d_deg, c->d_col, c->d_w, c->d_diag, x, y, partial, scal, fz);
    else k_matvec<DOT, 6, FUSE><<<grid, PR_BLOCK, 0, st>>>(ns, c->d_rowptr, c->d_deg, c->d_col, c->d_w, c->d_diag, x, y, partial, scal, fz);
    c->launches++;
}

int lv_pr_matvec(LvContext *c, const double *x, double *y) {
    if (!c->assembled) return lv_set_error(c, LV_EINVAL, "operator not assembled");
    LvProfScope prof(c, LV_PROF_MATVEC);
    const int ns = (int)c->nslot;
    if (ns == 0) return LV_OK;
    mv_launch<false>(c, mv_grid<false>(c, ns), c->stream, ns, x, y, nullptr, c->d_red);
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// ---- right-hand side  pressure.jl:162-203 ---------------------------------------------------------
struct Vbc { double w[8]; };

// First pass of find_pressure! (gp_step = false): b, the initial GP and -- kept for the later passes --
// bvel, the part of b that does not depend on P (velocity divergence + wall terms, pressure.jl:177,180-184).
// b itself is accumulated in the reference's order.
__global__ void __launch_bounds__(PR_BLOCK) k_rhs_first(LvGridParams g, int nslot, double dt, Vbc vbc, const unsigned char *__restrict__ own,
                                                        const double2 *__restrict__ ent_xy, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                        const int *__restrict__ col, const double2 *__restrict__ v1,
                                                        const double2 *__restrict__ v2, const double *__restrict__ lrr_in,
                                                        const double2 *__restrict__ mx_in, const double *__restrict__ area,
                                                        const double *__restrict__ mass, const double *__restrict__ rho,
                                                        const double *__restrict__ c2, const double *__restrict__ P,
                                                        const double2 *__restrict__ v, double *__restrict__ b, double *__restrict__ bvel,
                                                        double2 *__restrict__ GP, const unsigned *__restrict__ ent_label,
                                                        const int *__restrict__ bptr, const double2 *__restrict__ vbc_edge) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslot) return;
    if (!own[i]) { b[i] = 0.0; bvel[i] = 0.0; GP[i] = make_double2(0.0, 0.0); return; }
    int be = vbc_edge ? bptr[ent_label[i] & ~LV_IMAGE_BIT] : 0; // number of this polygon's first boundary edge
    const double2 x = ent_xy[i];
    const double Pi = P[i];
    const double2 vi = v[i];
    double bi = (area[i] * Pi) / ((rho[i] * c2[i]) * (dt * dt)); // pressure.jl:171
    double bv = 0.0;
    double gx = 0.0, gy = 0.0;
    const int r0 = rowptr[i], r1 = r0 + rdeg[i];
    for (int k = r0; k < r1; k++) { // neighbors(p, grid)
        const int j = col[k];
        if (j < 0) continue;
        const double2 y = lv_neighbor_pos(g, x, ent_xy[j]);
        const double2 m = mx_in[k]; // m - p.x
        const double lrr = lrr_in[k];
        const double2 a = v1[k], c = v2[k];
        const double mx = 0.5 * (a.x + c.x), my = 0.5 * (a.y + c.y);
        const double2 vj = v[j];
        const double t = (lrr / dt) * ((vi.x - vj.x) * (mx - y.x) + (vi.y - vj.y) * (my - y.y)); // :177
        bi -= t;
        bv -= t;
        const double s = lrr * (Pi - P[j]);
        gx -= s * m.x; // :178
        gy -= s * m.y;
    }
    for (int k = r0; k < r1; k++) { // boundaries(p)
        const int j = col[k];
        if (j >= 0) continue;
        const double2 a = v1[k], c = v2[k];
        const double sx = a.y - c.y, sy = c.x - a.x; // dS  :181
        const int wl = -j - 1;
        double bx = wl < 4 ? vbc.w[2 * wl] : 0.0, by = wl < 4 ? vbc.w[2 * wl + 1] : 0.0;
        if (vbc_edge) { const double2 ve = vbc_edge[be++]; bx = ve.x; by = ve.y; } // boundary_velocity(midpoint(e), e.label)  :182
        const double t = (sx * (bx - vi.x) + sy * (by - vi.y)) / dt; // :183
        bi -= t;
        bv -= t;
    }
    const double mi = mass[i];
    GP[i] = make_double2(gx / mi, gy / mi); // :185
    b[i] = bi;
    bvel[i] = bv;
}

#define RH_U 8 // edges whose loads are issued before first use (same latency argument as k_matvec)

// Later passes, sweep 1: GP_i = -sum lrr (P_i - P_j)(m - p.x) / mass_i   pressure.jl:178,185
__global__ void __launch_bounds__(PR_BLOCK) k_rhs_gp(int nslot, const unsigned char *__restrict__ own, const int *__restrict__ rowptr,
                                                     const unsigned char *__restrict__ rdeg, const int *__restrict__ col,
                                                     const double *__restrict__ lrr_in, const double2 *__restrict__ mx_in,
                                                     const double *__restrict__ mass, const double *__restrict__ P,
                                                     double2 *__restrict__ GP) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslot) return;
    if (!own[i]) { GP[i] = make_double2(0.0, 0.0); return; }
    const double Pi = P[i];
    const int r0 = rowptr[i], d = rdeg[i];
    int cj[RH_U];
    double lr[RH_U], pj[RH_U];
    double2 mm[RH_U];
#pragma unroll
    for (int k = 0; k < RH_U; k++) {
        cj[k] = k < d ? col[r0 + k] : -1;
        lr[k] = k < d ? lrr_in[r0 + k] : 0.0;
        mm[k] = k < d ? mx_in[r0 + k] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < RH_U; k++) pj[k] = cj[k] >= 0 ? P[cj[k]] : Pi;
    double gx = 0.0, gy = 0.0;
#pragma unroll
    for (int k = 0; k < RH_U; k++)
        if (k < d && cj[k] >= 0) {
            const double s = lr[k] * (Pi - pj[k]);
            gx -= s * mm[k].x;
            gy -= s * mm[k].y;
        }
    for (int k = RH_U; k < d; k++) {
        const int j = col[r0 + k];
        if (j < 0) continue;
        const double s = lrr_in[r0 + k] * (Pi - P[j]);
        const double2 m = mx_in[r0 + k];
        gx -= s * m.x;
        gy -= s * m.y;
    }
    const double mi = mass[i];
    GP[i] = make_double2(gx / mi, gy / mi);
}

// Later passes, sweep 1 fused with the first matvec of the solve: both gather P_j over the same row, so
// GP_i (pressure.jl:178,185) and (A P)_i (pressure.jl:119-130, the A x0 of the warm-started Krylov solve) come out of
// one CSR walk.  (A P)_i is accumulated exactly like k_matvec does (diagonal first, neighbours in edge order).
__global__ void __launch_bounds__(PR_BLOCK) k_rhs_gp_mv(int nslot, const unsigned char *__restrict__ own, const int *__restrict__ rowptr,
                                                        const unsigned char *__restrict__ rdeg, const int *__restrict__ col,
                                                        const double *__restrict__ lrr_in, const double2 *__restrict__ mx_in,
                                                        const double *__restrict__ w, const double *__restrict__ diag,
                                                        const double *__restrict__ mass, const double *__restrict__ P,
                                                        double2 *__restrict__ GP, double *__restrict__ AP) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslot) return;
    const double Pi = P[i];
    if (!own[i]) { GP[i] = make_double2(0.0, 0.0); AP[i] = diag[i] * Pi; return; }
    const int r0 = rowptr[i], d = rdeg[i];
    int cj[RH_U];
    double lr[RH_U], pj[RH_U], wk[RH_U];
    double2 mm[RH_U];
#pragma unroll
    for (int k = 0; k < RH_U; k++) {
        cj[k] = k < d ? col[r0 + k] : -1;
        lr[k] = k < d ? lrr_in[r0 + k] : 0.0;
        wk[k] = k < d ? w[r0 + k] : 0.0;
        mm[k] = k < d ? mx_in[r0 + k] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < RH_U; k++) pj[k] = cj[k] >= 0 ? P[cj[k]] : Pi;
    double gx = 0.0, gy = 0.0, yi = diag[i] * Pi;
#pragma unroll
    for (int k = 0; k < RH_U; k++) {
        if (k < d) yi += wk[k] * (Pi - pj[k]);
        if (k < d && cj[k] >= 0) {
            const double s = lr[k] * (Pi - pj[k]);
            gx -= s * mm[k].x;
            gy -= s * mm[k].y;
        }
    }
    for (int k = RH_U; k < d; k++) {
        const int j = col[r0 + k];
        const double pv = j >= 0 ? P[j] : Pi;
        yi += w[r0 + k] * (Pi - pv);
        if (j < 0) continue;
        const double s = lrr_in[r0 + k] * (Pi - pv);
        const double2 m = mx_in[r0 + k];
        gx -= s * m.x;
        gy -= s * m.y;
    }
    const double mi = mass[i];
    GP[i] = make_double2(gx / mi, gy / mi);
    AP[i] = yi;
}

// Later passes, sweep 2: b_i = A_i P_i/(rho c2 dt^2) + bvel_i + sum lrr (GP_i - GP_j).(m - z)   pressure.jl:171,189-202
// INIT: the kernel is also the CG initialisation (k_cg_init): r = b - A x0 with A x0 from k_rhs_gp_mv, p = r and the block
// partials of r.r and b.b -- grid-stride with k_cg_init's grid and accumulation order, so the sums are bit-identical to it.
template <bool INIT>
__global__ void __launch_bounds__(PR_BLOCK, INIT ? 4 : 8) k_rhs_corr(int nslot, double dt, const unsigned char *__restrict__ own, const int *__restrict__ rowptr,
                                                       const unsigned char *__restrict__ rdeg, const int *__restrict__ col,
                                                       const double *__restrict__ lrr_in, const double2 *__restrict__ mz_in,
                                                       const double *__restrict__ area, const double *__restrict__ rho,
                                                       const double *__restrict__ c2, const double *__restrict__ P,
                                                       const double *__restrict__ bvel, const double2 *__restrict__ GP,
                                                       double *__restrict__ b, const double *__restrict__ AP, double *__restrict__ r,
                                                       double *__restrict__ p, double *__restrict__ partial, int nblk_max,
                                                       const float *__restrict__ dinv) {
    __shared__ double sm[32];
    double rr = 0.0, bb = 0.0;
    const int stride = INIT ? gridDim.x * blockDim.x : nslot;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += stride) {
        double bi = 0.0;
        if (own[i]) {
            const double2 gi = GP[i];
            bi = (area[i] * P[i]) / ((rho[i] * c2[i]) * (dt * dt)) + bvel[i];
            const int r0 = rowptr[i], d = rdeg[i];
            int cj[RH_U];
            double lr[RH_U];
            double2 mm[RH_U], gj[RH_U];
#pragma unroll
            for (int k = 0; k < RH_U; k++) {
                cj[k] = k < d ? col[r0 + k] : -1;
                lr[k] = k < d ? lrr_in[r0 + k] : 0.0;
                mm[k] = k < d ? mz_in[r0 + k] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int k = 0; k < RH_U; k++) gj[k] = cj[k] >= 0 ? GP[cj[k]] : gi;
#pragma unroll
            for (int k = 0; k < RH_U; k++)
                if (k < d && cj[k] >= 0) bi += lr[k] * ((gi.x - gj[k].x) * mm[k].x + (gi.y - gj[k].y) * mm[k].y); // :198
            for (int k = RH_U; k < d; k++) {
                const int j = col[r0 + k];
                if (j < 0) continue;
                const double2 g2 = GP[j], m = mz_in[r0 + k];
                bi += lrr_in[r0 + k] * ((gi.x - g2.x) * m.x + (gi.y - g2.y) * m.y);
            }
        }
        b[i] = bi;
        if (INIT) {
            const double ri = bi - AP[i];
            r[i] = ri;
            rr += ri * ri;
            if (dinv) { const double zi = (double)dinv[i] * ri; p[i] = zi; bb += ri * zi; } // Jacobi: p = z, second sum = r.z
            else { p[i] = ri; bb += bi * bi; }
        }
    }
    if (INIT) {
        const double s1 = block_sum(rr, sm);
        const double s2 = block_sum(bb, sm);
        if (threadIdx.x == 0) { partial[blockIdx.x] = s1; partial[nblk_max + blockIdx.x] = s2; }
    }
}

static inline int pr_grid(const LvContext *c, int64_t n);
int lv_pr_rhs(LvContext *c, double dt, int gp_step, const double *vbc_wall, bool fuse_init, bool *init_done, bool pcg) {
    if (init_done) *init_done = false;
    LV_TRY(lv_pr_ensure(c));
    if (!c->pr_valid) return lv_set_error(c, LV_EINVAL, "fields not uploaded: call lv_fields_upload first");
    if (!c->assembled || c->asm_dt != dt) LV_TRY(lv_pr_assemble(c, dt)); // the per-edge factors come from the assembly
    LvProfScope prof(c, LV_PROF_ASSEMBLE);
    const int ns = (int)c->nslot;
    if (ns == 0) return LV_OK;
    Vbc vbc;
    bool same_vbc = true;
    for (int k = 0; k < 8; k++) { vbc.w[k] = vbc_wall ? vbc_wall[k] : 0.0; same_vbc &= (vbc.w[k] == c->last_vbc[k]); }
    const int nb = (ns + PR_BLOCK - 1) / PR_BLOCK;
    const bool first = !gp_step || !c->bvel_valid || !same_vbc;
    LV_TRY(lv_halo_exchange(c, c->d_P, 1)); // neighbours' pressures across the strip edges (no-op on one GPU)
    if (first) {
        k_rhs_first<<<nb, PR_BLOCK, 0, c->stream>>>(c->gp, ns, dt, vbc, c->d_own, c->d_ent_xy, c->d_rowptr, c->d_deg, c->d_col, c->d_v1,
                                                    c->d_v2, c->d_lrr, c->d_mx, c->d_area, c->d_mass, c->d_rho, c->d_c2, c->d_P, c->d_v, c->d_b,
                                                    c->d_bvel, c->d_GP, c->d_ent_label, c->d_bdry_ptr, c->vbc_edge_on ? c->d_vbc_edge : nullptr);
        c->launches++;
        for (int k = 0; k < 8; k++) c->last_vbc[k] = vbc.w[k];
        c->bvel_valid = true;
    }
    if (gp_step) {
        // fused with the start of the CG solve (its first matvec and k_cg_init) when the caller is find_pressure!
        const bool fused = fuse_init && !first;
        if (!first) {
            if (fused) k_rhs_gp_mv<<<nb, PR_BLOCK, 0, c->stream>>>(ns, c->d_own, c->d_rowptr, c->d_deg, c->d_col, c->d_lrr, c->d_mx, c->d_w, c->d_diag,
                                                                   c->d_mass, c->d_P, c->d_GP, c->d_vec[2]);
            else k_rhs_gp<<<nb, PR_BLOCK, 0, c->stream>>>(ns, c->d_own, c->d_rowptr, c->d_deg, c->d_col, c->d_lrr, c->d_mx, c->d_mass, c->d_P,
                                                          c->d_GP);
            c->launches++;
        }
        LV_TRY(lv_halo_exchange(c, (double *)c->d_GP, 2));
        if (fused) {
            k_rhs_corr<true><<<pr_grid(c, ns), PR_BLOCK, 0, c->stream>>>(ns, dt, c->d_own, c->d_rowptr, c->d_deg, c->d_col, c->d_lrr, c->d_mz, c->d_area,
                                                                         c->d_rho, c->d_c2, c->d_P, c->d_bvel, c->d_GP, c->d_b, c->d_vec[2],
                                                                         c->d_vec[0], c->d_vec[1], c->d_red + SC_COUNT, 4096, pcg ? (const float *)c->d_dinv : nullptr);
            if (init_done) *init_done = true;
        } else
            k_rhs_corr<false><<<nb, PR_BLOCK, 0, c->stream>>>(ns, dt, c->d_own, c->d_rowptr, c->d_deg, c->d_col, c->d_lrr, c->d_mz, c->d_area, c->d_rho,
                                                              c->d_c2, c->d_P, c->d_bvel, c->d_GP, c->d_b, nullptr, nullptr, nullptr, nullptr, 4096, nullptr);
        c->launches++;
    }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// ---- K5: conjugate gradients, scalars resident on the device ------------------------------------------
// r = b - A x ; p = r ; partial(r.r), partial(b.b)
// dinv (nullable): Jacobi preconditioner 1/A_ii -- then p = z = dinv r and the second partial is r.z instead of b.b
__global__ void __launch_bounds__(PR_BLOCK) k_cg_init(int nslot, const double *__restrict__ b, const double *__restrict__ Ax,
                                                      double *__restrict__ r, double *__restrict__ p, double *__restrict__ partial,
                                                      int nblk_max, const float *__restrict__ dinv) {
    __shared__ double sm[32];
    double rr = 0.0, bb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        const double bi = b[i];
        const double ri = bi - Ax[i];
        r[i] = ri;
        rr += ri * ri;
        if (dinv) { const double zi = (double)dinv[i] * ri; p[i] = zi; bb += ri * zi; }
        else { p[i] = ri; bb += bi * bi; }
    }
    const double s1 = block_sum(rr, sm);
    const double s2 = block_sum(bb, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s1; partial[nblk_max + blockIdx.x] = s2; }
}

// What the host looks at between batches (CONV, ITER, RES2, BNORM2) goes to mapped pinned memory with one tiny kernel
// per batch -- not a device->host memcpy, which would queue behind a background copy of the edge view, and not a
// store from every scalar kernel, whose system-scope fence would stall behind that copy's PCIe traffic.
__global__ void k_publish_scalars(const double *__restrict__ scal, double *host, const int *dead) {
    if (threadIdx.x == 0) {
        host[SC_DEAD] = (double)*dead;
        host[SC_CONV] = scal[SC_CONV]; host[SC_ITER] = scal[SC_ITER]; host[SC_SOLVED] = scal[SC_SOLVED];
        host[SC_RES2] = scal[SC_RES2]; host[SC_BNORM2] = scal[SC_BNORM2];
        __threadfence_system();
    }
}

// ---- MINRES vector kernels (vectors: r1, r2 (= v, M = I), y, w1, w2, x; see lv_pr_solve_minres) ----------------
__global__ void __launch_bounds__(PR_BLOCK) k_mr_init(int nslot, const double *__restrict__ b, const double *__restrict__ Ax0, double *__restrict__ r1,
                                                      double *__restrict__ r2, double *__restrict__ w1, double *__restrict__ w2,
                                                      double *__restrict__ x, double *__restrict__ partial) {
    __shared__ double sm[32];
    double rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        const double ri = b[i] - Ax0[i]; // warm start: r1 = b - A*x0
        r1[i] = ri; r2[i] = ri; w1[i] = 0.0; w2[i] = 0.0; x[i] = 0.0;
        rr += ri * ri;
    }
    const double s = block_sum(rr, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// y = y/beta - (beta/oldbeta) r1 (iter >= 2); partial(v.y)
__global__ void __launch_bounds__(PR_BLOCK) k_mr_a(int nslot, int iter, const double *__restrict__ scal, const double *__restrict__ v,
                                                   const double *__restrict__ r1, double *__restrict__ y, double *__restrict__ partial) {
    __shared__ double sm[32];
    if (scal[SC_CONV] != 0.0) return;
    const double beta = scal[MR_BETA], oldb = scal[MR_OLDB];
    const double ib = 1.0 / beta, c1 = iter >= 2 ? -beta / oldb : 0.0;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        double yi = y[i] * ib;
        if (iter >= 2) yi += c1 * r1[i];
        y[i] = yi;
        acc += v[i] * yi;
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// y -= (alpha/beta) r2; w update; r1 = r2; r2 = y; partial(r2.r2)
// `own` (nullable): rows this rank owns.  Ghost slots of v = r2 hold the neighbours' values after the halo exchange; they
// must not enter the recurrences or the norms, so they are read as zero (and left zero for the next exchange to refill).
__global__ void __launch_bounds__(PR_BLOCK) k_mr_b(int nslot, int iter, const double *__restrict__ scal, double *__restrict__ r1,
                                                   double *__restrict__ r2, const double *__restrict__ y, double *__restrict__ wa /* w1 */,
                                                   double *__restrict__ wb /* w2 */, double *__restrict__ partial,
                                                   const unsigned char *__restrict__ own) {
    __shared__ double sm[32];
    if (scal[SC_CONV] != 0.0) return;
    const double beta = scal[MR_BETA], alpha = scal[SC_ALPHA], eps_old = scal[MR_EPS], delta = scal[MR_DELTA];
    const double c2 = -alpha / beta, ib = 1.0 / beta;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        const double vi = (own == nullptr || own[i]) ? r2[i] : 0.0;
        const double yi = y[i] + c2 * vi;
        if (iter == 1) wb[i] = wb[i] + ib * vi;
        else {
            double w = wa[i];
            if (iter >= 3) w = w * (-eps_old);
            w = w + (-delta) * wb[i];
            wa[i] = w + ib * vi;
        }
        r1[i] = vi;
        r2[i] = yi;
        acc += yi * yi;
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// w /= gamma; x += phi w; partial(x.x)
__global__ void __launch_bounds__(PR_BLOCK) k_mr_c(int nslot, const double *__restrict__ scal, double *__restrict__ w, double *__restrict__ x,
                                                   double *__restrict__ partial) {
    __shared__ double sm[32];
    if (scal[SC_CONV] != 0.0) return;
    const double ig = 1.0 / scal[MR_GAMMA], phi = scal[MR_PHI];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        const double wi = w[i] * ig;
        w[i] = wi;
        const double xi = x[i] + phi * wi;
        x[i] = xi;
        acc += xi * xi;
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(PR_BLOCK) k_axpy1(int nslot, const double *__restrict__ x, double *__restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) y[i] += x[i];
}

// r -= alpha Ap ; partial(r.r).  The x update rides along with the p update below, which reads p anyway:
// 8 instead of 9 vector streams per iteration, same arithmetic per element.
// PC: Jacobi-preconditioned CG -- the kernel also accumulates r.z = sum dinv_i r_i^2 (second partial) and finishes with mode 9
template <bool FUSE, bool PC>
__global__ void __launch_bounds__(PR_BLOCK) k_cg_update_r(int nslot, double *scal, const double *__restrict__ Ap,
                                                          double *__restrict__ r, double *__restrict__ partial, FuseArgs fz,
                                                          const float *__restrict__ dinv, int nblk_max) {
    __shared__ double sm[32];
    const bool idle = scal[SC_CONV] != 0.0;
    if (idle && !FUSE) return;
    const double alpha = scal[SC_ALPHA];
    double rr = 0.0, rz = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (idle ? 0 : nslot); i += gridDim.x * blockDim.x) {
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        rr += ri * ri;
        if (PC) rz += ri * ((double)dinv[i] * ri);
    }
    const double s = block_sum(rr, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
    if (PC) {
        const double s2 = block_sum(rz, sm);
        if (threadIdx.x == 0) partial[nblk_max + blockIdx.x] = s2;
    }
    if (FUSE && lv_last_block(fz.ticket)) { // mode 2 (9 with the preconditioner): beta, iteration count, convergence test
        cg_finish(PC ? 9 : 2, fz.stage, gridDim.x, fz.nblk_max, partial, scal, fz.rtol, fz.atol, fz.mail);
        if (threadIdx.x == 0) *fz.ticket = 0;
    }
}

// x += alpha p ; p = r + beta p.  `iter` is this iteration's 1-based index: the iteration that converged still
// applies its x update (and skips the p update, which nobody reads); later queued iterations are no-ops.
// PACK (strip decomposition over peer memory): the kernel also stores the new p of every slot a neighbour needs into
// this rank's halo outbox (only slots outside [bounds[0], bounds[1]) can be such slots, so the lookup costs nothing in
// the interior) and its last block publishes the exchange's sequence word -- no pack / signal launches.
template <bool PACK>
__global__ void __launch_bounds__(PR_BLOCK) k_cg_update_xp(int nslot, int iter, const double *__restrict__ scal, const double *__restrict__ r,
                                                           double *__restrict__ x, double *__restrict__ p, LvHaloPack hp,
                                                           const float *__restrict__ dinv) {
    const bool conv = scal[SC_CONV] != 0.0;
    const bool idle = conv && scal[SC_ITER] != (double)iter;
    if (idle && !PACK) return;
    const double alpha = scal[SC_ALPHA], beta = scal[SC_BETA];
    bool wrote = false;
    if (idle) {
    } else if (conv) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) x[i] += alpha * p[i];
    } else {
        const int b0 = PACK ? hp.bounds[0] : 0, b1 = PACK ? hp.bounds[1] : 0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
            const double pi = p[i];
            x[i] += alpha * pi;
            const double pn = (dinv ? (double)dinv[i] * r[i] : r[i]) + beta * pi; // p = z + beta p, z = D^-1 r with the Jacobi preconditioner
            p[i] = pn;
            if (PACK && (i < b0 || i >= b1)) {
                const int s0 = hp.send_pos0[i];
                if (s0 >= 0) {
                    hp.outbox[s0] = pn;
                    const int s1 = hp.send_pos1[i];
                    if (s1 >= 0) hp.outbox[s1] = pn;
                    wrote = true;
                }
            }
        }
    }
    if (PACK) {
        if (wrote) __threadfence_system(); // the outbox is read by other GPUs
        if (lv_last_block(hp.ticket) && threadIdx.x == 0) {
            __threadfence_system();
            lv_st_release_sys(hp.flag, hp.seq);
            *hp.ticket = 0;
        }
    }
}

// residual check: partial(|b - Ax|^2), partial(|b|^2)
__global__ void __launch_bounds__(PR_BLOCK) k_resid(int nslot, const double *__restrict__ b, const double *__restrict__ Ax,
                                                    double *__restrict__ partial, int nblk_max) {
    __shared__ double sm[32];
    double rr = 0.0, bb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        const double bi = b[i], d = bi - Ax[i];
        rr += d * d;
        bb += bi * bi;
    }
    const double s1 = block_sum(rr, sm);
    const double s2 = block_sum(bb, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s1; partial[nblk_max + blockIdx.x] = s2; }
}

// A x = b with x = c->d_P (initial guess in, solution out), b = c->d_b
int lv_pr_solve(LvContext *c, int solver, double rtol, double atol, int itmax, int *iters, double *relres, bool pre_init) {
    if (!c->assembled) return lv_set_error(c, LV_EINVAL, "operator not assembled");
    if (solver != LV_SOLVER_CG && solver != LV_SOLVER_MINRES && solver != LV_SOLVER_PCG) return lv_set_error(c, LV_EINVAL, "unknown solver %d", solver);
    const bool minres = solver == LV_SOLVER_MINRES, pcg = solver == LV_SOLVER_PCG;
    const float *dinv = pcg ? (const float *)c->d_dinv : nullptr; // Jacobi preconditioner 1/A_ii (k_assemble), single precision
    const int ns = (int)c->nslot;
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    if (ns == 0) return LV_OK;
    cudaStream_t st = c->stream;
    double *x = c->d_P, *b = c->d_b, *r = c->d_vec[0], *p = c->d_vec[1], *Ap = c->d_vec[2];
    double *scal = c->d_red, *partial = c->d_red + SC_COUNT;
    const int NBMAX = 4096;
    const int nb = pr_grid(c, ns), nb_mv = mv_grid<true>(c, ns), nb_mv0 = mv_grid<false>(c, ns);
    // How a reduction is finished.  Single GPU, or several GPUs with mapped mailboxes: inside the producer kernel's last
    // block (CG modes 1 and 2) or by a one-block launch (everything else); the mailbox variant exchanges the two local sums
    // with every rank over NVLink and adds them in rank order.  Without mailboxes: reduce / ncclAllReduce / update.
    const bool multi = c->comm != nullptr;
    const bool fuse = !multi || c->mailbox_ready;
    const bool peer = multi && lv_strip_peer_mode(c); // halo values through the exchange areas (NVLink pulls)
    auto mail_next = [&]() -> MailArgs {
        if (multi && c->mailbox_ready) return MailArgs{(LvMailSlot *const *)c->d_mailbox_ptrs, c->nranks, c->rank, ++c->ar_seq, c->d_tickets + 7};
        return MailArgs{nullptr, 1, 0, 0, c->d_tickets + 7};
    };
    auto fuse_args = [&](int ticket) -> FuseArgs { return FuseArgs{c->d_tickets + ticket, (multi && c->mailbox_ready) ? 3 : 0, NBMAX, rtol, atol, mail_next()}; };
    auto finish = [&](int mode) -> int {
        const int nb = mode == 1 ? nb_mv : pr_grid(c, ns);
        if (fuse) {
            k_cg_scalars<<<1, 256, 0, st>>>(mode, (multi && c->mailbox_ready) ? 3 : 0, nb, NBMAX, partial, scal, rtol, atol, mail_next());
            c->launches++;
            return LV_OK;
        }
        const MailArgs mail{nullptr, 1, 0, 0, c->d_tickets + 7};
        k_cg_scalars<<<1, 256, 0, st>>>(mode, 1, nb, NBMAX, partial, scal, rtol, atol, mail);
        LV_TRY(lv_allreduce_sum(c, scal + SC_TMP0, 2));
        k_cg_scalars<<<1, 256, 0, st>>>(mode, 2, nb, NBMAX, partial, scal, rtol, atol, mail);
        c->launches += 2;
        return LV_OK;
    };
    auto matvec_plain = [&](const double *in, double *out) {
        LvProfScope prof(c, LV_PROF_MATVEC);
        mv_launch<false>(c, nb_mv0, st, ns, in, out, nullptr, scal);
    };
    auto publish = [&]() -> int {
        if (c->flags_mapped) k_publish_scalars<<<1, 32, 0, st>>>(scal, c->h_red, c->d_tickets + 7);
        else LV_CUDA(c, cudaMemcpyAsync(c->h_red, scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, st));
        LV_CUDA(c, cudaStreamSynchronize(st));
        if (c->flags_mapped && c->h_red[SC_DEAD] != 0.0) return lv_set_error(c, LV_ECUDA, "a peer GPU did not arrive at an exchange (timeout)");
        return LV_OK;
    };
    if (pre_init && minres) return lv_set_error(c, LV_EINVAL, "pre-initialised solves are CG only");
    if (pcg && !c->d_dinv) return lv_set_error(c, LV_EINVAL, "operator not assembled");
    if (!pre_init) {
        LV_TRY(lv_halo_exchange(c, x, 1)); // ghost columns of the initial guess
        matvec_plain(x, Ap);
    }
    if (!minres) {
        {
            LvProfScope prof(c, LV_PROF_VECOPS);
            if (!pre_init) { // otherwise the right-hand-side kernels left r, p and the partial sums behind (lv_pr_rhs)
                k_cg_init<<<nb, PR_BLOCK, 0, st>>>(ns, b, Ap, r, p, partial, NBMAX, dinv);
                c->launches++;
            }
            LV_TRY(finish(pcg ? 8 : 0));
            if (peer) LV_TRY(lv_strip_halo_post(c, p, 1)); // p = r is ready for the neighbours
        }
        // Iterations are queued in batches; kernels turn into no-ops once the device-side convergence
        // flag is set, so the host only has to look at the flag between batches.  The first batch is
        // sized by the previous solve (the fixed-point passes of find_pressure! need similar counts).
        // Launches per iteration: [halo pull] matvec (+ alpha) / r update (+ beta, convergence) / x, p update (+ halo pack and
        // signal) -- three on one GPU, four on several.
        int done = 0;
        while (done < itmax) {
            int batch = done == 0 ? (c->cg_hint > 8 ? c->cg_hint : 8) : 8;
            const int todo = itmax - done < batch ? itmax - done : batch;
            for (int it = 0; it < todo; it++) {
                if (peer) LV_TRY(lv_strip_halo_pull(c, p)); // ghost columns of the search direction, out of the neighbours' outboxes
                else if (multi) LV_TRY(lv_halo_exchange(c, p, 1));
                {
                    LvProfScope prof(c, LV_PROF_MATVEC);
                    if (fuse) mv_launch<true, true>(c, nb_mv, st, ns, p, Ap, partial, scal, fuse_args(0));
                    else mv_launch<true, false>(c, nb_mv, st, ns, p, Ap, partial, scal);
                }
                LvProfScope prof(c, LV_PROF_VECOPS);
                if (fuse) {
                    if (pcg) k_cg_update_r<true, true><<<nb, PR_BLOCK, 0, st>>>(ns, scal, Ap, r, partial, fuse_args(1), dinv, NBMAX);
                    else k_cg_update_r<true, false><<<nb, PR_BLOCK, 0, st>>>(ns, scal, Ap, r, partial, fuse_args(1), nullptr, NBMAX);
                } else {
                    LV_TRY(finish(1));
                    if (pcg) k_cg_update_r<false, true><<<nb, PR_BLOCK, 0, st>>>(ns, scal, Ap, r, partial, FuseArgs(), dinv, NBMAX);
                    else k_cg_update_r<false, false><<<nb, PR_BLOCK, 0, st>>>(ns, scal, Ap, r, partial, FuseArgs(), nullptr, NBMAX);
                    LV_TRY(finish(pcg ? 9 : 2));
                }
                if (peer) {
                    LvHaloPack hp;
                    LV_TRY(lv_strip_pack_args(c, &hp));
                    k_cg_update_xp<true><<<nb, PR_BLOCK, 0, st>>>(ns, done + it + 1, scal, r, x, p, hp, dinv);
                } else k_cg_update_xp<false><<<nb, PR_BLOCK, 0, st>>>(ns, done + it + 1, scal, r, x, p, LvHaloPack(), dinv);
                c->launches += 2;
            }
            done += todo;
            LV_TRY(publish());
            if (c->h_red[SC_CONV] != 0.0) break;
        }
    } else {
        // MINRES (the reference's Krylov method, pressure.jl:219), warm-started: solve A dx = b - A x0, x = x0 + dx.
        double *r1 = c->d_vec[0], *r2 = c->d_vec[1], *y = c->d_vec[2], *w1 = c->d_vec[3], *w2 = c->d_vec[4], *dx = c->d_vec[5];
        {
            LvProfScope prof(c, LV_PROF_VECOPS);
            k_mr_init<<<nb, PR_BLOCK, 0, st>>>(ns, b, Ap /* = A x0, same buffer as y */, r1, r2, w1, w2, dx, partial);
            c->launches++;
            LV_TRY(finish(4));
        }
        int done = 0;
        while (done < itmax) {
            int batch = done == 0 ? (c->cg_hint > 8 ? c->cg_hint : 8) : 8;
            const int todo = itmax - done < batch ? itmax - done : batch;
            for (int it = 0; it < todo; it++) {
                const int iter = done + it + 1;
                LV_TRY(lv_halo_exchange(c, r2, 1));
                {
                    LvProfScope prof(c, LV_PROF_MATVEC);
                    mv_launch<true, false>(c, nb_mv, st, ns, r2, y, partial, scal);
                }
                LvProfScope prof(c, LV_PROF_VECOPS);
                k_mr_a<<<nb, PR_BLOCK, 0, st>>>(ns, iter, scal, r2, r1, y, partial);
                LV_TRY(finish(5));
                double *wa = (iter == 1) ? w1 : w1, *wb = w2; // iter == 1 works on w2 in place, later iterations build w in w1
                k_mr_b<<<nb, PR_BLOCK, 0, st>>>(ns, iter, scal, r1, r2, y, wa, wb, partial, multi ? c->d_own : nullptr);
                LV_TRY(finish(6));
                k_mr_c<<<nb, PR_BLOCK, 0, st>>>(ns, scal, iter == 1 ? w2 : w1, dx, partial);
                LV_TRY(finish(7));
                c->launches += 3;
                if (iter >= 2) { double *t = w1; w1 = w2; w2 = t; }
            }
            done += todo;
            LV_TRY(publish());
            if (c->h_red[SC_CONV] != 0.0) break;
        }
        k_axpy1<<<nb, PR_BLOCK, 0, st>>>(ns, dx, x); // x = x0 + dx
        c->launches++;
    }
    LV_CUDA(c, cudaGetLastError());
    c->cg_hint = (int)c->h_red[SC_ITER];
    if (iters) *iters = (int)c->h_red[SC_ITER];
    if (relres) { // true residual ||b - A x|| / ||b||
        LV_TRY(lv_halo_exchange(c, x, 1));
        matvec_plain(x, Ap);
        k_resid<<<nb, PR_BLOCK, 0, st>>>(ns, b, Ap, partial, NBMAX);
        c->launches++;
        LV_TRY(finish(3));
        LV_TRY(publish());
        const double bn = c->h_red[SC_BNORM2], rn = c->h_red[SC_RES2];
        *relres = bn > 0.0 ? sqrt(rn / bn) : sqrt(rn);
    }
    return LV_OK;
}

// Cold-started MINRES on an arbitrary symmetric operator given as a callback (single GPU): used by the multiphase
// projector (relaxation.jl:182), which is matrix free and lives in label order.  Same iteration and stopping rules as
// the pressure MINRES above (Krylov.jl minres!).  x receives the solution; *solved = 0 when itmax was hit.
int lv_minres_apply(LvContext *c, int n, const std::function<int(const double *, double *)> &apply, const double *b, double *x,
                    double rtol, double atol, int itmax, int *iters, int *solved) {
    if (iters) *iters = 0;
    if (solved) *solved = 1;
    const bool multi = c->comm != nullptr;
    if (n == 0 && !multi) return LV_OK;
    if (multi && !c->mailbox_ready) return lv_set_error(c, LV_EINVAL, "lv_minres_apply on several GPUs needs the peer mailboxes");
    cudaStream_t st = c->stream;
    double *scal = c->d_red, *partial = c->d_red + SC_COUNT;
    const int NBMAX = 4096;
    const int nb = pr_grid(c, n);
    // reductions over the rows of this rank; on several GPUs the two sums of every step are exchanged through the mailboxes
    auto scalars = [&](int mode) {
        if (multi) k_cg_scalars<<<1, 256, 0, st>>>(mode, 3, nb, NBMAX, partial, scal, rtol, atol, MailArgs{(LvMailSlot *const *)c->d_mailbox_ptrs, c->nranks, c->rank, ++c->ar_seq, c->d_tickets + 7});
        else k_cg_scalars<<<1, 256, 0, st>>>(mode, 0, nb, NBMAX, partial, scal, rtol, atol, MailArgs{nullptr, 1, 0, 0, c->d_tickets + 7});
    };
    double *r1 = c->d_vec[0], *r2 = c->d_vec[1], *y = c->d_vec[2], *w1 = c->d_vec[3], *w2 = c->d_vec[4];
    LV_CUDA(c, cudaMemsetAsync(y, 0, sizeof(double) * (size_t)n, st)); // A*x0 with x0 = 0
    k_mr_init<<<nb, PR_BLOCK, 0, st>>>(n, b, y, r1, r2, w1, w2, x, partial);
    scalars(4);
    c->launches += 2;
    int done = 0;
    bool conv = false;
    while (done < itmax) {
        const int todo = itmax - done < 16 ? itmax - done : 16;
        for (int it = 0; it < todo; it++) {
            const int iter = done + it + 1;
            LV_TRY(apply(r2, y));
            k_mr_a<<<nb, PR_BLOCK, 0, st>>>(n, iter, scal, r2, r1, y, partial);
            scalars(5);
            k_mr_b<<<nb, PR_BLOCK, 0, st>>>(n, iter, scal, r1, r2, y, w1, w2, partial, nullptr);
            scalars(6);
            k_mr_c<<<nb, PR_BLOCK, 0, st>>>(n, scal, iter == 1 ? w2 : w1, x, partial);
            scalars(7);
            c->launches += 6;
            if (iter >= 2) { double *t = w1; w1 = w2; w2 = t; }
        }
        done += todo;
        if (c->flags_mapped) k_publish_scalars<<<1, 32, 0, st>>>(scal, c->h_red, c->d_tickets + 7);
        else LV_CUDA(c, cudaMemcpyAsync(c->h_red, scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, st));
        LV_CUDA(c, cudaStreamSynchronize(st));
        if (c->h_red[SC_CONV] != 0.0) { conv = true; break; }
    }
    LV_CUDA(c, cudaGetLastError());
    if (iters) *iters = (int)c->h_red[SC_ITER];
    if (solved) *solved = (conv && c->h_red[SC_SOLVED] != 0.0) ? 1 : 0; // relaxation.jl:184 warns when stats.solved is false
    return LV_OK;
}

// find_pressure!  pressure.jl:215-225
int lv_pr_find_pressure(LvContext *c, double dt, int niter, double rtol, double atol, int itmax, int solver, const double *vbc_wall,
                        int32_t *iters_out, double *relres_out) {
    LV_TRY(lv_pr_assemble(c, dt));
    for (int it = 1; it <= niter; it++) {
        bool init_done = false;
        LV_TRY(lv_pr_rhs(c, dt, it > 1, vbc_wall, solver == LV_SOLVER_CG || solver == LV_SOLVER_PCG, &init_done, solver == LV_SOLVER_PCG));
        int iters = 0;
        double relres = 0.0;
        LV_TRY(lv_pr_solve(c, solver, rtol, atol, itmax, &iters, relres_out ? &relres : nullptr, init_done));
        if (iters_out) iters_out[it - 1] = iters;
        if (relres_out) relres_out[it - 1] = relres;
    }
    return LV_OK;
}

// ---- C ABI ----------------------------------------------------------------------------------------
extern "C" {

int32_t lv_pressure_create(LvHandle c) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    return lv_pr_ensure(c);
}
int32_t lv_pressure_destroy(LvHandle c) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    double **one[] = {&c->d_mass, &c->d_rho, &c->d_c2, &c->d_P, &c->d_diag, &c->d_dinv, &c->d_b, &c->d_bvel, &c->d_vec[0], &c->d_vec[1],
                      &c->d_vec[2], &c->d_vec[3], &c->d_vec[4], &c->d_vec[5], &c->d_vec[6], &c->d_vec[7]};
    for (double **p : one) { lv_free(c, *p, sizeof(double) * (size_t)c->pr_cap); *p = nullptr; }
    lv_free(c, c->d_lrr, sizeof(double) * (size_t)c->cap_w); c->d_lrr = nullptr;
    lv_free(c, c->d_mx, sizeof(double2) * (size_t)c->cap_w); c->d_mx = nullptr;
    lv_free(c, c->d_mz, sizeof(double2) * (size_t)c->cap_w); c->d_mz = nullptr;
    lv_free(c, c->d_v, sizeof(double2) * (size_t)c->pr_cap); c->d_v = nullptr;
    lv_free(c, c->d_GP, sizeof(double2) * (size_t)c->pr_cap); c->d_GP = nullptr;
    lv_free(c, c->d_w, sizeof(double) * (size_t)c->cap_w); c->d_w = nullptr; c->cap_w = 0;
    c->pr_cap = 0; c->pr_valid = false; c->assembled = false;
    return LV_OK;
}

// staged != NULL: the five label-order arrays are already on their way into the device staging buffer on the upload stream
// of the pipelined mode (lv_pipe_upload_begin); the gathers wait for that stream
static int upload_fields(LvContext *c, const double *mass, const double *rho, const double *c2, const double *P, const double *v, bool dev,
                         const double *const *staged = nullptr) {
    LV_TRY(lv_pr_ensure(c));
    const int64_t n = c->n;
    struct Item { const double *src; double *dst; int nc; double fill; };
    Item items[] = {{mass, c->d_mass, 1, 0.0}, {rho, c->d_rho, 1, 1.0}, {c2, c->d_c2, 1, 1.0}, {P, c->d_P, 1, 0.0}, {v, (double *)c->d_v, 2, 0.0}};
    // label-order staging on the device: one persistent buffer with a sub-range per field, so that the five uploads
    // run back to back on the copy engine and no cudaMalloc / cudaFree (device-wide sync) sits on the path
    char *stage = nullptr;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    if (staged) { dev = true; LV_TRY(lv_pipe_upload_join(c)); }
    if (!dev) LV_TRY(lv_io_stage(c, (void **)&stage, sizeof(double) * 6 * nn));
    int st = LV_OK;
    size_t off = 0;
    const double *src_dev[5];
    int k = 0;
    for (const Item &it : items) {
        src_dev[k] = staged ? staged[k] : it.src;
        if (it.src && !dev) {
            cudaError_t e = cudaMemcpyAsync(stage + off, it.src, sizeof(double) * (size_t)it.nc * (size_t)n, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) { st = lv_set_error(c, LV_ECUDA, "field upload failed: %s", cudaGetErrorString(e)); break; }
            src_dev[k] = (const double *)(stage + off);
            off += sizeof(double) * (size_t)it.nc * nn;
        }
        k++;
    }
    k = 0;
    for (const Item &it : items) {
        if (st != LV_OK) break;
        if (it.src) st = lv_gather_to_slots(c, src_dev[k], it.dst, it.nc, it.fill);
        k++;
    }
    if (!dev || staged) cudaStreamSynchronize(c->stream); // the host buffers may be reused as soon as we return
    // neighbours' density, pressure and velocity across the strip edges
    if (st == LV_OK && rho) st = lv_halo_exchange(c, c->d_rho, 1);
    if (st == LV_OK && P) st = lv_halo_exchange(c, c->d_P, 1);
    if (st == LV_OK && v) st = lv_halo_exchange(c, (double *)c->d_v, 2);
    if (st == LV_OK && mass && rho && c2) c->pr_valid = true;
    if (mass || rho || c2) c->assembled = false;
    if (v) c->bvel_valid = false;
    return st;
}

int32_t lv_fields_upload(LvHandle c, const double *mass, const double *rho, const double *c2, const double *P, const double *v) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    return upload_fields(c, mass, rho, c2, P, v, false);
}
int32_t lv_fields_upload_dev(LvHandle c, const double *mass, const double *rho, const double *c2, const double *P, const double *v) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    return upload_fields(c, mass, rho, c2, P, v, true);
}

static int download_slots(LvContext *c, const double *src_slot, double *dst_host, int nc) {
    const int64_t n = c->n;
    if (n == 0) return LV_OK;
    // while an edge view is still streaming to the host the copy engine is taken: scatter straight into the caller's
    // buffer when it is pinned (mapped), otherwise through the staging buffer + cudaMemcpy
    double *direct = (c->stage_pending[0] || c->stage_pending[1] || lv_pipe_busy(c)) ? (double *)lv_mapped_alias(dst_host) : nullptr;
    void *stage = direct;
    if (!direct) LV_TRY(lv_io_stage(c, &stage, sizeof(double) * (size_t)nc * (size_t)n));
    int st = lv_scatter_to_labels(c, src_slot, (double *)stage, nc);
    if (st == LV_OK) {
        cudaError_t e = cudaSuccess;
        if (!direct) e = cudaMemcpyAsync(dst_host, stage, sizeof(double) * (size_t)nc * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) st = lv_set_error(c, LV_ECUDA, "download failed: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(c->stream);
    return st;
}

int32_t lv_pressure_download(LvHandle c, double *P_out) {
    if (!c || !P_out) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->d_P) return lv_set_error(c, LV_EINVAL, "no pressure workspace");
    return download_slots(c, c->d_P, P_out, 1);
}

int32_t lv_pressure_assemble(LvHandle c, double dt) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    return lv_pr_assemble(c, dt);
}

__global__ void __launch_bounds__(256) k_op_copy(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                 const int *__restrict__ rowptr_l, const int *__restrict__ col, const double *__restrict__ w,
                                                 const double *__restrict__ diag, const unsigned *__restrict__ ent_label,
                                                 long long *__restrict__ col_l, double *__restrict__ w_l, double *__restrict__ diag_l) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = prim[i];
    if (s < 0) { diag_l[i] = 0.0; return; }
    diag_l[i] = diag[s];
    int o = rowptr_l[i];
    for (int k = rowptr[s]; k < rowptr[s] + rdeg[s]; k++) {
        const int j = col[k];
        if (j < 0) continue;
        col_l[o] = (long long)(ent_label[j] & ~LV_IMAGE_BIT) + 1;
        w_l[o] = w[k];
        o++;
    }
}
__global__ void __launch_bounds__(256) k_op_deg(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                const int *__restrict__ col, int *__restrict__ deg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = prim[i];
    int d = 0;
    if (s >= 0)
        for (int k = rowptr[s]; k < rowptr[s] + rdeg[s]; k++) d += col[k] >= 0;
    deg[i] = d;
}

int32_t lv_pressure_operator(LvHandle c, int64_t *rowptr, int64_t *col, double *w, int64_t cap, double *diag) {
    if (!c || !rowptr) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->assembled) return lv_set_error(c, LV_EINVAL, "operator not assembled");
    const int64_t n = c->n, nnz = c->nnz;
    if (n == 0) { rowptr[0] = 0; return LV_OK; }
    size_t o_deg = 0, o_rl = sizeof(int) * (size_t)(n + 2), o_col = (o_rl + sizeof(int) * (size_t)(n + 2) + 15) & ~(size_t)15;
    size_t o_w = o_col + sizeof(long long) * (size_t)nnz, o_d = o_w + sizeof(double) * (size_t)nnz, total = o_d + sizeof(double) * (size_t)n + 64;
    void *stage = nullptr;
    LV_TRY(lv_alloc(c, &stage, total));
    char *base = (char *)stage;
    int *deg = (int *)(base + o_deg), *rl = (int *)(base + o_rl);
    long long *col_l = (long long *)(base + o_col);
    double *w_l = (double *)(base + o_w), *d_l = (double *)(base + o_d);
    const int nb = (int)((n + 255) / 256);
    int st = LV_OK;
    std::string msg;
    do {
        k_op_deg<<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, c->d_col, deg);
        c->launches++;
        if ((st = lv_exclusive_scan_i32(c, deg, rl, n)) != LV_OK) break;
        k_op_copy<<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, rl, c->d_col, c->d_w, c->d_diag, c->d_ent_label, col_l, w_l, d_l);
        c->launches++;
        int *h_rl = (int *)malloc(sizeof(int) * (size_t)(n + 1));
        cudaError_t e = cudaMemcpyAsync(h_rl, rl, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) {
            for (int64_t i = 0; i <= n; i++) rowptr[i] = h_rl[i];
            const int64_t m = h_rl[n];
            if ((col || w) && cap < m) st = lv_set_error(c, LV_ECAPACITY, "operator buffer too small: %lld > %lld", (long long)m, (long long)cap);
            else {
                if (col) e = cudaMemcpy(col, col_l, sizeof(long long) * (size_t)m, cudaMemcpyDeviceToHost);
                if (e == cudaSuccess && w) e = cudaMemcpy(w, w_l, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost);
                if (e == cudaSuccess && diag) e = cudaMemcpy(diag, d_l, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
            }
        }
        free(h_rl);
        if (e != cudaSuccess) st = lv_set_error(c, LV_ECUDA, "operator download failed: %s", cudaGetErrorString(e));
    } while (0);
    cudaStreamSynchronize(c->stream);
    lv_free(c, stage, total);
    return st;
}

int32_t lv_pressure_matvec(LvHandle c, const double *x, double *y) {
    if (!c || !x || !y) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->assembled) return lv_set_error(c, LV_EINVAL, "operator not assembled");
    const int64_t n = c->n;
    if (n == 0) return LV_OK;
    void *stage = nullptr;
    LV_TRY(lv_io_stage(c, &stage, sizeof(double) * (size_t)n));
    int st = LV_OK;
    cudaError_t e = cudaMemcpyAsync(stage, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) st = lv_set_error(c, LV_ECUDA, "upload failed: %s", cudaGetErrorString(e));
    if (st == LV_OK) st = lv_gather_to_slots(c, (const double *)stage, c->d_vec[3], 1, 0.0);
    if (st == LV_OK) st = lv_halo_exchange(c, c->d_vec[3], 1);
    if (st == LV_OK) st = lv_pr_matvec(c, c->d_vec[3], c->d_vec[4]);
    cudaStreamSynchronize(c->stream);
    if (st == LV_OK) st = download_slots(c, c->d_vec[4], y, 1);
    return st;
}

int32_t lv_pressure_rhs(LvHandle c, double dt, int32_t gp_step, const double *vbc_wall, double *b, double *GP) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(lv_pr_rhs(c, dt, gp_step, vbc_wall, false, nullptr, false));
    if (b) LV_TRY(download_slots(c, c->d_b, b, 1));
    if (GP) LV_TRY(download_slots(c, (const double *)c->d_GP, GP, 2));
    return LV_OK;
}

int32_t lv_find_pressure_dev(LvHandle c, double dt, int32_t niter, double rtol, double atol, int32_t itmax, int32_t solver,
                             const double *vbc_wall, int32_t *iters_out, double *relres_out) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    return lv_pr_find_pressure(c, dt, niter, rtol, atol, itmax, solver, vbc_wall, iters_out, relres_out);
}

int32_t lv_find_pressure(LvHandle c, double dt, int32_t niter, double rtol, double atol, int32_t itmax, int32_t solver,
                         const double *mass, const double *rho, const double *c2, const double *P_in, const double *v,
                         const double *vbc_wall, const double *vbc_edge, int64_t n_vbc_edge, double *P_out, int32_t *iters_out,
                         double *relres_out) {
    if (!c || !mass || !rho || !c2 || !P_in || !v || !P_out) return lv_set_error(c, LV_EINVAL, "null argument");
    if (c->pipe_mode && c->n > 0) {
        // pipelined mode: the five fields go up on the upload stream while the clip kernel of the remesh just queued is
        // still running; only then does the host wait for that remesh
        LV_CUDA(c, cudaSetDevice(c->device));
        const double *src[5] = {mass, rho, c2, P_in, v};
        const int nc[5] = {1, 1, 1, 1, 2};
        const double *staged[5];
        LV_TRY(lv_pipe_upload_begin(c, src, nc, staged));
        LV_TRY(lv_pipe_finish(c));
        LV_TRY(lv_set_boundary_velocity(c, vbc_edge, n_vbc_edge));
        LV_TRY(upload_fields(c, mass, rho, c2, P_in, v, false, staged));
        LV_TRY(lv_pr_find_pressure(c, dt, niter, rtol, atol, itmax, solver, vbc_wall, iters_out, relres_out));
        return download_slots(c, c->d_P, P_out, 1);
    }
    LV_ENTER(c);
    LV_TRY(lv_set_boundary_velocity(c, vbc_edge, n_vbc_edge)); // NULL: the four per-wall constants
    LV_TRY(upload_fields(c, mass, rho, c2, P_in, v, false));
    LV_TRY(lv_pr_find_pressure(c, dt, niter, rtol, atol, itmax, solver, vbc_wall, iters_out, relres_out));
    return download_slots(c, c->d_P, P_out, 1);
}

int32_t lv_pressure_solve(LvHandle c, int32_t solver, const double *b, double *x, double rtol, double atol, int32_t itmax,
                          int32_t *iters, double *relres) {
    if (!c || !b || !x) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->assembled) return lv_set_error(c, LV_EINVAL, "operator not assembled");
    const int64_t n = c->n;
    if (n == 0) return LV_OK;
    void *stage = nullptr;
    LV_TRY(lv_io_stage(c, &stage, sizeof(double) * (size_t)n));
    int st = LV_OK;
    const double *srcs[2] = {b, x};
    double *dsts[2] = {c->d_b, c->d_P};
    for (int k = 0; k < 2 && st == LV_OK; k++) {
        cudaError_t e = cudaMemcpyAsync(stage, srcs[k], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) st = lv_set_error(c, LV_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        if (st == LV_OK) st = lv_gather_to_slots(c, (const double *)stage, dsts[k], 1, 0.0);
        cudaStreamSynchronize(c->stream);
    }
    if (st == LV_OK) st = lv_pr_solve(c, solver, rtol, atol, itmax, iters, relres, false);
    if (st == LV_OK) st = download_slots(c, c->d_P, x, 1);
    return st;
}

} // extern "C"
