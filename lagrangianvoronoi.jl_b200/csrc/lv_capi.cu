// lv_capi.cu -- C-ABI entry points of liblvb200.so (see include/lv_capi.h), context
// management, the truncated magic_path table and the label-order views of the mesh.
#include "lv_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static thread_local std::string g_create_error;

int lv_set_error(LvContext *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

int lv_alloc(LvContext *c, void **ptr, size_t bytes) {
    *ptr = nullptr;
    if (bytes == 0) bytes = 16;
    LV_CUDA(c, cudaMalloc(ptr, bytes));
    c->dev_bytes += (int64_t)bytes;
    return LV_OK;
}
void lv_free(LvContext *c, void *ptr, size_t bytes) {
    if (!ptr) return;
    cudaFree(ptr);
    c->dev_bytes -= (int64_t)bytes;
}
// persistent staging for host<->device field transfers (contents undefined on return)
int lv_io_stage(LvContext *c, void **ptr, size_t bytes) {
    LV_TRY(lv_ensure(c, &c->d_io_stage, &c->cap_io_stage, (int64_t)(bytes ? bytes : 16), 1));
    *ptr = c->d_io_stage;
    return LV_OK;
}
// grow-only buffer; contents are NOT preserved
int lv_ensure(LvContext *c, void **ptr, int64_t *cap, int64_t need, size_t elt) {
    if (*ptr && *cap >= need) return LV_OK;
    if (*ptr) {
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        lv_free(c, *ptr, (size_t)*cap * elt);
        *ptr = nullptr;
        *cap = 0;
    }
    LV_TRY(lv_alloc(c, ptr, (size_t)need * elt));
    *cap = need;
    return LV_OK;
}

__global__ void k_publish_flags(const int *flags, const int *extra, int *host) {
    const int t = threadIdx.x;
    if (t < 8) host[t] = flags[t];
    if (t == 8 && extra) host[8] = *extra;
    __threadfence_system();
}
int lv_publish_flags(LvContext *c, const int *extra) {
    if (c->flags_mapped) {
        k_publish_flags<<<1, 32, 0, c->stream>>>(c->d_flags, extra, c->h_flags);
        c->launches++;
        LV_CUDA(c, cudaGetLastError());
    } else {
        if (extra) LV_CUDA(c, cudaMemcpyAsync(c->h_flags + 8, extra, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LV_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int) * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}

static cudaEvent_t prof_event(LvContext *c) {
    if (!c->prof_free.empty()) { cudaEvent_t e = c->prof_free.back(); c->prof_free.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
LvProfScope::LvProfScope(LvContext *c_, int slot_) : c(c_), slot(slot_) {
    if (c->prof_on) { a = prof_event(c); cudaEventRecord(a, c->stream); }
}
LvProfScope::~LvProfScope() {
    c->prof[slot].launches++;
    if (a) {
        cudaEvent_t b = prof_event(c);
        cudaEventRecord(b, c->stream);
        c->prof_pending.push_back({slot, a, b});
        if (c->prof_pending.size() > 8192) lv_prof_resolve(c);
    }
}
void lv_prof_resolve(LvContext *c) {
    for (auto &p : c->prof_pending) {
        cudaEventSynchronize(p.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->prof[p.slot].ms += ms;
        c->prof_free.push_back(p.a);
        c->prof_free.push_back(p.b);
    }
    c->prof_pending.clear();
}

// ---- magic_path (neighborlist.jl:29-41), truncated ---------------------------------------------
// The reference enumerates every offset i2 in (1-n1):(n1-1) (outer), i1 in (1-n2):(n2-1) (inner;
// the bounds are swapped in the source), rr = h^2*(max(0,|i1|-1)^2 + max(0,|i2|-1)^2), and sorts
// stably by i1^2+i2^2 and then stably by rr.  voronoicut!(grid, poly) stops at the first node
// with rr > prr and throws at the first node with rr > rr_max (voronoigrid.jl:57-65), so only
// the prefix up to and including the first node with rr > rr_max can ever be visited.  All such
// nodes have |i| <= K = floor(r_max/h)+3; enumerating only that window and applying the same
// two stable sorts yields the same prefix (stable sorts keep the relative order of the window's
// nodes, and every node outside the window has rr >= (K h)^2, larger than the sentinel's rr).
static void build_magic_path(double h, double r_max, int n1, int n2, std::vector<LvPathNode> &out) {
    const double rr_max = r_max * r_max;
    long K = (long)std::floor(r_max / h) + 3;
    long b2 = std::min<long>(n1 - 1, K), b1 = std::min<long>(n2 - 1, K);
    struct Node { long i1, i2; double rr; };
    std::vector<Node> nodes;
    for (long i2 = -b2; i2 <= b2; i2++)
        for (long i1 = -b1; i1 <= b1; i1++) {
            long a1 = std::max<long>(0, std::labs(i1) - 1), a2 = std::max<long>(0, std::labs(i2) - 1);
            nodes.push_back({i1, i2, (h * h) * (double)(a1 * a1 + a2 * a2)});
        }
    std::stable_sort(nodes.begin(), nodes.end(),
                     [](const Node &a, const Node &b) { return a.i1 * a.i1 + a.i2 * a.i2 < b.i1 * b.i1 + b.i2 * b.i2; });
    std::stable_sort(nodes.begin(), nodes.end(), [](const Node &a, const Node &b) { return a.rr < b.rr; });
    out.clear();
    for (const Node &nd : nodes) {
        out.push_back({(int)nd.i1, (int)nd.i2, nd.rr});
        if (nd.rr > rr_max) break; // sentinel: the walk can never pass it
    }
}

extern "C" {

int32_t lv_create(const LvGridDesc *d, int32_t device, LvHandle *out) {
    if (!d || !out) return lv_set_error(nullptr, LV_EINVAL, "null argument");
    *out = nullptr;
    const double dr = d->dr;
    const double h = d->h > 0.0 ? d->h : 2.0 * dr;          // voronoigrid.jl:27
    const double r_max = d->r_max > 0.0 ? d->r_max : 10.0 * dr;
    if (!(h > 0.0)) return lv_set_error(nullptr, LV_EINVAL, "h must be positive"); // neighborlist.jl:19-21
    if (!(d->bmax[0] > d->bmin[0]) || !(d->bmax[1] > d->bmin[1]))
        return lv_set_error(nullptr, LV_EINVAL, "empty boundary rectangle");
    LvContext *c = new LvContext();
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        lv_set_error(nullptr, LV_ECUDA, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
        delete c;
        return LV_ECUDA;
    }
    c->dr = dr; c->h = h; c->r_max = r_max;
    { const char *fm = getenv("LV_FLAG_MODE"); c->flags_mapped = !(fm && !strcmp(fm, "memcpy")); }
    const int xp = d->xperiodic != 0, yp = d->yperiodic != 0;
    for (int k = 0; k < 2; k++) { c->bmin[k] = d->bmin[k]; c->bmax[k] = d->bmax[k]; }
    // cropping_rect  voronoigrid.jl:29-30 (same association as the source)
    const double px = r_max * (double)xp, py = r_max * (double)yp;
    c->cmin[0] = (c->bmin[0] - px * 1.0) - py * 0.0; c->cmin[1] = (c->bmin[1] - px * 0.0) - py * 1.0;
    c->cmax[0] = (c->bmax[0] + px * 1.0) + py * 0.0; c->cmax[1] = (c->bmax[1] + px * 0.0) + py * 1.0;
    LvGridParams &g = c->gp;
    g.h = h;
    g.rr_max = r_max * r_max; // voronoigrid.jl:40
    g.xper = xp; g.yper = yp;
    g.xperiod = c->bmax[0] - c->bmin[0]; // voronoigrid.jl:31-32
    g.yperiod = c->bmax[1] - c->bmin[1];
    g.cminx = c->cmin[0]; g.cminy = c->cmin[1]; g.cmaxx = c->cmax[0]; g.cmaxy = c->cmax[1];
    g.ox = c->cmin[0] - h; g.oy = c->cmin[1] - h; // neighborlist.jl:23
    const double f1 = std::floor((c->cmax[0] - c->cmin[0]) / h), f2 = std::floor((c->cmax[1] - c->cmin[1]) / h);
    if (!(f1 >= 0 && f2 >= 0) || (f1 + 3.0) * (f2 + 3.0) > 2.0e9) {
        lv_set_error(nullptr, LV_EINVAL, "cell list of %g x %g buckets is out of range", f1 + 3.0, f2 + 3.0);
        delete c;
        return LV_EINVAL;
    }
    g.n1 = (int)f1 + 3; g.n2 = (int)f2 + 3; // neighborlist.jl:24-25
    c->ncell = (int64_t)g.n1 * g.n2;
    std::vector<LvPathNode> path;
    build_magic_path(h, r_max, g.n1, g.n2, path);
    g.npath = (int)path.size();
    if (g.npath > 4096) {
        lv_set_error(nullptr, LV_EINVAL, "r_max/h = %g needs a %d-node path table (limit 4096)", r_max / h, g.npath);
        delete c;
        return LV_EINVAL;
    }
    c->h_path = (LvPathNode *)malloc(sizeof(LvPathNode) * path.size());
    memcpy(c->h_path, path.data(), sizeof(LvPathNode) * path.size());
    int st = LV_OK;
    do {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { st = LV_ECUDA; break; }
        c->stream = c->own_stream;
        if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) { st = LV_ECUDA; break; }
        if ((st = lv_alloc(c, (void **)&c->d_path, sizeof(LvPathNode) * path.size())) != LV_OK) break;
        if (cudaMemcpy(c->d_path, path.data(), sizeof(LvPathNode) * path.size(), cudaMemcpyHostToDevice) != cudaSuccess) { st = LV_ECUDA; break; }
        if ((st = lv_alloc(c, (void **)&c->d_cell_cnt, sizeof(int) * (size_t)(c->ncell + 2))) != LV_OK) break;
        if ((st = lv_alloc(c, (void **)&c->d_cell_start, sizeof(int) * (size_t)(c->ncell + 2))) != LV_OK) break;
        if ((st = lv_alloc(c, (void **)&c->d_flags, sizeof(int) * 8)) != LV_OK) break;
        if ((st = lv_alloc(c, (void **)&c->d_tickets, sizeof(int) * 8)) != LV_OK) break;
        if (cudaMemset(c->d_tickets, 0, sizeof(int) * 8) != cudaSuccess) { st = LV_ECUDA; break; }
        if (cudaHostAlloc((void **)&c->h_flags, sizeof(int) * 16, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { st = LV_ECUDA; break; }
        if (cudaHostAlloc((void **)&c->h_red, sizeof(double) * 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { st = LV_ECUDA; break; }
    } while (0);
    if (st != LV_OK) {
        if (g_create_error.empty() || st == LV_ECUDA)
            lv_set_error(nullptr, st, "device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        lv_destroy(c);
        return st;
    }
    *out = c;
    return LV_OK;
}

int32_t lv_destroy(LvHandle c) {
    if (!c) return LV_OK;
    cudaSetDevice(c->device);
    lv_pipe_destroy(c); // joins the download threads before anything they read is freed
    cudaDeviceSynchronize();
    void *bufs[] = {c->d_path, c->d_xy, c->d_cell_cnt, c->d_cell_start, c->d_ent_label, c->d_ent_xy, c->d_prim_of_label,
                    c->d_rowptr, c->d_col, c->d_v1, c->d_v2, c->d_area, c->d_cen, c->d_tile_state, c->d_park_v, c->d_park_l, c->d_park_nxt, c->d_park_hdr, c->d_flags, c->d_tickets, c->d_bdry_ptr, c->d_vbc_edge, c->d_bf_on, c->d_scratch,
                    c->d_mass, c->d_rho, c->d_c2, c->d_P, c->d_v, c->d_GP, c->d_diag, c->d_dinv, c->d_w, c->d_b, c->d_red, c->d_lrr, c->d_mx, c->d_mz, c->d_bvel, c->d_deg, c->d_own, c->d_stage_buf[0], c->d_stage_buf[1], c->d_io_stage};
    for (void *b : bufs) if (b) cudaFree(b);
    for (double *v : c->d_vec) if (v) cudaFree(v);
    if (c->st_x_alias) c->st_field[0] = nullptr; // owned by the strip state
    for (double *v : c->st_field) if (v) cudaFree(v);
    if (c->st_tmp) cudaFree(c->st_tmp);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    if (c->h_red) cudaFreeHost(c->h_red);
    lv_dist_destroy(c);
    lv_prof_resolve(c);
    for (cudaEvent_t e : c->prof_free) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_conv_done) cudaEventDestroy(c->ev_conv_done);
    for (int k = 0; k < 2; k++) if (c->ev_stage_done[k]) cudaEventDestroy(c->ev_stage_done[k]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    free(c->h_path);
    delete c;
    return LV_OK;
}

const char *lv_last_error(LvHandle c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int32_t lv_set_rects(LvHandle c, const double bmin[2], const double bmax[2], const double cmin[2], const double cmax[2]) {
    if (!c || !bmin || !bmax || !cmin || !cmax) return lv_set_error(c, LV_EINVAL, "null argument");
    for (int k = 0; k < 2; k++) { c->bmin[k] = bmin[k]; c->bmax[k] = bmax[k]; c->cmin[k] = cmin[k]; c->cmax[k] = cmax[k]; }
    // only reset! reads the cropping rectangle (polygon.jl:37-47); the cell list and the periods
    // keep their construction-time values, as in the reference (examples/piston.jl:43-47)
    c->gp.cminx = cmin[0]; c->gp.cminy = cmin[1]; c->gp.cmaxx = cmax[0]; c->gp.cmaxy = cmax[1];
    return LV_OK;
}

int32_t lv_grid_info(LvHandle c, int64_t *n1, int64_t *n2, double origin[2], int64_t *npath) {
    if (!c) return LV_EINVAL;
    if (n1) *n1 = c->gp.n1;
    if (n2) *n2 = c->gp.n2;
    if (origin) { origin[0] = c->gp.ox; origin[1] = c->gp.oy; }
    if (npath) *npath = c->gp.npath;
    return LV_OK;
}

int32_t lv_magic_path(LvHandle c, int64_t cap, int64_t *i1, int64_t *i2, double *rr, int64_t *count) {
    if (!c) return LV_EINVAL;
    int64_t m = std::min<int64_t>(cap, c->gp.npath);
    for (int64_t k = 0; k < m; k++) {
        if (i1) i1[k] = c->h_path[k].i1;
        if (i2) i2[k] = c->h_path[k].i2;
        if (rr) rr[k] = c->h_path[k].rr;
    }
    if (count) *count = c->gp.npath;
    return LV_OK;
}

int32_t lv_set_stream(LvHandle c, void *stream, int32_t own) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    cudaStreamSynchronize(c->stream);
    // a NULL stream is a real stream (the legacy default stream torch uses): it must not fall back to the
    // handle's non-blocking stream, which does not order against it
    c->stream = own ? c->own_stream : (cudaStream_t)stream;
    return LV_OK;
}

int32_t lv_sync(LvHandle c) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}

// ---- remesh -----------------------------------------------------------------------------------
static int ensure_generators(LvContext *c, int64_t n, bool need_xy) {
    // labels (and the global ordering keys of lv_remesh_owned_dev) are packed as (key << 1 | image bit) into 32 bits by
    // the bucket sort (lv_cells.cu:k_cells_order), so they must stay below 2^30
    if (n < 0 || n >= ((int64_t)1 << 30)) return lv_set_error(c, LV_EINVAL, "n = %lld out of range (limit 2^30)", (long long)n);
    if (!c->d_prim_of_label || c->cap_n < n) {
        // grow both label-indexed buffers together so that cap_n describes both
        if (c->d_xy) { LV_CUDA(c, cudaStreamSynchronize(c->stream)); lv_free(c, c->d_xy, sizeof(double2) * (size_t)c->cap_n); c->d_xy = nullptr; }
        int64_t cap = c->d_prim_of_label ? c->cap_n : 0;
        const int64_t ncap = n + n / 16 + 64;
        LV_TRY(lv_ensure(c, (void **)&c->d_prim_of_label, &cap, ncap, sizeof(int)));
        c->cap_n = ncap;
    }
    if (need_xy && !c->d_xy) LV_TRY(lv_alloc(c, (void **)&c->d_xy, sizeof(double2) * (size_t)c->cap_n));
    return LV_OK;
}

} // extern "C"
int lv_remesh_common(LvContext *c, int64_t n) {
    if (c->pipe) LV_TRY(lv_pipe_settle(c)); // a pipelined download of the mesh about to be replaced may need full records
    c->mesh_valid = false;
    c->assembled = false;
    // the pressure fields live in slot order and every remesh re-sorts the slots: whatever was uploaded before is
    // permuted garbage now, so the "fields not uploaded" guards of lv_pr_assemble / lv_pr_rhs must fire again
    c->pr_valid = false;
    c->bvel_valid = false;
    c->bdry_valid = false;   // boundary-edge numbering and per-edge wall data belong to the old mesh
    c->vbc_edge_on = false;
    c->n = n;
    LV_TRY(lv_cells_build(c));
    LV_TRY(lv_clip_run(c));
    c->mesh_valid = true;
    return LV_OK;
}
extern "C" {

int32_t lv_remesh_dev(LvHandle c, int64_t n, const double *xy_dev) {
    if (!c || (n > 0 && !xy_dev)) return lv_set_error(c, LV_EINVAL, "null argument");
    LV_ENTER(c);
    LV_TRY(ensure_generators(c, n, false));
    c->xy = (const double2 *)xy_dev; // used in place: positions are only read
    int st = lv_remesh_common(c, n);
    c->owned_mask = nullptr; // one-shot: set by lv_remesh_owned_dev
    c->order_key = nullptr;
    return st;
}

int32_t lv_clip_info(LvHandle c, int32_t *level, int64_t *anomalies) {
    if (!c) return LV_EINVAL;
    if (level) *level = c->clip_last_level;
    if (anomalies) *anomalies = c->clip_anomalies;
    return LV_OK;
}

int32_t lv_mesh_nnz(LvHandle c, int64_t *nnz) {
    if (!c || !nnz) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->mesh_valid) return lv_set_error(c, LV_EINVAL, "no valid mesh: call lv_remesh first");
    *nnz = c->nnz;
    return LV_OK;
}

} // extern "C"

// label-order view: degree gather, scan, copy.  40-byte Edge records with 1-based labels.
__global__ void __launch_bounds__(256) k_label_deg(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                   int *__restrict__ deg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = prim[i];
    deg[i] = s >= 0 ? rdeg[s] : 0; // ghost generators (multi-GPU) have no row here
}

__global__ void __launch_bounds__(256) k_label_copy(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg,
                                                    const int *__restrict__ rowptr_l, const int *__restrict__ col,
                                                    const double2 *__restrict__ v1, const double2 *__restrict__ v2,
                                                    const unsigned *__restrict__ ent_label, const double *__restrict__ area,
                                                    const double2 *__restrict__ cen, long long *__restrict__ rowptr64,
                                                    LvEdge *__restrict__ edges, double *__restrict__ area_l,
                                                    double2 *__restrict__ cen_l) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { if (rowptr64) rowptr64[n] = rowptr_l[n]; return; }
    const int s = prim[i];
    const int o = rowptr_l[i];
    if (rowptr64) rowptr64[i] = o;
    if (s < 0) {
        if (area_l) area_l[i] = 0.0;
        if (cen_l) cen_l[i] = make_double2(0.0, 0.0);
        return;
    }
    const int r0 = rowptr[s], r1 = r0 + rdeg[s];
    if (area_l) area_l[i] = area[s];
    if (cen_l) cen_l[i] = cen[s];
    if (edges)
        for (int k = r0; k < r1; k++) {
            LvEdge e;
            const double2 a = v1[k], b = v2[k];
            e.v1[0] = a.x; e.v1[1] = a.y; e.v2[0] = b.x; e.v2[1] = b.y;
            const int cc = col[k];
            e.label = cc >= 0 ? (long long)(ent_label[cc] & ~LV_IMAGE_BIT) + 1 : (long long)cc;
            edges[o + (k - r0)] = e;
        }
}

// device-visible alias of a pinned (mapped) host buffer, or nullptr for pageable memory; LV_DIRECT_STORE=0 disables it
void *lv_mapped_alias(const void *host) {
    static const bool enabled = [] { const char *e = getenv("LV_DIRECT_STORE"); return !(e && e[0] == '0'); }();
    if (!enabled || !host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int lv_mesh_to_labels(LvContext *c, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid) {
    if (!c->mesh_valid) return lv_set_error(c, LV_EINVAL, "no valid mesh: call lv_remesh first");
    if (c->pipe) LV_TRY(lv_pipe_drain(c)); // the download threads of the pipelined mode read the same staging buffers
    const int64_t n = c->n, nnz = c->nnz;
    if (edges && cap < nnz) return lv_set_error(c, LV_ECAPACITY, "edge buffer too small: nnz = %lld, cap = %lld", (long long)nnz, (long long)cap);
    if (n == 0) { if (rowptr) rowptr[0] = 0; return LV_OK; }
    // staging: deg[n+1] | rowptr_l[n+1] | rowptr64[n+1] | area[n] | cen[2n] | edges[nnz]
    size_t off_deg = 0, off_rl = off_deg + sizeof(int) * (size_t)(n + 2), off_r64 = (off_rl + sizeof(int) * (size_t)(n + 2) + 15) & ~(size_t)15;
    size_t off_area = off_r64 + sizeof(long long) * (size_t)(n + 2);
    size_t off_cen = off_area + sizeof(double) * (size_t)n;
    size_t off_e = (off_cen + sizeof(double2) * (size_t)n + 15) & ~(size_t)15;
    size_t total = off_e + (edges ? sizeof(LvEdge) * (size_t)nnz : 0) + 64;
    // two staging buffers alternate in the lazy mode, so that the background copy of one edge view overlaps the next
    // remesh AND its conversion; a buffer is reused only after the copy that read it has finished
    const int sb = c->async_edges ? (c->stage_cur ^= 1) : 0;
    if (c->stage_pending[sb]) {
        if ((int64_t)total > c->cap_stage_buf[sb]) { LV_CUDA(c, cudaEventSynchronize(c->ev_stage_done[sb])); }
        else LV_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_stage_done[sb], 0));
        c->stage_pending[sb] = false;
    }
    LV_TRY(lv_ensure(c, &c->d_stage_buf[sb], &c->cap_stage_buf[sb], (int64_t)total, 1)); // grow-only: no cudaMalloc/cudaFree per remesh
    char *base = (char *)c->d_stage_buf[sb];
    int *deg = (int *)(base + off_deg), *rl = (int *)(base + off_rl);
    long long *r64 = (long long *)(base + off_r64);
    double *area_l = (double *)(base + off_area);
    double2 *cen_l = (double2 *)(base + off_cen);
    LvEdge *e_l = (LvEdge *)(base + off_e);
    const int nb = (int)((n + 256) / 256);
    int st = LV_OK;
    // lazy mode: the copy engine is busy with the previous edge view for ~80 ms, and a cudaMemcpy of the small per-cell
    // arrays would queue behind it.  When the caller's buffers are pinned (hence mapped under UVA) the conversion kernel
    // stores them straight into host memory instead: coalesced 256-512 B per warp, concurrent with the DMA.
    long long *r64_out = r64; double *area_out = area_l; double2 *cen_out = cen_l;
    bool direct_r = false, direct_a = false, direct_c = false;
    // async_all (lv_set_async_edges(h, 2)): the per-cell arrays travel in the background as well, ahead of the edge records on
    // the copy stream; nothing on the caller's critical path touches PCIe and lv_mesh_wait orders the host against all of it
    const bool all_lazy = c->async_edges && c->async_all && edges;
    if (c->async_edges && !all_lazy) {
        direct_r = rowptr && (r64_out = (long long *)lv_mapped_alias(rowptr)) != nullptr;
        direct_a = area && (area_out = (double *)lv_mapped_alias(area)) != nullptr;
        direct_c = centroid && (cen_out = (double2 *)lv_mapped_alias(centroid)) != nullptr;
        if (!direct_r) r64_out = r64;
        if (!direct_a) area_out = area_l;
        if (!direct_c) cen_out = cen_l;
    }
    do {
        k_label_deg<<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, deg);
        c->launches++;
        if ((st = lv_exclusive_scan_i32(c, deg, rl, n)) != LV_OK) break;
        k_label_copy<<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, rl, c->d_col, c->d_v1, c->d_v2, c->d_ent_label,
                                                c->d_area, c->d_cen, rowptr ? r64_out : nullptr, edges ? e_l : nullptr,
                                                area ? area_out : nullptr, centroid ? cen_out : nullptr);
        c->launches++;
        cudaError_t e = cudaGetLastError();
        cudaStream_t cell_stream = c->stream;
        if (all_lazy) {
            e = cudaEventRecord(c->ev_conv_done, c->stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, c->ev_conv_done, 0);
            cell_stream = c->copy_stream;
        }
        if (e == cudaSuccess && rowptr && !direct_r) e = cudaMemcpyAsync(rowptr, r64, sizeof(long long) * (size_t)(n + 1), cudaMemcpyDeviceToHost, cell_stream);
        if (e == cudaSuccess && area && !direct_a) e = cudaMemcpyAsync(area, area_l, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, cell_stream);
        if (e == cudaSuccess && centroid && !direct_c) e = cudaMemcpyAsync(centroid, cen_l, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost, cell_stream);
        if (e == cudaSuccess && edges) {
            if (c->async_edges) {
                // the 40 B/edge view is the bulk of the device->host traffic: copy it on a second stream so that it
                // overlaps whatever the caller runs next (lv_mesh_wait orders the host against it)
                if (!all_lazy) e = cudaEventRecord(c->ev_conv_done, c->stream);
                if (e == cudaSuccess && !all_lazy) e = cudaStreamWaitEvent(c->copy_stream, c->ev_conv_done, 0);
                if (e == cudaSuccess) e = cudaMemcpyAsync(edges, e_l, sizeof(LvEdge) * (size_t)nnz, cudaMemcpyDeviceToHost, c->copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(c->ev_stage_done[sb], c->copy_stream);
                c->stage_pending[sb] = (e == cudaSuccess);
            } else e = cudaMemcpyAsync(edges, e_l, sizeof(LvEdge) * (size_t)nnz, cudaMemcpyDeviceToHost, c->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) st = lv_set_error(c, LV_ECUDA, "mesh download failed: %s", cudaGetErrorString(e));
    } while (0);
    cudaStreamSynchronize(c->stream);
    return st;
}

extern "C" {

extern "C" int32_t lv_mesh_wait(LvHandle c);
// lazy edge view: with on != 0 lv_remesh / lv_mesh_download return as soon as rowptr, areas and centroids are on the
// host and stream the edge records in the background; lv_mesh_wait blocks until they have landed
int32_t lv_set_async_edges(LvHandle c, int32_t on) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    if (on && !c->copy_stream) {
        LV_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_conv_done, cudaEventDisableTiming));
        LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage_done[0], cudaEventDisableTiming));
        LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage_done[1], cudaEventDisableTiming));
    }
    if (!on || on == 3 || c->pipe_mode) LV_TRY(lv_mesh_wait(c));
    if (on == 3) LV_TRY(lv_pipe_enable(c));
    c->pipe_mode = on == 3;
    c->async_edges = on == 1 || on == 2;
    c->async_all = on == 2;
    return LV_OK;
}
int32_t lv_mesh_wait(LvHandle c) {
    if (!c) return LV_EINVAL;
    if (c->pipe) { LV_CUDA(c, cudaSetDevice(c->device)); LV_TRY(lv_pipe_wait(c)); }
    for (int k = 0; k < 2; k++)
        if (c->stage_pending[k]) {
            LV_ENTER(c);
            LV_CUDA(c, cudaEventSynchronize(c->ev_stage_done[k]));
            c->stage_pending[k] = false;
        }
    return LV_OK;
}

int32_t lv_mesh_download(LvHandle c, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    // pipelined mode: 20 B/edge wire format + host-side expansion in the background (NO output may be read before lv_mesh_wait)
    if (c->pipe_mode && c->mesh_valid && c->n > 0 && (rowptr || edges || area || centroid))
        return lv_pipe_download(c, rowptr, edges, cap, area, centroid);
    return lv_mesh_to_labels(c, rowptr, edges, cap, area, centroid);
}

int32_t lv_remesh(LvHandle c, int64_t n, const double *xy, int64_t *rowptr, LvEdge *edges, int64_t cap, int64_t *nnz,
                  double *area, double *centroid) {
    if (!c || (n > 0 && !xy)) return lv_set_error(c, LV_EINVAL, "null argument");
    if (c->pipe_mode && n > 0 && (rowptr || edges || area || centroid)) {
        // pipelined mode: the upload is issued BEFORE the previous remesh is completed (lv_pipe_remesh), the call returns
        // with the clip kernel queued, and *nnz is not known yet (-1; lv_mesh_nnz after lv_mesh_wait)
        LV_CUDA(c, cudaSetDevice(c->device));
        LV_TRY(ensure_generators(c, n, false));
        if (nnz) *nnz = -1;
        return lv_pipe_remesh(c, n, xy, rowptr, edges, cap, area, centroid);
    }
    LV_ENTER(c);
    LV_TRY(ensure_generators(c, n, true));
    if (n > 0) LV_CUDA(c, cudaMemcpyAsync(c->d_xy, xy, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    c->xy = c->d_xy;
    c->owned_mask = nullptr;
    c->order_key = nullptr;
    LV_TRY(lv_remesh_common(c, n));
    if (nnz) *nnz = c->nnz;
    if (rowptr || edges || area || centroid) LV_TRY(lv_mesh_to_labels(c, rowptr, edges, cap, area, centroid));
    return LV_OK;
}

__global__ void __launch_bounds__(256) k_faces(int64_t nnz, const double2 *__restrict__ v1, const double2 *__restrict__ v2,
                                               double *__restrict__ len, double2 *__restrict__ mid) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const double2 a = v1[k], b = v2[k];
    const double ex = a.x - b.x, ey = a.y - b.y;
    if (len) len[k] = sqrt(ex * ex + ey * ey);                                  // geometry.jl:136-138
    if (mid) mid[k] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));       // geometry.jl:145-147
}

int32_t lv_mesh_faces(LvHandle c, double *length, double *midpoint, int64_t cap) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->mesh_valid) return lv_set_error(c, LV_EINVAL, "no valid mesh: call lv_remesh first");
    const int64_t n = c->n, nnz = c->nnz;
    if (cap < nnz) return lv_set_error(c, LV_ECAPACITY, "face buffer too small");
    if (nnz == 0) return LV_OK;
    // label-order edges first, then lengths / midpoints of those records
    (void)n;
    std::vector<LvEdge> tmp((size_t)nnz); // host staging keeps this rarely used call simple
    LV_TRY(lv_mesh_to_labels(c, nullptr, tmp.data(), nnz, nullptr, nullptr));
    for (int64_t k = 0; k < nnz; k++) {
        const LvEdge &e = tmp[(size_t)k];
        const double ex = e.v1[0] - e.v2[0], ey = e.v1[1] - e.v2[1];
        if (length) length[k] = std::sqrt(ex * ex + ey * ey);
        if (midpoint) { midpoint[2 * k] = 0.5 * (e.v1[0] + e.v2[0]); midpoint[2 * k + 1] = 0.5 * (e.v1[1] + e.v2[1]); }
    }
    return LV_OK;
}

} // extern "C"

// ---- order-independent mesh witness ---------------------------------------------------------------
// Sum over all owned rows and their edges of mix64(id_i, id_j), id = 1-based GLOBAL label (wall codes stay
// negative), accumulated as four 16-bit-chunk sums so that partial sums of several ranks can be added with
// ordinary integer arithmetic.  Equal for any decomposition of the same generator set iff the connectivity is.
__device__ __forceinline__ unsigned long long lv_mix64(unsigned long long z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    return z;
}
__global__ void __launch_bounds__(256) k_mesh_hash(int nslot, const unsigned char *__restrict__ own, const unsigned *__restrict__ ent_label,
                                                   const int *__restrict__ rowptr, const unsigned char *__restrict__ rdeg, const int *__restrict__ col,
                                                   const int *__restrict__ key, unsigned long long *__restrict__ out /*[6]*/) {
    unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0, cnt = 0, rows = 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nslot; s += gridDim.x * blockDim.x) {
        if (!own[s]) continue;
        rows++;
        const unsigned li = ent_label[s] & ~LV_IMAGE_BIT;
        const long long gi = key ? (long long)key[li] : (long long)li + 1;
        const int r0 = rowptr[s], d = rdeg[s];
        for (int k = 0; k < d; k++) {
            const int j = col[r0 + k];
            long long gj = j;
            if (j >= 0) { const unsigned lj = ent_label[j] & ~LV_IMAGE_BIT; gj = key ? (long long)key[lj] : (long long)lj + 1; }
            const unsigned long long h = lv_mix64(lv_mix64((unsigned long long)gi) ^ ((unsigned long long)gj * 0x9E3779B97F4A7C15ull));
            a0 += h & 0xffffull; a1 += (h >> 16) & 0xffffull; a2 += (h >> 32) & 0xffffull; a3 += (h >> 48) & 0xffffull;
            cnt++;
        }
    }
    atomicAdd(&out[0], a0); atomicAdd(&out[1], a1); atomicAdd(&out[2], a2); atomicAdd(&out[3], a3);
    atomicAdd(&out[4], cnt); atomicAdd(&out[5], rows);
}

extern "C" int32_t lv_mesh_hash(LvHandle c, const int32_t *global_label_dev, uint64_t out[6]) {
    if (!c || !out) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->mesh_valid) return lv_set_error(c, LV_EINVAL, "no valid mesh: call lv_remesh first");
    for (int k = 0; k < 6; k++) out[k] = 0;
    if (c->nslot == 0) return LV_OK;
    LV_TRY(lv_ensure(c, &c->d_scratch, &c->cap_scratch, 64, 1));
    unsigned long long *d = (unsigned long long *)c->d_scratch;
    LV_CUDA(c, cudaMemsetAsync(d, 0, 48, c->stream));
    k_mesh_hash<<<c->num_sms * 4, 256, 0, c->stream>>>((int)c->nslot, c->d_own, c->d_ent_label, c->d_rowptr, c->d_deg, c->d_col, global_label_dev, d);
    c->launches++;
    LV_CUDA(c, cudaMemcpyAsync(out, d, 48, cudaMemcpyDeviceToHost, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}

// ---- boundary edges (boundaries(p), iterators.jl:50-57) numbered in CSR order ------------------------------------
__global__ void __launch_bounds__(256) k_bdry_count(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr,
                                                    const unsigned char *__restrict__ rdeg, const int *__restrict__ col, int *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = prim[i];
    int d = 0;
    if (s >= 0)
        for (int k = rowptr[s]; k < rowptr[s] + rdeg[s]; k++) d += col[k] < 0;
    cnt[i] = d;
}
__global__ void __launch_bounds__(256) k_bdry_list(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr,
                                                   const unsigned char *__restrict__ rdeg, const int *__restrict__ col,
                                                   const double2 *__restrict__ v1, const double2 *__restrict__ v2, const int *__restrict__ bptr,
                                                   double2 *__restrict__ mid, long long *__restrict__ label, long long *__restrict__ poly) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = prim[i];
    if (s < 0) return;
    int o = bptr[i];
    for (int k = rowptr[s]; k < rowptr[s] + rdeg[s]; k++) {
        const int cc = col[k];
        if (cc >= 0) continue;
        const double2 a = v1[k], b = v2[k];
        if (mid) mid[o] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y)); // midpoint(e)  geometry.jl:145-147
        if (label) label[o] = cc;
        if (poly) poly[o] = i + 1;
        o++;
    }
}

int lv_bdry_index(LvContext *c) {
    if (!c->mesh_valid) return lv_set_error(c, LV_EINVAL, "no valid mesh: call lv_remesh first");
    if (c->bdry_valid) return LV_OK;
    const int64_t n = c->n;
    c->n_bedge = 0;
    if (n > 0) {
        int64_t cap = c->cap_bdry;
        LV_TRY(lv_ensure(c, (void **)&c->d_bdry_ptr, &cap, 2 * (c->cap_n > n ? c->cap_n : n) + 8, sizeof(int)));
        c->cap_bdry = cap;
        int *cnt = c->d_bdry_ptr + (cap / 2);
        k_bdry_count<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, c->d_col, cnt);
        c->launches++;
        LV_TRY(lv_exclusive_scan_i32(c, cnt, c->d_bdry_ptr, n));
        int total = 0;
        LV_CUDA(c, cudaMemcpyAsync(&total, c->d_bdry_ptr + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        c->n_bedge = total;
    }
    c->bdry_valid = true;
    return LV_OK;
}

extern "C" int32_t lv_boundary_edges(LvHandle c, int64_t *count, double *midpoint, int64_t *label, int64_t *polygon, int64_t cap) {
    if (!c || !count) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(lv_bdry_index(c));
    *count = c->n_bedge;
    if (!midpoint && !label && !polygon) return LV_OK;
    if (cap < c->n_bedge) return lv_set_error(c, LV_ECAPACITY, "boundary-edge buffer too small: %lld > %lld", (long long)c->n_bedge, (long long)cap);
    const int64_t m = c->n_bedge;
    if (m == 0) return LV_OK;
    void *stage = nullptr;
    LV_TRY(lv_alloc(c, &stage, (size_t)m * 32));
    double2 *mid = (double2 *)stage;
    long long *lab = (long long *)((char *)stage + (size_t)m * 16), *pol = lab + m;
    k_bdry_list<<<(int)((c->n + 255) / 256), 256, 0, c->stream>>>(c->n, c->d_prim_of_label, c->d_rowptr, c->d_deg, c->d_col, c->d_v1, c->d_v2,
                                                                c->d_bdry_ptr, mid, lab, pol);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && midpoint) e = cudaMemcpyAsync(midpoint, mid, (size_t)m * 16, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && label) e = cudaMemcpyAsync(label, lab, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && polygon) e = cudaMemcpyAsync(polygon, pol, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream);
    lv_free(c, stage, (size_t)m * 32);
    if (e != cudaSuccess) return lv_set_error(c, LV_ECUDA, "boundary-edge download failed: %s", cudaGetErrorString(e));
    return LV_OK;
}

// boundary_velocity per boundary edge (order of lv_boundary_edges), used by every right-hand side until the next remesh;
// NULL returns to the four per-wall constants
extern "C" int32_t lv_set_boundary_velocity(LvHandle c, const double *vbc_edge, int64_t n_edge) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    c->bvel_valid = false;
    if (!vbc_edge) { c->vbc_edge_on = false; return LV_OK; }
    LV_TRY(lv_bdry_index(c));
    if (n_edge != c->n_bedge) return lv_set_error(c, LV_EINVAL, "boundary velocity for %lld edges, the mesh has %lld boundary edges", (long long)n_edge, (long long)c->n_bedge);
    if (n_edge > c->cap_bedge || !c->d_vbc_edge) {
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->d_vbc_edge) cudaFree(c->d_vbc_edge);
        if (c->d_bf_on) cudaFree(c->d_bf_on);
        c->cap_bedge = n_edge + n_edge / 8 + 64;
        LV_CUDA(c, cudaMalloc((void **)&c->d_vbc_edge, sizeof(double2) * (size_t)c->cap_bedge));
        LV_CUDA(c, cudaMalloc((void **)&c->d_bf_on, (size_t)c->cap_bedge));
    }
    if (n_edge > 0) LV_CUDA(c, cudaMemcpyAsync(c->d_vbc_edge, vbc_edge, sizeof(double2) * (size_t)n_edge, cudaMemcpyHostToDevice, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    c->vbc_edge_on = true;
    return LV_OK;
}

extern "C" {
// ---- instrumentation ----------------------------------------------------------------------------
int32_t lv_prof_enable(LvHandle c, int32_t on) { if (!c) return LV_EINVAL; c->prof_on = on != 0; return LV_OK; }
int32_t lv_prof_reset(LvHandle c) {
    if (!c) return LV_EINVAL;
    lv_prof_resolve(c);
    for (auto &p : c->prof) p = LvProfSlot();
    return LV_OK;
}
int32_t lv_prof_get(LvHandle c, int32_t slot, double *ms, int64_t *launches) {
    if (!c || slot < 0 || slot >= LV_PROF_COUNT) return LV_EINVAL;
    lv_prof_resolve(c);
    if (ms) *ms = c->prof[slot].ms;
    if (launches) *launches = c->prof[slot].launches;
    return LV_OK;
}
int64_t lv_launch_count(LvHandle c) { return c ? c->launches : 0; }
int64_t lv_device_bytes(LvHandle c) { return c ? c->dev_bytes : 0; }

} // extern "C"
