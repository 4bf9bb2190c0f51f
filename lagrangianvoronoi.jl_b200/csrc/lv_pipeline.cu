// lv_pipeline.cu -- pipelined host-buffer mode of remesh! / find_pressure!  (lv_set_async_edges(h, 3)).
//
// The plain host-buffer calls are synchronous: the caller's next upload starts only after the previous kernel has
// finished, and every mesh leaves the device as 40-byte Edge records (geometry.jl:82-87), 4 GB per 16.8M-cell remesh --
// at PCIe rates that is as long as the step itself.  This mode keeps the call surface (lv_remesh, lv_find_pressure,
// lv_mesh_wait) and changes three things underneath:
//
//  * lv_remesh returns as soon as K1 has sized the slot arrays and K2 is QUEUED.  Its status words (overflow, anomaly,
//    nnz) land in a per-remesh snapshot in mapped pinned memory; the next call that needs the mesh (or the next remesh)
//    completes it -- waits for the kernel, reads the snapshot, climbs the capacity / exactness ladder of lv_clip_run if
//    it has to (lv_pipe_finish).  Uploads of the NEXT call (positions of the second remesh, the five fields of
//    find_pressure!) are issued on their own stream before that wait, so they run while K2 is still busy.
//  * the mesh crosses PCIe in a compact wire format, 20 B per edge instead of 40: the start vertex and a 32-bit word
//    (label or wall code, "last edge of the row" bit).  The end vertex of an edge is the start vertex of its successor in
//    the chain sort_edges! leaves (IO.jl:35-48); the conversion kernel checks that bit for bit and the remesh falls
//    back to full records when a chain does not close (degenerate input).
//  * a small thread pool inside the library receives the wire format chunk by chunk (ring of pinned buffers, one
//    cudaEvent per chunk) and expands it into the caller's Edge records with non-temporal stores while the GPU is
//    already on the next remesh / the pressure solve.  lv_mesh_wait returns when every record is in place.
//
// Nothing here computes mesh or pressure values on the CPU: the host threads only re-materialise the redundant end
// vertices of records the GPU produced.
#include "lv_internal.cuh"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <sched.h>
#include <thread>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

// Chunk = unit of the device->host copies and of the host-side expansion.  A worker expands 2^18 edges in ~1.2 ms and the
// copy engine delivers one in ~0.1 ms, so the ring must hold more chunks than there are workers or the ring, not PCIe,
// sets the pace (measured: 8 slots of 2^20 edges capped the download at 33 GB/s).
#define PIPE_CH (1 << 18) // edges per chunk of the wire format (4 MiB of vertices + 1 MiB of label words)
#define PIPE_RING 40      // pinned chunk buffers (200 MiB)
#define WIRE_END 0x80000000u
#define WIRE_WALL 0x40000000u
#define WIRE_PAYLOAD 0x3fffffffu

namespace {

struct PipeJob {
    int sb = 0;
    int64_t n = 0;
    // device staging (label order)
    const long long *d_rowptr64 = nullptr;
    const double *d_area = nullptr;
    const double2 *d_cen = nullptr, *d_v = nullptr, *d_hdr = nullptr;
    const unsigned *d_lab = nullptr;
    // host destinations
    int64_t *rowptr = nullptr;
    LvEdge *edges = nullptr;
    int64_t cap = 0;
    double *area = nullptr, *centroid = nullptr;
    const int *snap = nullptr; // status words of the clip attempt + [9] = "chains are not closed"
    cudaEvent_t ev_conv = nullptr;
    // results
    int64_t nnz_known = -1; // >= 0: download of a mesh that already stands (lv_pipe_download); the snapshot only carries [9]
    int decided = 0; // 0 not yet, 1 the job delivers the mesh, 2 it will not (status becomes 2)
    bool handled = false; // a skipped job whose mesh was delivered synchronously (lv_pipe_finish / lv_pipe_settle)
    int status = 0; // 0 queued / running, 1 done, 2 skipped (lv_pipe_finish replays or downloads synchronously), < 0 error
    int err_code = LV_OK;
    std::string err;
    std::atomic<int64_t> chunks_left{0};
};

struct ChunkTask {
    PipeJob *job;
    int slot;
    int64_t k0, len;
    bool has_next;
    double2 rs; // start vertex of the row the chunk begins in
};

} // namespace

struct LvPipe {
    LvContext *c = nullptr;
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_up = nullptr, ev_conv[2] = {nullptr, nullptr};
    double2 *d_xy_alt[2] = {nullptr, nullptr};
    int64_t cap_xy = 0;
    int xy_cur = 0;
    int *snap[2] = {nullptr, nullptr}; // mapped pinned, 16 ints each
    double2 *h_hdr[2] = {nullptr, nullptr}; // pinned chunk headers
    int64_t cap_hdr = 0;
    char *ring[PIPE_RING] = {nullptr};
    cudaEvent_t ring_ev[PIPE_RING] = {nullptr};
    bool ring_free[PIPE_RING];
    PipeJob *job[2] = {nullptr, nullptr}; // last job of each staging buffer
    int stage_cur = 0;
    cudaEvent_t ev_hdr = nullptr;
    int *d_clean_flags = nullptr; // 8 zero words: "status" of a mesh that already stands (lv_pipe_download)
    // deferred remesh: staging buffer / snapshot index and the caller's output buffers
    int pend_sb = 0;
    int64_t *pend_rowptr = nullptr;
    LvEdge *pend_edges = nullptr;
    int64_t pend_cap = 0;
    double *pend_area = nullptr, *pend_cen = nullptr;
    // threads
    std::mutex mu;
    std::condition_variable cv_jobs, cv_tasks, cv_ring, cv_done;
    std::deque<PipeJob *> jobs;
    std::deque<ChunkTask> tasks;
    bool stop = false;
    std::thread downloader;
    std::vector<std::thread> workers;
    int64_t bytes_d2h = 0; // wire bytes of the finished jobs (diagnostics)
    // LV_PIPE_TRACE=1: per-job timeline on stderr
    bool trace = false;
    std::chrono::steady_clock::time_point epoch;
    double tr_expand_ms = 0, tr_evwait_ms = 0; // summed over the workers, per job
};

static inline double pipe_now_ms(const LvPipe *P) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - P->epoch).count();
}

// ---- device side: label-order wire format -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pipe_deg(int64_t n, const int *__restrict__ prim, const unsigned char *__restrict__ rdeg,
                                                  const int *__restrict__ flags, int *__restrict__ deg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool bad = flags[LVF_OVERFLOW] | flags[LVF_NAN] | flags[LVF_DESTROYED]; // the attempt does not stand: nothing to convert
    const int s = prim[i];
    deg[i] = (!bad && s >= 0) ? rdeg[s] : 0;
}

// CHECK: compare every end vertex with its successor's start vertex.  The linked-slot kernel (levels 0-1) emits v2 as the
// very value it emits as the successor's v1, so its chains are closed by construction and the 16 B/edge of v2 stay unread;
// the edge-list kernel (levels 2-4) chains with ==, where -0.0 == 0.0.
template <bool CHECK>
__global__ void __launch_bounds__(256) k_pipe_copy(int64_t n, const int *__restrict__ prim, const int *__restrict__ rowptr,
                                                   const unsigned char *__restrict__ rdeg, const int *__restrict__ rowptr_l,
                                                   const int *__restrict__ col, const double2 *__restrict__ v1, const double2 *__restrict__ v2,
                                                   const unsigned *__restrict__ ent_label, const double *__restrict__ area,
                                                   const double2 *__restrict__ cen, const int *__restrict__ flags,
                                                   long long *__restrict__ rowptr64, double2 *__restrict__ vout, unsigned *__restrict__ lab,
                                                   double *__restrict__ area_l, double2 *__restrict__ cen_l, int *__restrict__ open_chain) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (flags[LVF_OVERFLOW] | flags[LVF_NAN] | flags[LVF_DESTROYED]) return;
    if (i == n) { rowptr64[n] = rowptr_l[n]; return; }
    const int s = prim[i];
    const int o = rowptr_l[i];
    rowptr64[i] = o;
    if (s < 0) { area_l[i] = 0.0; cen_l[i] = make_double2(0.0, 0.0); return; }
    const int r0 = rowptr[s], d = rdeg[s];
    area_l[i] = area[s];
    cen_l[i] = cen[s];
    bool open = false;
    for (int k = 0; k < d; k++) {
        const double2 a = v1[r0 + k];
        if (CHECK) {
            const double2 b = v2[r0 + k], nx = v1[r0 + (k + 1 < d ? k + 1 : 0)];
            // bitwise: -0.0 == 0.0 chains for sort_edges! but is not the same record
            open |= (__double_as_longlong(b.x) != __double_as_longlong(nx.x)) | (__double_as_longlong(b.y) != __double_as_longlong(nx.y));
        }
        const int cc = col[r0 + k];
        unsigned w = cc >= 0 ? ((ent_label[cc] & ~LV_IMAGE_BIT) + 1u) : (WIRE_WALL | (unsigned)(-cc));
        if (k == d - 1) w |= WIRE_END;
        vout[o + k] = a;
        lab[o + k] = w;
    }
    if (open) *open_chain = 1;
}

// start vertex of the row that chunk ci begins in
__global__ void k_pipe_hdr(int64_t n, const int *__restrict__ rowptr_l, const double2 *__restrict__ vout, const int *__restrict__ flags,
                           int nchunks, double2 *__restrict__ hdr) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nchunks) return;
    if (flags[LVF_OVERFLOW] | flags[LVF_NAN] | flags[LVF_DESTROYED]) return;
    const long long k0 = (long long)ci * PIPE_CH;
    if (k0 >= rowptr_l[n]) return;
    int64_t lo = 0, hi = n; // last i with rowptr_l[i] <= k0 (empty rows before it share the offset)
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (rowptr_l[mid] <= k0) lo = mid; else hi = mid - 1;
    }
    hdr[ci] = vout[rowptr_l[lo]];
}

__global__ void k_pipe_publish(const int *flags, const int *open_chain, int *snap) {
    const int t = threadIdx.x;
    if (t < 8) snap[t] = flags[t];
    if (t == 9) snap[9] = *open_chain;
    __threadfence_system();
}

// ---- host side: expansion of one chunk -------------------------------------------------------------------------------
static inline void store_rec(LvEdge *o, double2 a, double2 b, long long label) {
#if defined(__x86_64__)
    long long w[5];
    memcpy(&w[0], &a.x, 8); memcpy(&w[1], &a.y, 8); memcpy(&w[2], &b.x, 8); memcpy(&w[3], &b.y, 8);
    w[4] = label;
    long long *p = (long long *)o;
    _mm_stream_si64(p + 0, w[0]); _mm_stream_si64(p + 1, w[1]); _mm_stream_si64(p + 2, w[2]);
    _mm_stream_si64(p + 3, w[3]); _mm_stream_si64(p + 4, w[4]);
#else
    o->v1[0] = a.x; o->v1[1] = a.y; o->v2[0] = b.x; o->v2[1] = b.y; o->label = label;
#endif
}

static void expand_chunk(const double2 *V, const unsigned *lab, int64_t len, bool has_next, double2 rs, LvEdge *out) {
    for (int64_t k = 0; k < len; k++) {
        const unsigned u = lab[k];
        const double2 a = V[k];
        const bool end = (u & WIRE_END) != 0;
        const bool more = (k + 1 < len) || has_next;
        const double2 b = end ? rs : V[k + 1];
        const long long label = (u & WIRE_WALL) ? -(long long)(u & WIRE_PAYLOAD) : (long long)(u & WIRE_PAYLOAD);
        store_rec(out + k, a, b, label);
        if (end && more) rs = V[k + 1];
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

// Decoder of the wire format, exported for callers that want to check or reuse it: `len` edges starting somewhere in
// the label-order edge list; v1[2 * (len + has_next)] start vertices (has_next: one more vertex follows the chunk),
// word[len], row_start = start vertex of the row the first edge belongs to.  Host-only; needs no device.
extern "C" int32_t lv_wire_expand(const double *v1, const uint32_t *word, int64_t len, int32_t has_next, const double row_start[2],
                                  LvEdge *out) {
    if (len < 0 || (len > 0 && (!v1 || !word || !row_start || !out))) return LV_EINVAL;
    if (len > 0 && !has_next && !(word[len - 1] & WIRE_END)) return LV_EINVAL; // the list cannot stop inside a row
    expand_chunk((const double2 *)v1, word, len, has_next != 0, make_double2(row_start[0], row_start[1]), out);
    return LV_OK;
}

static void pipe_worker(LvPipe *P) {
    cudaSetDevice(P->c->device);
    for (;;) {
        ChunkTask t;
        {
            std::unique_lock<std::mutex> lk(P->mu);
            P->cv_tasks.wait(lk, [&] { return P->stop || !P->tasks.empty(); });
            if (P->tasks.empty()) return;
            t = P->tasks.front();
            P->tasks.pop_front();
        }
        const double tw0 = P->trace ? pipe_now_ms(P) : 0.0;
        cudaError_t e = cudaEventSynchronize(P->ring_ev[t.slot]);
        const double tw1 = P->trace ? pipe_now_ms(P) : 0.0;
        if (e == cudaSuccess) {
            const double2 *V = (const double2 *)P->ring[t.slot];
            const unsigned *lab = (const unsigned *)(P->ring[t.slot] + sizeof(double2) * (size_t)(PIPE_CH + 1));
            expand_chunk(V, lab, t.len, t.has_next, t.rs, t.job->edges + t.k0);
        }
        const double tw2 = P->trace ? pipe_now_ms(P) : 0.0;
        {
            std::lock_guard<std::mutex> lk(P->mu);
            P->tr_evwait_ms += tw1 - tw0; P->tr_expand_ms += tw2 - tw1;
            if (e != cudaSuccess && t.job->err_code == LV_OK) {
                t.job->err_code = LV_ECUDA;
                t.job->err = std::string("edge download failed: ") + cudaGetErrorString(e);
            }
            P->ring_free[t.slot] = true;
            t.job->chunks_left--;
        }
        P->cv_ring.notify_all();
        P->cv_done.notify_all();
    }
}

static void job_fail(LvPipe *P, PipeJob *j, int code, const std::string &msg) {
    std::lock_guard<std::mutex> lk(P->mu);
    if (j->err_code == LV_OK) { j->err_code = code; j->err = msg; }
}

// returns the final state of the job: 1 done, 2 skipped, -1 failed (err_code / err set)
static int pipe_run_job(LvPipe *P, PipeJob *j) {
    LvContext *c = P->c;
    cudaError_t e = cudaEventSynchronize(j->ev_conv);
    if (e != cudaSuccess) {
        job_fail(P, j, LV_ECUDA, std::string("remesh failed: ") + cudaGetErrorString(e));
        { std::lock_guard<std::mutex> lk(P->mu); j->decided = 2; j->handled = true; }
        P->cv_done.notify_all();
        return -1;
    }
    const int *hf = j->snap;
    {
        const bool skip = hf[LVF_NAN] || hf[LVF_DESTROYED] || hf[LVF_OVERFLOW] || hf[9]; // lv_pipe_finish / lv_pipe_settle deal with these
        {
            std::lock_guard<std::mutex> lk(P->mu);
            j->decided = skip ? 2 : 1;
        }
        P->cv_done.notify_all();
        if (skip) return 2;
    }
    const double tr0 = P->trace ? pipe_now_ms(P) : 0.0;
    double tr1 = 0.0;
    if (P->trace) { std::lock_guard<std::mutex> lk(P->mu); P->tr_expand_ms = P->tr_evwait_ms = 0.0; }
    const int64_t n = j->n, nnz = j->nnz_known >= 0 ? j->nnz_known : hf[LVF_NNZ];
    if (j->edges && j->cap < nnz) {
        char buf[128];
        snprintf(buf, sizeof(buf), "edge buffer too small: nnz = %lld, cap = %lld", (long long)nnz, (long long)j->cap);
        job_fail(P, j, LV_ECAPACITY, buf);
        return -1;
    }
    cudaStream_t cs = c->copy_stream;
    const int64_t nchunks = j->edges ? (nnz + PIPE_CH - 1) / PIPE_CH : 0;
    // the chunk headers first: they are read on this thread when the tasks are built
    if (nchunks > 0) e = cudaMemcpyAsync(P->h_hdr[j->sb], j->d_hdr, sizeof(double2) * (size_t)nchunks, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaEventRecord(P->ev_hdr, cs);
    if (e == cudaSuccess && j->rowptr) e = cudaMemcpyAsync(j->rowptr, j->d_rowptr64, sizeof(long long) * (size_t)(n + 1), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && j->area) e = cudaMemcpyAsync(j->area, j->d_area, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && j->centroid) e = cudaMemcpyAsync(j->centroid, j->d_cen, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaEventSynchronize(P->ev_hdr);
    if (e != cudaSuccess) { job_fail(P, j, LV_ECUDA, std::string("mesh download failed: ") + cudaGetErrorString(e)); return -1; }
    j->chunks_left = nchunks;
    for (int64_t ci = 0; ci < nchunks; ci++) {
        int slot = -1;
        {
            std::unique_lock<std::mutex> lk(P->mu);
            P->cv_ring.wait(lk, [&] {
                for (int s = 0; s < PIPE_RING; s++) if (P->ring_free[s]) { slot = s; return true; }
                return false;
            });
            P->ring_free[slot] = false;
        }
        const int64_t k0 = ci * PIPE_CH, len = std::min<int64_t>(PIPE_CH, nnz - k0);
        const bool has_next = k0 + len < nnz;
        char *dst = P->ring[slot];
        e = cudaMemcpyAsync(dst, j->d_v + k0, sizeof(double2) * (size_t)(len + (has_next ? 1 : 0)), cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dst + sizeof(double2) * (size_t)(PIPE_CH + 1), j->d_lab + k0, sizeof(unsigned) * (size_t)len, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaEventRecord(P->ring_ev[slot], cs);
        if (e != cudaSuccess) {
            job_fail(P, j, LV_ECUDA, std::string("edge download failed: ") + cudaGetErrorString(e));
            std::lock_guard<std::mutex> lk(P->mu);
            P->ring_free[slot] = true;
            j->chunks_left -= (nchunks - ci);
            break;
        }
        {
            std::lock_guard<std::mutex> lk(P->mu);
            P->tasks.push_back({j, slot, k0, len, has_next, P->h_hdr[j->sb][ci]});
        }
        P->cv_tasks.notify_one();
    }
    if (P->trace) tr1 = pipe_now_ms(P);
    // records of one mesh must be complete before the next mesh may write into the same caller buffers
    {
        std::unique_lock<std::mutex> lk(P->mu);
        P->cv_done.wait(lk, [&] { return j->chunks_left <= 0; });
    }
    if (nchunks == 0) e = cudaStreamSynchronize(cs); // per-cell arrays only
    if (e != cudaSuccess) job_fail(P, j, LV_ECUDA, std::string("mesh download failed: ") + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lk(P->mu);
    P->bytes_d2h += (int64_t)sizeof(long long) * (n + 1) + 24 * n + 20 * (j->edges ? nnz : 0);
    if (P->trace)
        fprintf(stderr, "[pipe] job sb=%d nnz=%lld: kernels done %.1f ms, copies issued %.1f, delivered %.1f | workers: expand %.1f ms, event wait %.1f ms (summed over %d threads)\n",
                j->sb, (long long)nnz, tr0, tr1, pipe_now_ms(P), P->tr_expand_ms, P->tr_evwait_ms, (int)P->workers.size());
    return j->err_code == LV_OK ? 1 : -1;
}

static void pipe_downloader(LvPipe *P) {
    cudaSetDevice(P->c->device);
    for (;;) {
        PipeJob *j;
        {
            std::unique_lock<std::mutex> lk(P->mu);
            P->cv_jobs.wait(lk, [&] { return P->stop || !P->jobs.empty(); });
            if (P->jobs.empty()) return;
            j = P->jobs.front();
            P->jobs.pop_front();
        }
        const int st = pipe_run_job(P, j);
        {
            std::lock_guard<std::mutex> lk(P->mu);
            j->status = st; // the job may be deleted by a waiter from here on
        }
        P->cv_done.notify_all();
    }
}

static int host_threads_default() {
    if (const char *e = getenv("LV_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return v > 64 ? 64 : v; }
    unsigned hc = std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = (unsigned)CPU_COUNT(&set);
#endif
    int ranks = 1; // torchrun: the ranks of a node share its cores
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > 1) ranks = v; }
    int t = ((int)hc - 2) / ranks;
    if (t < 3) t = 3; // the workers mostly wait (events, memory): a little oversubscription beats a starved download
    if (t > 16) t = 16;
    return t;
}

// ---- life cycle -------------------------------------------------------------------------------------------------------
int lv_pipe_enable(LvContext *c) {
    if (c->pipe) return LV_OK;
    LvPipe *P = new LvPipe();
    P->c = c;
    c->pipe = P;
    for (int s = 0; s < PIPE_RING; s++) P->ring_free[s] = true;
    if (!c->copy_stream) LV_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    LV_CUDA(c, cudaStreamCreateWithFlags(&P->up_stream, cudaStreamNonBlocking));
    LV_CUDA(c, cudaEventCreateWithFlags(&P->ev_up, cudaEventDisableTiming));
    LV_CUDA(c, cudaEventCreateWithFlags(&P->ev_hdr, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        LV_CUDA(c, cudaEventCreateWithFlags(&P->ev_conv[k], cudaEventDisableTiming));
        LV_CUDA(c, cudaHostAlloc((void **)&P->snap[k], sizeof(int) * 16, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(P->snap[k], 0, sizeof(int) * 16);
    }
    const size_t slot_bytes = sizeof(double2) * (size_t)(PIPE_CH + 1) + sizeof(unsigned) * (size_t)PIPE_CH;
    for (int s = 0; s < PIPE_RING; s++) {
        LV_CUDA(c, cudaHostAlloc((void **)&P->ring[s], slot_bytes, cudaHostAllocPortable));
        LV_CUDA(c, cudaEventCreateWithFlags(&P->ring_ev[s], cudaEventDisableTiming));
    }
    { const char *t = getenv("LV_PIPE_TRACE"); P->trace = t && t[0] == '1'; }
    P->epoch = std::chrono::steady_clock::now();
    LV_CUDA(c, cudaMalloc((void **)&P->d_clean_flags, sizeof(int) * 8));
    LV_CUDA(c, cudaMemset(P->d_clean_flags, 0, sizeof(int) * 8));
    P->downloader = std::thread(pipe_downloader, P);
    const int nw = host_threads_default();
    for (int k = 0; k < nw; k++) P->workers.emplace_back(pipe_worker, P);
    return LV_OK;
}

// every queued job has reached a final state; reports the first failure
int lv_pipe_drain(LvContext *c) {
    LvPipe *P = c->pipe;
    if (!P) return LV_OK;
    int code = LV_OK;
    std::string msg;
    {
        std::unique_lock<std::mutex> lk(P->mu);
        P->cv_done.wait(lk, [&] {
            if (!P->jobs.empty()) return false;
            for (int k = 0; k < 2; k++) if (P->job[k] && P->job[k]->status == 0) return false;
            return true;
        });
        for (int k = 0; k < 2; k++)
            if (P->job[k] && P->job[k]->status < 0 && code == LV_OK) { code = P->job[k]->err_code; msg = P->job[k]->err; }
        for (int k = 0; k < 2; k++) { delete P->job[k]; P->job[k] = nullptr; }
    }
    if (code != LV_OK) return lv_set_error(c, code, "%s", msg.c_str());
    return LV_OK;
}

bool lv_pipe_busy(const LvContext *c) {
    LvPipe *P = c->pipe;
    if (!P) return false;
    std::lock_guard<std::mutex> lk(P->mu);
    if (!P->jobs.empty()) return true;
    for (int k = 0; k < 2; k++) if (P->job[k] && P->job[k]->status == 0) return true;
    return false;
}

void lv_pipe_destroy(LvContext *c) {
    LvPipe *P = c->pipe;
    if (!P) return;
    lv_pipe_drain(c);
    {
        std::lock_guard<std::mutex> lk(P->mu);
        P->stop = true;
    }
    P->cv_jobs.notify_all();
    P->cv_tasks.notify_all();
    if (P->downloader.joinable()) P->downloader.join();
    for (auto &w : P->workers) if (w.joinable()) w.join();
    for (int s = 0; s < PIPE_RING; s++) {
        if (P->ring[s]) cudaFreeHost(P->ring[s]);
        if (P->ring_ev[s]) cudaEventDestroy(P->ring_ev[s]);
    }
    for (int k = 0; k < 2; k++) {
        if (P->snap[k]) cudaFreeHost(P->snap[k]);
        if (P->h_hdr[k]) cudaFreeHost(P->h_hdr[k]);
        if (P->ev_conv[k]) cudaEventDestroy(P->ev_conv[k]);
        if (P->d_xy_alt[k]) cudaFree(P->d_xy_alt[k]);
    }
    if (P->d_clean_flags) cudaFree(P->d_clean_flags);
    if (P->ev_up) cudaEventDestroy(P->ev_up);
    if (P->ev_hdr) cudaEventDestroy(P->ev_hdr);
    if (P->up_stream) cudaStreamDestroy(P->up_stream);
    delete P;
    c->pipe = nullptr;
    c->pipe_mode = false;
    c->pipe_pending = false;
}

// ---- the deferred remesh ------------------------------------------------------------------------------------------------
// queue conversion + snapshot behind the clip attempt and hand the download to the thread pool
static int pipe_queue_download(LvContext *c, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid,
                               int64_t nnz_known = -1) {
    LvPipe *P = c->pipe;
    // status words the conversion looks at: those of the clip attempt just queued, or "clean" for a mesh that stands
    // (d_flags is reused by the stepping and strip kernels after a remesh)
    const int *flags = nnz_known >= 0 ? P->d_clean_flags : c->d_flags;
    const int64_t n = c->n, capz = c->cap_nnz;
    const int sb = (P->stage_cur ^= 1);
    // the previous user of this staging buffer (two remeshes ago) must be through with it
    {
        std::unique_lock<std::mutex> lk(P->mu);
        P->cv_done.wait(lk, [&] { return !P->job[sb] || P->job[sb]->status != 0; });
        if (P->job[sb] && P->job[sb]->status < 0) {
            const int code = P->job[sb]->err_code;
            const std::string msg = P->job[sb]->err;
            delete P->job[sb];
            P->job[sb] = nullptr;
            lk.unlock();
            return lv_set_error(c, code, "%s", msg.c_str());
        }
        delete P->job[sb];
        P->job[sb] = nullptr;
    }
    const int64_t nchunks_cap = (capz + PIPE_CH - 1) / PIPE_CH + 1;
    if (nchunks_cap > P->cap_hdr) {
        for (int k = 0; k < 2; k++) {
            if (P->h_hdr[k]) { LV_TRY(lv_pipe_drain(c)); cudaFreeHost(P->h_hdr[k]); P->h_hdr[k] = nullptr; }
            LV_CUDA(c, cudaHostAlloc((void **)&P->h_hdr[k], sizeof(double2) * (size_t)(nchunks_cap + 16), cudaHostAllocPortable));
        }
        P->cap_hdr = nchunks_cap + 16;
    }
    // staging: deg[n+2] | rowptr_l[n+2] | rowptr64[n+2] | area[n] | cen[n] | hdr[nchunks] | open flag | v[cap] | lab[cap]
    size_t off_deg = 0, off_rl = off_deg + sizeof(int) * (size_t)(n + 2), off_r64 = (off_rl + sizeof(int) * (size_t)(n + 2) + 15) & ~(size_t)15;
    size_t off_area = off_r64 + sizeof(long long) * (size_t)(n + 2);
    size_t off_cen = off_area + sizeof(double) * (size_t)n;
    size_t off_hdr = (off_cen + sizeof(double2) * (size_t)n + 15) & ~(size_t)15;
    size_t off_open = off_hdr + sizeof(double2) * (size_t)nchunks_cap;
    size_t off_v = off_open + 16;
    size_t off_lab = off_v + sizeof(double2) * (size_t)(capz + 1);
    size_t total = off_lab + sizeof(unsigned) * (size_t)(capz + 1) + 64;
    LV_TRY(lv_ensure(c, &c->d_stage_buf[sb], &c->cap_stage_buf[sb], (int64_t)total, 1));
    char *base = (char *)c->d_stage_buf[sb];
    int *deg = (int *)(base + off_deg), *rl = (int *)(base + off_rl);
    long long *r64 = (long long *)(base + off_r64);
    double *area_l = (double *)(base + off_area);
    double2 *cen_l = (double2 *)(base + off_cen), *hdr = (double2 *)(base + off_hdr), *vout = (double2 *)(base + off_v);
    int *open_chain = (int *)(base + off_open);
    unsigned *lab = (unsigned *)(base + off_lab);
    const int nb = (int)((n + 256) / 256);
    LV_CUDA(c, cudaMemsetAsync(open_chain, 0, sizeof(int), c->stream));
    k_pipe_deg<<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_deg, flags, deg);
    c->launches++;
    LV_TRY(lv_exclusive_scan_i32(c, deg, rl, n));
    if (c->clip_last_level >= 2)
        k_pipe_copy<true><<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, rl, c->d_col, c->d_v1, c->d_v2, c->d_ent_label,
                                                     c->d_area, c->d_cen, flags, r64, vout, lab, area_l, cen_l, open_chain);
    else
        k_pipe_copy<false><<<nb, 256, 0, c->stream>>>(n, c->d_prim_of_label, c->d_rowptr, c->d_deg, rl, c->d_col, c->d_v1, c->d_v2, c->d_ent_label,
                                                      c->d_area, c->d_cen, flags, r64, vout, lab, area_l, cen_l, open_chain);
    k_pipe_hdr<<<(int)((nchunks_cap + 127) / 128), 128, 0, c->stream>>>(n, rl, vout, flags, (int)nchunks_cap, hdr);
    k_pipe_publish<<<1, 32, 0, c->stream>>>(flags, open_chain, P->snap[sb]);
    c->launches += 3;
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaEventRecord(P->ev_conv[sb], c->stream));
    PipeJob *j = new PipeJob();
    j->sb = sb; j->n = n;
    j->d_rowptr64 = r64; j->d_area = area_l; j->d_cen = cen_l; j->d_v = vout; j->d_hdr = hdr; j->d_lab = lab;
    j->rowptr = rowptr; j->edges = edges; j->cap = cap; j->area = area; j->centroid = centroid;
    j->snap = P->snap[sb];
    j->ev_conv = P->ev_conv[sb];
    j->nnz_known = nnz_known;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        P->job[sb] = j;
        P->jobs.push_back(j);
    }
    P->cv_jobs.notify_one();
    P->pend_sb = sb;
    return LV_OK;
}

int lv_clip_attempt_first(LvContext *c);
int lv_clip_resume(LvContext *c, const int *snap, bool *replayed);

int lv_pipe_remesh(LvContext *c, int64_t n, const double *xy, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid) {
    LvPipe *P = c->pipe;
    // 1. positions go up on the upload stream, into the buffer the previous remesh is NOT using: this copy runs while
    //    the previous clip kernel is still busy
    if (n > P->cap_xy) {
        LV_TRY(lv_pipe_finish(c));
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 2; k++) {
            if (P->d_xy_alt[k]) cudaFree(P->d_xy_alt[k]);
            P->d_xy_alt[k] = nullptr;
            LV_CUDA(c, cudaMalloc((void **)&P->d_xy_alt[k], sizeof(double2) * (size_t)(n + n / 16 + 64)));
        }
        P->cap_xy = n + n / 16 + 64;
    }
    if (P->trace) fprintf(stderr, "[pipe] remesh enter %.1f ms\n", pipe_now_ms(P));
    double2 *dst = P->d_xy_alt[P->xy_cur ^= 1];
    LV_CUDA(c, cudaMemcpyAsync(dst, xy, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, P->up_stream));
    LV_CUDA(c, cudaEventRecord(P->ev_up, P->up_stream));
    // 2. complete the previous remesh (host waits for its kernel; errors of that remesh surface here)
    LV_TRY(lv_pipe_finish(c));
    LV_TRY(lv_pipe_settle(c));
    LV_CUDA(c, cudaStreamWaitEvent(c->stream, P->ev_up, 0));
    c->xy = dst;
    c->owned_mask = nullptr;
    c->order_key = nullptr;
    // 3. K1 (synchronises once for the slot count), K2 queued, conversion + snapshot queued behind it
    c->mesh_valid = false; c->assembled = false; c->pr_valid = false; c->bvel_valid = false; c->bdry_valid = false; c->vbc_edge_on = false;
    c->n = n;
    LV_TRY(lv_cells_build(c));
    if (c->nslot == 0) { c->nnz = 0; c->mesh_valid = true; if (rowptr) rowptr[0] = 0; return LV_OK; }
    // the label-indexed arrays of the conversion need prim_of_label (K1) and the clip outputs
    LV_TRY(lv_clip_attempt_first(c));
    P->pend_rowptr = rowptr; P->pend_edges = edges; P->pend_cap = cap; P->pend_area = area; P->pend_cen = centroid;
    LV_TRY(pipe_queue_download(c, rowptr, edges, cap, area, centroid));
    c->pipe_pending = true;
    if (P->trace) fprintf(stderr, "[pipe] remesh queued %.1f ms\n", pipe_now_ms(P));
    return LV_OK;
}

// Completes the deferred remesh: host waits for clip + conversion, reads the snapshot; replays / falls back when the
// first attempt does not stand.  Called by every entry point (LV_ENTER) before it touches the mesh.
int lv_pipe_finish(LvContext *c) {
    LvPipe *P = c->pipe;
    if (!P || !c->pipe_pending) return LV_OK;
    c->pipe_pending = false;
    const int sb = P->pend_sb;
    LV_CUDA(c, cudaEventSynchronize(P->ev_conv[sb]));
    if (P->trace) fprintf(stderr, "[pipe] finish: clip + conversion done %.1f ms\n", pipe_now_ms(P));
    int snap[16];
    memcpy(snap, P->snap[sb], sizeof(snap));
    bool replayed = false;
    LV_TRY(lv_clip_resume(c, snap, &replayed));
    c->mesh_valid = true;
    if (!replayed && !snap[9]) return LV_OK; // the job in flight delivers this mesh
    // the job skipped itself (same snapshot): deliver synchronously with full records -- replayed meshes (capacity /
    // anomaly ladder) and meshes whose chains are not closed bit for bit
    {
        std::unique_lock<std::mutex> lk(P->mu);
        P->cv_done.wait(lk, [&] { return !P->job[sb] || P->job[sb]->status != 0; });
        if (P->job[sb]) P->job[sb]->handled = true;
    }
    LV_TRY(lv_pipe_drain(c));
    if (P->pend_rowptr || P->pend_edges || P->pend_area || P->pend_cen)
        LV_TRY(lv_mesh_to_labels(c, P->pend_rowptr, P->pend_edges, P->pend_cap, P->pend_area, P->pend_cen));
    return LV_OK;
}

// A download of the CURRENT mesh queued outside lv_pipe_remesh (lv_mesh_download in pipelined mode): the mesh stands, so
// the only way the job can decline is an open chain; lv_pipe_settle then delivers full records while the mesh is still
// there -- every remesh calls it before it touches the mesh arrays.
int lv_pipe_download(LvContext *c, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area, double *centroid) {
    if (edges && cap < c->nnz) return lv_set_error(c, LV_ECAPACITY, "edge buffer too small: nnz = %lld, cap = %lld", (long long)c->nnz, (long long)cap);
    return pipe_queue_download(c, rowptr, edges, cap, area, centroid, c->nnz);
}
int lv_pipe_settle(LvContext *c) {
    LvPipe *P = c->pipe;
    if (!P) return LV_OK;
    for (int k = 0; k < 2; k++) {
        PipeJob *j = nullptr;
        {
            std::unique_lock<std::mutex> lk(P->mu);
            if (!P->job[k]) continue;
            P->cv_done.wait(lk, [&] { return P->job[k]->decided != 0 || P->job[k]->status != 0; });
            if (P->job[k]->decided == 2 && !P->job[k]->handled && P->job[k]->err_code == LV_OK) { j = P->job[k]; j->handled = true; }
        }
        if (!j) continue;
        // wait for the job to leave the thread pool, then deliver synchronously (the staging buffers are shared)
        {
            std::unique_lock<std::mutex> lk(P->mu);
            P->cv_done.wait(lk, [&] { return j->status != 0; });
        }
        int64_t *r = j->rowptr; LvEdge *e = j->edges; const int64_t cp = j->cap; double *a = j->area, *ce = j->centroid;
        if (c->mesh_valid) LV_TRY(lv_mesh_to_labels(c, r, e, cp, a, ce)); // drains (and frees) the jobs first
    }
    return LV_OK;
}

int lv_pipe_wait(LvContext *c) {
    LV_TRY(lv_pipe_finish(c));
    LV_TRY(lv_pipe_settle(c));
    return lv_pipe_drain(c);
}

// uploads of find_pressure! on the upload stream, issued before the pending remesh is completed
int lv_pipe_upload_begin(LvContext *c, const double *const src[5], const int nc[5], const double *dev[5]) {
    LvPipe *P = c->pipe;
    const int64_t n = c->n;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    char *stage = nullptr;
    LV_TRY(lv_io_stage(c, (void **)&stage, sizeof(double) * 6 * nn));
    size_t off = 0;
    for (int k = 0; k < 5; k++) {
        dev[k] = nullptr;
        if (!src[k]) continue;
        LV_CUDA(c, cudaMemcpyAsync(stage + off, src[k], sizeof(double) * (size_t)nc[k] * (size_t)n, cudaMemcpyHostToDevice, P->up_stream));
        dev[k] = (const double *)(stage + off);
        off += sizeof(double) * (size_t)nc[k] * nn;
    }
    LV_CUDA(c, cudaEventRecord(P->ev_up, P->up_stream));
    return LV_OK;
}
int lv_pipe_upload_join(LvContext *c) {
    return cudaStreamWaitEvent(c->stream, c->pipe->ev_up, 0) == cudaSuccess ? LV_OK : lv_set_error(c, LV_ECUDA, "upload join failed");
}
