/*
 * lv_capi.h -- C ABI of liblvb200.so, the B200 (sm_100a) implementation of
 * LagrangianVoronoi.jl's per-timestep mesh-and-pressure hot path.
 *
 * The reference (pure Julia) has no FFI of its own; the drop-in boundary is its Julia call
 * surface.  Each entry point below names the reference body it replaces (file:line relative
 * to /root/reference/src).  The Julia shim that binds these symbols with `ccall` is in
 * julia/LagrangianVoronoiB200.jl and described in INTEGRATION.md; the Python host mirror in
 * lagrangianvoronoi.jl_b200/ binds the same symbols with ctypes.
 *
 * Conventions
 *  - every function returns an int32 status (LV_OK, ...); text via lv_last_error()
 *  - pointers are caller-owned HOST buffers unless the name ends in _dev (device pointers on
 *    the handle's GPU); the library owns all other device memory behind the opaque handle
 *  - generator labels are 1-based int64 at the boundary (Julia indices); labels <= 0 are the
 *    wall codes UP=-1 RIGHT=-2 DOWN=-3 LEFT=-4 (polygon.jl:4-7)
 *  - positions / vectors are interleaved double[2n] (= Vector{SVector{2,Float64}})
 *  - LvEdge is the 40-byte isbits layout of `Edge` (geometry.jl:82-87)
 *  - host-buffer calls are synchronous (lv_set_async_edges switches the lazy / pipelined modes on); *_dev calls are
 *    asynchronous on the handle's stream
 *    and report device-side failures at the next lv_sync()/host-buffer call
 *  - one caller thread per handle; no callbacks cross the boundary
 */
#ifndef LV_CAPI_H
#define LV_CAPI_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LvContext *LvHandle;

typedef struct LvEdge { double v1[2]; double v2[2]; int64_t label; } LvEdge; /* geometry.jl:82-87 */

/* VoronoiGrid{T}(boundary_rect, dr; h, r_max, xperiodic, yperiodic)  voronoigrid.jl:26-28.
 * h <= 0 selects the default 2dr, r_max <= 0 the default 10dr. */
typedef struct LvGridDesc {
    double dr, h, r_max;
    int32_t xperiodic, yperiodic;
    double bmin[2], bmax[2]; /* boundary_rect */
} LvGridDesc;

enum {
    LV_OK = 0,
    LV_EINVAL = 1,     /* invalid argument (also ArgumentError("h must be positive"), neighborlist.jl:19-21) */
    LV_EDESTROYED = 2, /* "The Voronoi Mesh has been destroyed."  voronoigrid.jl:63-65 */
    LV_ENAN = 3,       /* NaN/Inf position or velocity ("Velocity field invalidated." move.jl:24-26) */
    LV_ECUDA = 4,      /* CUDA / NCCL failure, see lv_last_error */
    LV_ECAPACITY = 5   /* caller buffer too small (nnz > cap) or polygon exceeded the internal limit */
};

enum { LV_SOLVER_CG = 0, LV_SOLVER_MINRES = 1, LV_SOLVER_PCG = 2 /* CG with the Jacobi preconditioner 1/A_ii; same stopping rule on ||r||_2 */ };

/* profiling slots for lv_prof_get */
enum {
    LV_PROF_CELLS = 0,   /* K1: cell-list build (count, scan, fill, bucket order) */
    LV_PROF_CLIP = 1,    /* K2: half-plane clipping kernel */
    LV_PROF_ASSEMBLE = 2,/* K3: operator + RHS assembly */
    LV_PROF_MATVEC = 3,  /* K4: Voronoi-Laplacian matvec (+ fused dot) */
    LV_PROF_VECOPS = 4,  /* K5: fused Krylov vector updates */
    LV_PROF_COUNT = 5
};

/* ---- lifetime ------------------------------------------------------------------------- */
/* replaces the VoronoiGrid constructor + CellList constructor (voronoigrid.jl:26-49,
 * neighborlist.jl:18-43).  `device` is the CUDA ordinal. */
int32_t lv_create(const LvGridDesc *desc, int32_t device, LvHandle *out);
int32_t lv_destroy(LvHandle h);
const char *lv_last_error(LvHandle h); /* h may be NULL: error of the last failed lv_create */
/* examples/piston.jl:43-47 mutates grid.boundary_rect / grid.cropping_rect at run time */
int32_t lv_set_rects(LvHandle h, const double bmin[2], const double bmax[2], const double cmin[2],
                     const double cmax[2]);
/* cell-list geometry (neighborlist.jl:23-25) and the length of the truncated magic_path */
int32_t lv_grid_info(LvHandle h, int64_t *n1, int64_t *n2, double origin[2], int64_t *npath);
int32_t lv_magic_path(LvHandle h, int64_t cap, int64_t *i1, int64_t *i2, double *rr, int64_t *count);
/* run on a caller-provided cudaStream_t (own = 0; NULL is the legacy default stream) or go back to the
 * handle's own non-blocking stream (own = 1) */
int32_t lv_set_stream(LvHandle h, void *cuda_stream, int32_t own);
int32_t lv_sync(LvHandle h); /* stream-synchronise and report any pending device-side status */

/* ---- remesh!(grid)  voronoigrid.jl:89-108 ----------------------------------------------- */
/* Host-buffer form.  xy[2n] in; out: rowptr[n+1] (0-based offsets), edges[cap] in the storage
 * order the reference leaves in p.edges after sort_edges! (IO.jl:35-48), *nnz, and optionally
 * area[n] (polygon.jl:114-122) and centroid[2n] (polygon.jl:210-219).  rowptr/edges/area/
 * centroid may each be NULL (edges NULL: the mesh stays device-resident only). */
int32_t lv_remesh(LvHandle h, int64_t n, const double *xy, int64_t *rowptr, LvEdge *edges, int64_t cap,
                  int64_t *nnz, double *area, double *centroid);
/* Device-resident form: positions already in HBM, results stay in HBM. */
int32_t lv_remesh_dev(LvHandle h, int64_t n, const double *xy_dev);
int32_t lv_mesh_nnz(LvHandle h, int64_t *nnz); /* synchronises */
int32_t lv_mesh_download(LvHandle h, int64_t *rowptr, LvEdge *edges, int64_t cap, double *area,
                         double *centroid);
/* Lazy edge view: after lv_set_async_edges(h, 1), lv_remesh / lv_mesh_download return once rowptr, areas and centroids
 * are on the host and copy the 40-byte edge records on a second stream, overlapping whatever runs next (typically
 * find_pressure!); the host must call lv_mesh_wait before it reads the edge buffer.  The buffer should be pinned.
 * While such a copy is in flight the copy engine is busy, so rowptr / area / centroid (lv_remesh) and P (lv_find_pressure,
 * lv_pressure_download) are written straight into the caller's buffers by the conversion kernels when those buffers are
 * pinned (cudaHostAlloc / cudaHostRegister: mapped under UVA); pageable buffers take the ordinary cudaMemcpy path.
 * on = 2 ("everything lazy"): rowptr, areas and centroids travel in the background too, ahead of the edge records on the
 * same copy stream; lv_remesh returns as soon as the device-side conversion is queued and NO output buffer may be read
 * before lv_mesh_wait.  For callers that launch the next device call (a second remesh, find_pressure!) straight away.
 * on = 3 ("pipelined", lv_pipeline.cu): lv_remesh additionally returns while its clipping kernel is still QUEUED (*nnz = -1;
 * lv_mesh_nnz after lv_mesh_wait).  The next lv_remesh / lv_find_pressure issues its uploads on a separate stream first and
 * only then completes the pending remesh, so host->device copies overlap the kernel; errors of a deferred remesh
 * ("The Voronoi Mesh has been destroyed.", capacity) are reported by that next call or by lv_mesh_wait.  The mesh crosses
 * PCIe as 20 B per edge (start vertex + 32-bit label word; the end vertex of an edge is the start vertex of its successor
 * in the chain sort_edges! leaves, checked bit for bit on the device) and host threads of the library (LV_HOST_THREADS,
 * default min(16, cores - 2)) expand it into the caller's 40-byte records; a mesh with an open chain is delivered as full
 * records instead.  NO output buffer may be read before lv_mesh_wait.  lv_mesh_download in this mode sends the mesh that
 * stands (after lv_remesh_dev / lv_strip_remesh) the same way and returns at once; the next remesh first settles such a
 * download (full records if the chains are open) before it replaces the mesh.
 * Environment switches (diagnostics): LV_DIRECT_STORE=0 disables the direct stores, LV_FLAG_MODE=memcpy reads status
 * words with cudaMemcpy instead of mapped memory. */
int32_t lv_set_async_edges(LvHandle h, int32_t on);
int32_t lv_mesh_wait(LvHandle h);
/* Decoder of the 20 B/edge wire format of the pipelined mode (host-only, no device needed): `len` consecutive edges of
 * the label-order edge list; v1[2*(len + has_next)] = their start vertices (+ the next edge's when has_next), word[len]:
 * bits 0-29 = 1-based label or wall number 1..4, bit 30 = wall, bit 31 = last edge of its row; row_start = start vertex
 * of the row the first edge belongs to.  Writes the 40-byte records (v2 = successor's v1, or the row's first v1). */
int32_t lv_wire_expand(const double *v1, const uint32_t *word, int64_t len, int32_t has_next, const double row_start[2],
                       LvEdge *out);
/* which kernel produced the current mesh: level 0/1 = linked-slot kernel (12/16 slots), 2/3/4 =
 * edge-list kernel (16/32/128 edges); anomalies = remeshes replayed by the edge-list kernel because
 * the linked-slot kernel met a degenerate configuration (results are identical either way) */
int32_t lv_clip_info(LvHandle h, int32_t *level, int64_t *anomalies);
/* Order-independent witness of the connectivity of the rows this handle owns: out[0..3] = 16-bit-chunk sums of
 * mix64(id_i, id_j) over all (row, edge) pairs, out[4] = number of pairs, out[5] = number of rows; id = 1-based global
 * label (global_label_dev[local label], NULL: local label + 1; wall codes -1..-4 as they are).  Sums of several ranks
 * add up, so a strip-decomposed mesh can be compared with the single-GPU mesh without gathering it. */
int32_t lv_mesh_hash(LvHandle h, const int32_t *global_label_dev, uint64_t out[6]);
/* face lengths len(e) (geometry.jl:136) and midpoints (geometry.jl:145) in the same CSR order */
int32_t lv_mesh_faces(LvHandle h, double *length, double *midpoint, int64_t cap);

/* ---- pressure system  pressure.jl:84-225 ------------------------------------------------ */
/* PressureSolver(grid) (pressure.jl:150-158): workspace for the current n */
int32_t lv_pressure_create(LvHandle h);
int32_t lv_pressure_destroy(LvHandle h);
/* per-polygon fields read by the solve, label order, any pointer may be NULL to keep the
 * resident value: mass[n] rho[n] c2[n] P[n] v[2n] */
int32_t lv_fields_upload(LvHandle h, const double *mass, const double *rho, const double *c2,
                         const double *P, const double *v);
int32_t lv_fields_upload_dev(LvHandle h, const double *mass_dev, const double *rho_dev,
                             const double *c2_dev, const double *P_dev, const double *v_dev);
int32_t lv_pressure_download(LvHandle h, double *P_out);
/* refresh!(A, grid, dt) (pressure.jl:104-117) on the resident fields */
int32_t lv_pressure_assemble(LvHandle h, double dt);
/* copy of A.neighbors / A.lr_ratios / A.diagonal in label order; col is 1-based (rowptr[n+1],
 * col/w[cap]); for parity tests of the assembly */
int32_t lv_pressure_operator(LvHandle h, int64_t *rowptr, int64_t *col, double *w, int64_t cap,
                             double *diag);
/* mul!(y, A, x) (pressure.jl:119-130), x and y in label order */
int32_t lv_pressure_matvec(LvHandle h, const double *x, double *y);
/* boundaries(p) (iterators.jl:50-57) of the whole mesh, numbered polygon by polygon in label order and, inside a polygon,
 * in edge order: *count, and for every boundary edge its midpoint (2 doubles), wall label (-1..-4) and 1-based polygon
 * (any array may be NULL; cap = capacity in edges).  This is what a host needs to evaluate the reference's closures
 * boundary_velocity(midpoint(e), e.label) (pressure.jl:182), vDirichlet(m) and charfun(m) (diffusion.jl:64-70) per edge. */
int32_t lv_boundary_edges(LvHandle h, int64_t *count, double *midpoint, int64_t *label, int64_t *polygon, int64_t cap);
/* boundary velocity per boundary edge (2 doubles each, numbering of lv_boundary_edges), used by every right-hand side built
 * from now on until the next remesh; NULL returns to the four per-wall constants vbc_wall */
int32_t lv_set_boundary_velocity(LvHandle h, const double *vbc_edge, int64_t n_edge);
/* refresh!(solver, dt, gp_step, boundary_velocity) (pressure.jl:162-203).  vbc_wall[4][2] is
 * the boundary velocity per wall code (row -label-1), NULL = zero_vbc (pressure.jl:205); per-edge values set with
 * lv_set_boundary_velocity take precedence.  Outputs b[n], GP[2n] in label order (either may be NULL). */
int32_t lv_pressure_rhs(LvHandle h, double dt, int32_t gp_step, const double *vbc_wall, double *b,
                        double *GP);
/* find_pressure!(solver, dt, niter; boundary_velocity) (pressure.jl:215-225).
 * Reference defaults: niter 10, rtol = atol = 1e-6, itmax 1000 (pressure.jl:219).
 * Host form gathers the fields, solves and writes P_out[n]; iters_out[niter] and
 * relres_out[niter] (true relative residual ||b - A P|| / ||b|| after each pass) may be NULL.
 * vbc_edge (nullable): boundary_velocity evaluated by the caller per boundary edge (2 doubles each, n_vbc_edge edges in
 * the numbering of lv_boundary_edges); NULL = the four per-wall constants vbc_wall (exact for the reference's examples). */
int32_t lv_find_pressure(LvHandle h, double dt, int32_t niter, double rtol, double atol, int32_t itmax,
                         int32_t solver, const double *mass, const double *rho, const double *c2,
                         const double *P_in, const double *v, const double *vbc_wall, const double *vbc_edge,
                         int64_t n_vbc_edge, double *P_out, int32_t *iters_out, double *relres_out);
/* Device-resident form on the fields uploaded with lv_fields_upload*; P stays resident. */
int32_t lv_find_pressure_dev(LvHandle h, double dt, int32_t niter, double rtol, double atol,
                             int32_t itmax, int32_t solver, const double *vbc_wall, int32_t *iters_out,
                             double *relres_out);
/* one Krylov solve A x = b on the assembled operator (label order, x holds the initial guess) */
int32_t lv_pressure_solve(LvHandle h, int32_t solver, const double *b, double *x, double rtol,
                          double atol, int32_t itmax, int32_t *iters, double *relres);

/* ---- device-resident time step: the callers either side of the hot path ------------ */
/* Polygon fields of @Euler_vars (celldefs.jl:7-27) kept in HBM in label order so that a whole step! runs without
 * PCIe round trips.  Field names: x v dv momentum (2 doubles per polygon), rho e P c2 mass energy quality mu phase
 * (1), D (4, column-major).  Set "x" first (it fixes n); every other field defaults to zero. */
int32_t lv_state_set(LvHandle h, const char *name, const double *host, int64_t n);
int32_t lv_state_get(LvHandle h, const char *name, double *host);
int32_t lv_state_ptr(LvHandle h, const char *name, void **dev_ptr, int64_t *n);
int32_t lv_state_remesh(LvHandle h);                                   /* remesh!(grid)            voronoigrid.jl:89-108 */
int32_t lv_step_move(LvHandle h, double dt);                           /* move!(grid, dt)          move.jl:9-33 (remeshes) */
int32_t lv_step_eos(LvHandle h, double gamma, double p0, int32_t stiffened); /* stiffened_eos!(grid, gamma, P0) pressure.jl:64-70 /
                                                                          ideal_eos!(grid, gamma; Pmin = p0) pressure.jl:49-55 */
int32_t lv_step_find_pressure(LvHandle h, double dt, int32_t niter, double rtol, double atol, int32_t itmax,
                              int32_t solver, const double *vbc_wall, int32_t *iters_out, double *relres_out);
int32_t lv_step_pressure_step(LvHandle h, double dt);                  /* pressure_step!           pressure.jl:10-25 */
int32_t lv_step_gravity(LvHandle h, double gx, double gy, double dt);  /* gravity_step!            pressure.jl:77-82 */
int32_t lv_step_find_D(LvHandle h);                                    /* find_D!                  diffusion.jl:8-19 */
int32_t lv_step_viscous_step(LvHandle h, double dt, int32_t artificial_viscosity); /* viscous_step! diffusion.jl:39-53 */
/* bdary_friction!(grid, vDirichlet, dt) diffusion.jl:64-80; vwall[4][2] = the closure's value on the walls UP, RIGHT, DOWN, LEFT
 * (NULL = all walls at rest), charfun = everywhere */
int32_t lv_step_bdary_friction(LvHandle h, double dt, const double *vwall);
/* general form: wall_on[4] (nullable = all) switches whole walls off (charfun constant per wall, e.g. examples/bubble.jl);
 * v_edge (2 doubles) / on_edge (1 byte) per boundary edge in the numbering of lv_boundary_edges carry vDirichlet(m) and
 * charfun(m) evaluated by the host at every boundary-edge midpoint (diffusion.jl:64-80); NULL falls back to vwall / on */
int32_t lv_step_bdary_friction_ex(LvHandle h, double dt, const double *vwall, const uint8_t *wall_on, const double *v_edge,
                                  const uint8_t *on_edge, int64_t n_edge);
int32_t lv_step_find_dv(LvHandle h, double dt, double alpha);          /* find_dv!                 relaxation.jl:10-25 */
int32_t lv_step_relaxation_step(LvHandle h, double dt, int32_t rusanov); /* relaxation_step!       relaxation.jl:36-73 (remeshes) */
/* multiphase_projection!(solver) (relaxation.jl:179-206; MultiphaseProjector mul! :91-123, refresh! :162-177).  Reference
 * settings: quality_threshold 0.25, atol = rtol = 1e-4, itmax 200; *solved = 0 where the reference would @warn */
int32_t lv_step_multiphase_projection(LvHandle h, double quality_threshold, double rtol, double atol, int32_t itmax,
                                      int32_t *iters, int32_t *solved);
/* y = A x with A the MultiphaseProjector (mul!, relaxation.jl:91-123) and b = its right-hand side from the resident dv
 * (refresh!, :162-177); host vectors in label order, NULL skips.  Exposed for direct operator parity checks. */
int32_t lv_step_multiphase_apply(LvHandle h, const double *x, double *y, double *b);
int32_t lv_step_lloyd(LvHandle h, int32_t niter);                      /* populate_lloyd! loop     populate.jl:132-145 */

/* ---- multi-GPU: y-strips, one process per GPU ------------------------------ */
/* The reference is shared-memory only; these entry points have no counterpart there.  The host
 * (lagrangianvoronoi.jl_b200/distributed.py) moves ghost generators between neighbouring strips with
 * torch.distributed and hands the library the NCCL id, the ownership mask and the halo plan. */
int32_t lv_comm_unique_id(uint8_t *out128);                       /* ncclGetUniqueId on rank 0 */
int32_t lv_comm_init(LvHandle h, int32_t rank, int32_t nranks, const uint8_t *id128);
/* remesh!(grid) on the generators present on this rank (owned + ghosts).  Only polygons with
 * owned_mask_dev[i] != 0 are clipped, ghosts are candidates only.  order_key_dev[i] (the global label,
 * < 2^31) orders the labels inside a bucket so that the candidate order -- hence every vertex bit -- is
 * the single-GPU order (`julia -t 1` order) regardless of how generators are laid out locally. */
int32_t lv_remesh_owned_dev(LvHandle h, int64_t n_local, const double *xy_dev, const uint8_t *owned_mask_dev,
                            const int32_t *order_key_dev);
/* device pointers into the slot-ordered cell list (0 ent_label u32, 1 prim_of_label i32, 2 own u8,
 * 3 ent_xy f64x2, 4 P f64, 5 area f64; strip mode: 6 local positions f64x2, 7 local global labels i32, 8 the same
 * array counted to the owned generators only) */
int32_t lv_device_array(LvHandle h, int32_t which, void **ptr, int64_t *count);
/* per peer rank: how many values this rank sends / receives and the (device) slot lists, concatenated */
int32_t lv_halo_plan(LvHandle h, int32_t npeers, const int32_t *peer_rank, const int64_t *send_count,
                     const int32_t *send_slots_dev, const int64_t *recv_count, const int32_t *recv_slots_dev);
/* Strip exchange over peer memory (csrc/lv_strip.cu): the library owns the local generator arrays (owned generators
 * first, then the ghosts peer by peer) and one exchange area per rank that its strip neighbours map once over CUDA IPC.
 * lv_strip_setup: peers (rank, this rank's index in the peer's own peer list, the bucket rows [lo, hi) that peer needs =
 *   its rows +- the halo), the ghost capacity per peer (same on every rank) and the capacity of the local arrays;
 *   returns the IPC handle (64 bytes) of this rank's area.  lv_strip_map: the peers' handles, in peer order.
 * lv_strip_set_owned: positions (host or device) and global labels (device int32; NULL keeps them) of the generators this
 *   rank owns.
 * lv_strip_remesh: remesh!(grid) (voronoigrid.jl:89-108) on owned + ghosts: ghost selection, count + generator exchange
 *   (NVLink pulls ordered by sequence words; one host synchronisation), cell list, clipping of the owned polygons with
 *   buckets ordered by global label (=> meshes do not depend on the GPU count), halo plan.  counts_out (nullable, 9 values):
 *   local generator count, ghosts sent to / received from each peer.  Collective over the strip neighbours.
 * Afterwards every halo exchange of the pressure solve (fields, right-hand side, Krylov vectors) pulls from the neighbours'
 * halo outboxes; the CG search direction is packed by the kernel that produces it. */
int32_t lv_strip_setup(LvHandle h, int32_t npeers, const int32_t *peer_rank, const int32_t *idx_there, const int32_t *lo,
                       const int32_t *hi, int64_t ghost_capacity, int64_t local_capacity, uint8_t *out64);
int32_t lv_strip_map(LvHandle h, const uint8_t *handles);
int32_t lv_strip_set_owned(LvHandle h, int64_t n_own, const double *xy, int32_t xy_on_host, const int32_t *global_label_dev);
int32_t lv_strip_remesh(LvHandle h, int64_t *counts_out);
/* bucket-row ownership of all ranks (R[0..world]: rank r owns rows [R[r], R[r+1])): needed by the migration that the
 * device-resident move! / relaxation_step! run on strips */
int32_t lv_strip_set_rows(LvHandle h, int32_t world, const int32_t *R);
/* Device-resident stepping on strips (move.jl:9-33, relaxation.jl:36-73 across GPUs): make this rank's owned generators
 * the resident state ("x" aliases the strip's position array; other fields via lv_state_set with n_own values).  Every
 * lv_step_* sweep then refreshes the ghost entries it reads from their owners (NVLink pulls), the moving sweeps hand
 * generators that left the strip -- with all their fields -- to the new owner before the strip remesh, and the multiphase
 * projector's MINRES sums its dot products over the ranks. */
int32_t lv_state_attach_strip(LvHandle h);
/* Peer-memory allreduce of the two CG scalars: every rank exports a mailbox (CUDA IPC handle, 64 bytes), maps the
 * mailboxes of all ranks and from then on posts / collects partial sums with NVLink stores and flags; sums are
 * taken in rank order, so the result is deterministic and identical on every rank. */
int32_t lv_mailbox_export(LvHandle h, uint8_t *out64);
int32_t lv_mailbox_plan(LvHandle h, int32_t nranks, const uint8_t *handles);
/* back to ncclSend/Recv halos and ncclAllReduce dots (every rank must call it when any rank failed to map a peer) */
int32_t lv_peer_disable(LvHandle h);
/* unmap all peer memory (call on every rank, then barrier, then lv_destroy) */
int32_t lv_peer_close(LvHandle h);
/* fill the ghost slots of a slot-ordered device vector (ncomp 1 or 2) from their owners */
int32_t lv_halo_exchange_dev(LvHandle h, double *vec_dev, int32_t ncomp);

/* ---- instrumentation -------------------------------------------------------------------- */
int32_t lv_prof_enable(LvHandle h, int32_t on);
int32_t lv_prof_reset(LvHandle h);
/* accumulated CUDA-event milliseconds and launch count of one profiling slot */
int32_t lv_prof_get(LvHandle h, int32_t slot, double *ms, int64_t *launches);
/* total kernels launched by this handle since creation */
int64_t lv_launch_count(LvHandle h);
/* device memory currently owned by the handle, in bytes */
int64_t lv_device_bytes(LvHandle h);

#ifdef __cplusplus
}
#endif
#endif
