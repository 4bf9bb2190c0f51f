/*
 * lv_oracle.h -- CPU restatement of the LagrangianVoronoi.jl mesh-and-pressure path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and there only as the checker / CPU baseline.
 *
 * PARITY STATUS: the reference is pure Julia and cannot run in this environment
 * (no julia binary, no package depot, no network) and its test-suite holds no
 * golden vectors for this path.  The restatement is pinned only by
 *   (1) the reference's single integration test thresholds
 *       (tests/taylorgreen.jl:112-114, restated in tests/test_oracle_taylorgreen.py),
 *   (2) the semi-analytic Sedov profile the reference ships for examples/sedov.jl
 *       (examples/reference/sedov.csv; tests/test_oracle.py::test_sedov_blast_*), a physics-level pin,
 *   (3) invariants derived from the reference's definitions.
 * The Krylov.jl 0.9.8 MINRES iteration (Manifest.toml:465-469) is un-vendored
 * third-party code restated from its published algorithm: iterate-level
 * "parity unpinned".  Solution-level parity is solver independent.
 *
 * All citations are file:line relative to /root/reference/src unless noted.
 */
#ifndef LV_ORACLE_H
#define LV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double v1[2]; double v2[2]; int64_t label; } lvo_edge; /* geometry.jl:82-87 */

typedef struct lvo_grid lvo_grid;

/* status codes shared with include/lv_capi.h */
enum { LVO_OK = 0, LVO_EINVAL = 1, LVO_EDESTROYED = 2, LVO_ENAN = 3 };

/* voronoigrid.jl:26-49 + neighborlist.jl:18-43.  full_path != 0 builds the complete
 * (2n1-1)(2n2-1) magic_path like the reference; 0 builds the provably equivalent
 * truncated window (see lv_oracle.c:build_magic_path). */
lvo_grid *lvo_grid_create(const double bmin[2], const double bmax[2], double dr, double h,
                          double r_max, int xperiodic, int yperiodic, int full_path);
void lvo_grid_destroy(lvo_grid *g);
void lvo_set_rects(lvo_grid *g, const double bmin[2], const double bmax[2], const double cmin[2],
                   const double cmax[2]);
void lvo_grid_info(const lvo_grid *g, int64_t *n1, int64_t *n2, int64_t *npath, double origin[2]);
/* magic_path prefix: offsets (i1,i2) and rr of the first `cap` nodes */
int64_t lvo_magic_path(const lvo_grid *g, int64_t cap, int64_t *i1, int64_t *i2, double *rr);

/* polygons: push!(grid.polygons, T(x=x)) for every point; clears previous polygons */
int lvo_set_points(lvo_grid *g, int64_t n, const double *xy);
int64_t lvo_npolygons(const lvo_grid *g);

/* per-polygon field access (celldefs.jl:7-27). name in
 * x(2) rho v(2) e P c2 dv(2) mass momentum(2) energy phase quality D(4) mu */
int lvo_get_field(const lvo_grid *g, const char *name, double *out);
int lvo_set_field(lvo_grid *g, const char *name, const double *in);

/* voronoigrid.jl:89-108 */
int lvo_remesh(lvo_grid *g);
/* polygon edge lists in storage order (after sort_edges!), CSR over polygons */
int64_t lvo_nnz(const lvo_grid *g);
void lvo_get_mesh(const lvo_grid *g, int64_t *rowptr, lvo_edge *edges);
/* polygon.jl:114-122, 210-219 */
void lvo_area(const lvo_grid *g, double *area);
void lvo_centroid(const lvo_grid *g, double *cxy);

/* pressure.jl:89-117 operator assembly; CSR copy-out mirrors A.neighbors / A.lr_ratios / A.diagonal */
int lvo_pressure_assemble(lvo_grid *g, double dt);
void lvo_pressure_get_operator(const lvo_grid *g, int64_t *rowptr, int64_t *col, double *w,
                               double *diag);
/* pressure.jl:119-130 */
void lvo_pressure_matvec(const lvo_grid *g, const double *x, double *y);
/* pressure.jl:162-203; vbc_wall[4][2] indexed by -label-1 (UP,RIGHT,DOWN,LEFT) */
void lvo_pressure_rhs(lvo_grid *g, double dt, int gp_step, const double *vbc_wall, double *b,
                      double *P0, double *GP);
/* pressure.jl:215-225 with the MINRES restatement (solver=0) or plain CG (solver=1).
 * iters_out[niter], relres_out[niter] (true relative residual) may be NULL */
int lvo_find_pressure(lvo_grid *g, double dt, int niter, double rtol, double atol, int itmax,
                      int solver, const double *vbc_wall, int32_t *iters_out, double *relres_out);
/* stand-alone Krylov solves on the assembled operator; x holds the initial guess on entry */
int lvo_minres(const lvo_grid *g, const double *b, double *x, double rtol, double atol, int itmax,
               int warm_start);
int lvo_cg(const lvo_grid *g, const double *b, double *x, double rtol, double atol, int itmax);

/* callers either side of the hot path, needed to pin the oracle with
 * the reference's own Taylor-Green thresholds */
int lvo_populate_hex(lvo_grid *g);                                /* populate.jl:149-174 (no ic!) */
/* the other seeding strategies (charfun = everywhere, no ic!) and the relaxation loop of populate_lloyd! */
int lvo_populate_circ(lvo_grid *g, double cx, double cy);        /* populate.jl:18-35 */
int lvo_populate_rect(lvo_grid *g);                               /* populate.jl:46-66 */
int lvo_populate_vogel(lvo_grid *g, double cx, double cy);       /* populate.jl:105-121 */
int lvo_populate_rand(lvo_grid *g, const double *s, int64_t ns); /* populate.jl:76-95, uniform samples from the caller */
int lvo_lloyd(lvo_grid *g, int niter);                            /* populate.jl:132-145 */
int lvo_move(lvo_grid *g, double dt);                             /* move.jl:9-33 */
void lvo_stiffened_eos(lvo_grid *g, double gamma, double P0);     /* pressure.jl:64-70 */
void lvo_ideal_eos(lvo_grid *g, double gamma, double Pmin);       /* pressure.jl:49-55 */
void lvo_pressure_step(lvo_grid *g, double dt);                   /* pressure.jl:10-25 */
void lvo_find_D(lvo_grid *g);                                     /* diffusion.jl:8-19 */
void lvo_viscous_step(lvo_grid *g, double dt, int artificial_viscosity); /* diffusion.jl:39-53 */
void lvo_bdary_friction(lvo_grid *g, double dt, const double *vwall); /* diffusion.jl:64-80; vwall[4][2] by -label-1 */
/* the reference's closures evaluated per boundary edge by the caller: boundaries(p) numbered polygon by polygon */
int64_t lvo_boundary_edges(const lvo_grid *g, double *mid, int64_t *label, int64_t *polygon);
void lvo_set_vbc_edge(lvo_grid *g, const double *v, int64_t n); /* boundary_velocity per edge for the RHS (pressure.jl:182); NULL clears */
void lvo_bdary_friction_ex(lvo_grid *g, double dt, const double *vwall, const unsigned char *wall_on, const double *v_edge,
                           const unsigned char *on_edge); /* vDirichlet(m), charfun(m) per edge (diffusion.jl:64-80) */
void lvo_find_dv(lvo_grid *g, double dt, double alpha);           /* relaxation.jl:10-25 */
int lvo_relaxation_step(lvo_grid *g, double dt, int rusanov);     /* relaxation.jl:36-73 */

/* relaxation.jl:75-206 multiphase projector: MINRES (cold start) + correction of dv */
int lvo_multiphase_projection(lvo_grid *g, double quality_threshold, double rtol, double atol, int itmax, int *iters, int *solved);
void lvo_multiphase_apply(lvo_grid *g, const double *x, double *y, double *b); /* y = A x (:91-123), b = refresh! (:162-177); NULL skips */
void lvo_gravity_step(lvo_grid *g, double gx, double gy, double dt); /* pressure.jl:77-82 */

void lvo_set_threads(int nthreads);
int lvo_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
