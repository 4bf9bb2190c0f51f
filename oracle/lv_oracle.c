/*
 * lv_oracle.c -- CPU restatement of LagrangianVoronoi.jl's remesh + pressure path.
 *
 * TEST INFRASTRUCTURE ONLY (see lv_oracle.h).  Parity status: "parity unpinned" at the
 * Krylov-iterate level; pinned by tests/taylorgreen.jl:112-114 thresholds and derived
 * invariants otherwise.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).  No FMA
 * contraction: Julia does not contract a*b+c without @fastmath/muladd and the reference
 * uses neither.  Every function cites the reference lines it restates (relative to
 * /root/reference/src).  Data-structure shape follows the reference on purpose (array of
 * heap polygons each owning a grow-only edge vector; dense bucket grid of per-bucket
 * vectors with one lock each; per-row operator vectors) because this file is also the CPU
 * baseline that is timed beside the GPU path.
 *
 * Thread semantics: the reference's bucket contents are in ascending label order for
 * `julia -t 1` and racy for -t N.  Here buckets are filled under per-bucket locks like the
 * reference and then each bucket is put into ascending order, i.e. the -t 1 result,
 * independent of the thread count.
 */
#include "lv_oracle.h"
#include <math.h>
#include <omp.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ geometry.jl */
typedef struct { double x, y; } vec2;
static inline vec2 V(double x, double y) { vec2 r = {x, y}; return r; }
static inline vec2 vadd(vec2 a, vec2 b) { return V(a.x + b.x, a.y + b.y); }
static inline vec2 vsub(vec2 a, vec2 b) { return V(a.x - b.x, a.y - b.y); }
static inline vec2 vscale(double s, vec2 a) { return V(s * a.x, s * a.y); }
static inline double vdot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; } /* StaticArrays dot */
static inline double cross2(vec2 a, vec2 b) { return a.x * b.y - a.y * b.x; } /* geometry.jl:72-74 */
static inline double norm_squared(vec2 a) { return vdot(a, a); }             /* geometry.jl:163-165 */
static inline double vnorm(vec2 a) { return sqrt(a.x * a.x + a.y * a.y); }
static inline int veq(vec2 a, vec2 b) { return a.x == b.x && a.y == b.y; }
static inline int isnullvector(vec2 a) { return isnan(a.x) && isnan(a.y); }  /* geometry.jl:62-64 */

typedef struct { vec2 v1, v2; int64_t label; } edge_t;                        /* geometry.jl:82-87 */
static inline edge_t mkedge(vec2 a, vec2 b, int64_t label) { edge_t e = {a, b, label}; return e; }
static inline edge_t invert(edge_t e) { return mkedge(e.v2, e.v1, e.label); } /* IO.jl:5-7 */
static inline vec2 midpoint_e(edge_t e) { return vscale(0.5, vadd(e.v1, e.v2)); } /* geometry.jl:145-147 */
static inline vec2 midpoint_v(vec2 a, vec2 b) { return vscale(0.5, vadd(a, b)); } /* geometry.jl:154-156 */
static inline double len_e(edge_t e) { return vnorm(vsub(e.v1, e.v2)); }      /* geometry.jl:136-138 */

/* ------------------------------------------------------------------ fastvector.jl */
#define FV_DECL(NAME, T)                                                                    \
    typedef struct { T *data; int64_t last, cap; } NAME;                                    \
    static void NAME##_init(NAME *a, int64_t hint) { /* fastvector.jl:11 */                 \
        a->data = (T *)malloc(sizeof(T) * (size_t)(hint > 0 ? hint : 1));                   \
        a->last = 0; a->cap = hint > 0 ? hint : 1;                                          \
    }                                                                                       \
    static void NAME##_free(NAME *a) { free(a->data); a->data = NULL; a->last = a->cap = 0; } \
    static inline void NAME##_push(NAME *a, T val) { /* fastvector.jl:14-22 */              \
        a->last += 1;                                                                       \
        if (a->last > a->cap) {                                                             \
            a->cap *= 2;                                                                    \
            a->data = (T *)realloc(a->data, sizeof(T) * (size_t)a->cap);                    \
        }                                                                                   \
        a->data[a->last - 1] = val;                                                         \
    }                                                                                       \
    static inline void NAME##_deleteat(NAME *a, int64_t ind) { /* 1-based; fastvector.jl:44-50 */ \
        if (a->last > ind) a->data[ind - 1] = a->data[a->last - 1];                         \
        a->last -= 1;                                                                       \
    }
FV_DECL(fv_edge, edge_t)
FV_DECL(fv_int, int64_t)
FV_DECL(fv_f64, double)

/* ------------------------------------------------------------------ polygon.jl / celldefs.jl */
#define POLYGON_SIZEHINT 10           /* polygon.jl:1 */
#define SIGNUM_EPS (2.0 * 2.220446049250313e-16) /* polygon.jl:2  2*eps(Float64) */
#define BDARY_UP (-1)                 /* polygon.jl:4-7 */
#define BDARY_RIGHT (-2)
#define BDARY_DOWN (-3)
#define BDARY_LEFT (-4)
#define CELL_SIZEHINT 8               /* neighborlist.jl:1 */

static inline int signum(double x) { /* polygon.jl:9-16 */
    if (x < -SIGNUM_EPS) return -1;
    else if (x > SIGNUM_EPS) return 1;
    return 0;
}

typedef struct { /* celldefs.jl:7-27 (@Euler_vars) + PolygonNS :35-39 */
    vec2 x;
    double rho;
    vec2 v;
    double e, P, c2;
    vec2 dv;
    double mass;
    vec2 momentum;
    double energy;
    int64_t phase;
    fv_edge edges;
    double quality;
    double D[4]; /* column-major D11 D21 D12 D22 (geometry.jl:13) */
    double mu;
} poly_t;

typedef struct { int64_t i1, i2; double rr; } pathnode_t; /* neighborlist.jl:6-9 */

struct lvo_grid {
    /* voronoigrid.jl:14-25 */
    double dr, h, rr_max;
    vec2 bmin, bmax, cmin, cmax;
    int xperiodic, yperiodic;
    double xperiod, yperiod;
    poly_t **polygons;
    int64_t n;
    /* neighborlist.jl:11-17 */
    fv_int *cells;
    omp_lock_t *locks;
    vec2 origin;
    double clh;
    int64_t n1, n2;
    pathnode_t *path;
    int64_t npath;
    /* pressure.jl:89-93, 142-149 */
    fv_int *A_nb;
    fv_f64 *A_lrr;
    double *A_diag;
    int64_t A_n;
    double *b, *Psol;
    vec2 *GP;
    /* per-boundary-edge wall velocities: boundary_velocity(midpoint(e), e.label) evaluated by the caller (pressure.jl:182) */
    double *vbc_edge;
    int64_t n_vbc_edge;
};

static int g_threads = 0;
void lvo_set_threads(int nthreads) { g_threads = nthreads; omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs()); }
int lvo_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

/* stable merge sort of path nodes (Julia's default sort! is stable; neighborlist.jl:40-41) */
static void msort_nodes(pathnode_t *a, pathnode_t *tmp, int64_t n, int by_rr) {
    if (n < 2) return;
    int64_t m = n / 2;
    msort_nodes(a, tmp, m, by_rr);
    msort_nodes(a + m, tmp, n - m, by_rr);
    int64_t i = 0, j = m, k = 0;
    while (i < m && j < n) {
        int take_right;
        if (by_rr) take_right = a[j].rr < a[i].rr;
        else take_right = (a[j].i1 * a[j].i1 + a[j].i2 * a[j].i2) < (a[i].i1 * a[i].i1 + a[i].i2 * a[i].i2);
        tmp[k++] = take_right ? a[j++] : a[i++];
    }
    while (i < m) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(pathnode_t) * (size_t)n);
}

static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

/* neighborlist.jl:29-41.  The reference enumerates i2 in (1-n1):(n1-1) (outer) and i1 in
 * (1-n2):(n2-1) (inner) -- bounds swapped as in the source -- and sorts twice, stably.
 * With full_path == 0 only offsets |i| <= K = floor(r_max/h)+3 are generated.  This is
 * equivalent for voronoicut!(grid, poly) (voronoigrid.jl:57-65): every node outside the
 * window has rr >= (K h)^2 > ((K-1) h)^2 > rr_max, while the window already contains every
 * node with rr <= the smallest value above rr_max, and the walk stops (break or throw) at
 * the first node with rr > rr_max at the latest.  Relative order inside the window is
 * unchanged because both sorts are stable and the enumeration order is the same. */
static void build_magic_path(lvo_grid *g, double r_max, int full_path) {
    int64_t b2 = g->n1 - 1, b1 = g->n2 - 1; /* i2 bound uses n1, i1 bound uses n2 (sic) */
    if (!full_path) {
        int64_t K = (int64_t)floor(r_max / g->clh) + 3;
        b2 = imin64(b2, K);
        b1 = imin64(b1, K);
    }
    int64_t cnt = (2 * b2 + 1) * (2 * b1 + 1);
    pathnode_t *p = (pathnode_t *)malloc(sizeof(pathnode_t) * (size_t)cnt);
    pathnode_t *tmp = (pathnode_t *)malloc(sizeof(pathnode_t) * (size_t)cnt);
    int64_t k = 0;
    double h = g->clh;
    for (int64_t i2 = -b2; i2 <= b2; i2++)
        for (int64_t i1 = -b1; i1 <= b1; i1++) {
            int64_t a1 = imax64(0, llabs(i1) - 1), a2 = imax64(0, llabs(i2) - 1);
            double rr = (h * h) * (double)(a1 * a1 + a2 * a2); /* neighborlist.jl:34 */
            p[k].i1 = i1; p[k].i2 = i2; p[k].rr = rr; k++;
        }
    msort_nodes(p, tmp, cnt, 0); /* neighborlist.jl:40 */
    msort_nodes(p, tmp, cnt, 1); /* neighborlist.jl:41 */
    free(tmp);
    g->path = p;
    g->npath = cnt;
}

lvo_grid *lvo_grid_create(const double bmin[2], const double bmax[2], double dr, double h,
                          double r_max, int xperiodic, int yperiodic, int full_path) {
    if (!(h > 0.0)) return NULL; /* neighborlist.jl:19-21 */
    lvo_grid *g = (lvo_grid *)calloc(1, sizeof(lvo_grid));
    g->dr = dr; g->h = h; g->rr_max = r_max * r_max; /* voronoigrid.jl:38-40 */
    g->bmin = V(bmin[0], bmin[1]); g->bmax = V(bmax[0], bmax[1]);
    g->xperiodic = xperiodic != 0; g->yperiodic = yperiodic != 0;
    /* voronoigrid.jl:29-30: xmin - r_max*xperiodic*VECX - r_max*yperiodic*VECY */
    double px = r_max * (double)g->xperiodic, py = r_max * (double)g->yperiodic;
    g->cmin = V((g->bmin.x - px * 1.0) - py * 0.0, (g->bmin.y - px * 0.0) - py * 1.0);
    g->cmax = V((g->bmax.x + px * 1.0) + py * 0.0, (g->bmax.y + px * 0.0) + py * 1.0);
    g->xperiod = g->bmax.x - g->bmin.x; /* voronoigrid.jl:31-32 */
    g->yperiod = g->bmax.y - g->bmin.y;
    /* neighborlist.jl:23-27 */
    g->clh = h;
    g->origin = V(g->cmin.x - h, g->cmin.y - h);
    g->n1 = (int64_t)floor((g->cmax.x - g->cmin.x) / h) + 3;
    g->n2 = (int64_t)floor((g->cmax.y - g->cmin.y) / h) + 3;
    int64_t nc = g->n1 * g->n2;
    g->cells = (fv_int *)malloc(sizeof(fv_int) * (size_t)nc);
    g->locks = (omp_lock_t *)malloc(sizeof(omp_lock_t) * (size_t)nc);
    for (int64_t c = 0; c < nc; c++) { fv_int_init(&g->cells[c], CELL_SIZEHINT); omp_init_lock(&g->locks[c]); }
    build_magic_path(g, r_max, full_path);
    return g;
}

static void free_polygons(lvo_grid *g) {
    for (int64_t i = 0; i < g->n; i++) { fv_edge_free(&g->polygons[i]->edges); free(g->polygons[i]); }
    free(g->polygons); g->polygons = NULL; g->n = 0;
}
static void free_pressure(lvo_grid *g) {
    for (int64_t i = 0; i < g->A_n; i++) { fv_int_free(&g->A_nb[i]); fv_f64_free(&g->A_lrr[i]); }
    free(g->A_nb); free(g->A_lrr); free(g->A_diag); free(g->b); free(g->Psol); free(g->GP);
    g->A_nb = NULL; g->A_lrr = NULL; g->A_diag = NULL; g->b = g->Psol = NULL; g->GP = NULL; g->A_n = 0;
}

void lvo_grid_destroy(lvo_grid *g) {
    if (g) free(g->vbc_edge);
    if (!g) return;
    free_pressure(g);
    free_polygons(g);
    int64_t nc = g->n1 * g->n2;
    for (int64_t c = 0; c < nc; c++) { fv_int_free(&g->cells[c]); omp_destroy_lock(&g->locks[c]); }
    free(g->cells); free(g->locks); free(g->path); free(g);
}

void lvo_set_rects(lvo_grid *g, const double bmin[2], const double bmax[2], const double cmin[2],
                   const double cmax[2]) { /* examples/piston.jl:43-47 mutate the rectangles */
    g->bmin = V(bmin[0], bmin[1]); g->bmax = V(bmax[0], bmax[1]);
    g->cmin = V(cmin[0], cmin[1]); g->cmax = V(cmax[0], cmax[1]);
}

void lvo_grid_info(const lvo_grid *g, int64_t *n1, int64_t *n2, int64_t *npath, double origin[2]) {
    *n1 = g->n1; *n2 = g->n2; *npath = g->npath; origin[0] = g->origin.x; origin[1] = g->origin.y;
}
int64_t lvo_magic_path(const lvo_grid *g, int64_t cap, int64_t *i1, int64_t *i2, double *rr) {
    int64_t m = cap < g->npath ? cap : g->npath;
    for (int64_t k = 0; k < m; k++) { i1[k] = g->path[k].i1; i2[k] = g->path[k].i2; rr[k] = g->path[k].rr; }
    return m;
}

static poly_t *new_polygon(vec2 x) { /* celldefs.jl:7-27 defaults */
    poly_t *p = (poly_t *)calloc(1, sizeof(poly_t));
    p->x = x;
    fv_edge_init(&p->edges, POLYGON_SIZEHINT); /* polygon.jl:33 */
    return p;
}

int lvo_set_points(lvo_grid *g, int64_t n, const double *xy) {
    free_pressure(g);
    free_polygons(g);
    g->polygons = (poly_t **)malloc(sizeof(poly_t *) * (size_t)(n > 0 ? n : 1));
    g->n = n;
    for (int64_t i = 0; i < n; i++) g->polygons[i] = new_polygon(V(xy[2 * i], xy[2 * i + 1]));
    return LVO_OK;
}
int64_t lvo_npolygons(const lvo_grid *g) { return g->n; }

/* field access by name */
static int field_info(const char *name, size_t *off, int *ncomp, int *is_int) {
    *is_int = 0;
#define F(NM, MEMBER, NC) if (!strcmp(name, NM)) { *off = offsetof(poly_t, MEMBER); *ncomp = NC; return 1; }
    F("x", x, 2) F("rho", rho, 1) F("v", v, 2) F("e", e, 1) F("P", P, 1) F("c2", c2, 1) F("dv", dv, 2)
    F("mass", mass, 1) F("momentum", momentum, 2) F("energy", energy, 1) F("quality", quality, 1)
    F("D", D, 4) F("mu", mu, 1)
#undef F
    if (!strcmp(name, "phase")) { *off = offsetof(poly_t, phase); *ncomp = 1; *is_int = 1; return 1; }
    return 0;
}
int lvo_get_field(const lvo_grid *g, const char *name, double *out) {
    size_t off; int nc, is_int;
    if (!field_info(name, &off, &nc, &is_int)) return LVO_EINVAL;
    for (int64_t i = 0; i < g->n; i++) {
        const char *base = (const char *)g->polygons[i] + off;
        if (is_int) out[i] = (double)*(const int64_t *)base;
        else for (int c = 0; c < nc; c++) out[(int64_t)nc * i + c] = ((const double *)base)[c];
    }
    return LVO_OK;
}
int lvo_set_field(lvo_grid *g, const char *name, const double *in) {
    size_t off; int nc, is_int;
    if (!field_info(name, &off, &nc, &is_int)) return LVO_EINVAL;
    for (int64_t i = 0; i < g->n; i++) {
        char *base = (char *)g->polygons[i] + off;
        if (is_int) *(int64_t *)base = (int64_t)in[i];
        else for (int c = 0; c < nc; c++) ((double *)base)[c] = in[(int64_t)nc * i + c];
    }
    return LVO_OK;
}

/* ------------------------------------------------------------------ neighborlist.jl */
static inline int findkey(const lvo_grid *g, vec2 x, int64_t *i1, int64_t *i2) { /* neighborlist.jl:47-52 */
    vec2 d = vsub(x, g->origin);
    double q1 = floor(d.x / g->clh), q2 = floor(d.y / g->clh);
    /* floor(Int, .) throws InexactError for NaN/Inf/out-of-Int64 range */
    if (!(fabs(q1) < 9.0e18) || !(fabs(q2) < 9.0e18)) return 0;
    *i1 = (int64_t)q1 + 1; *i2 = (int64_t)q2 + 1;
    return 1;
}
static inline int inbounds(const lvo_grid *g, int64_t i1, int64_t i2) {
    return i1 >= 1 && i1 <= g->n1 && i2 >= 1 && i2 <= g->n2;
}
#define CELL(g, i1, i2) ((g)->cells[((i1) - 1) + (g)->n1 * ((i2) - 1)])

static int cl_insert(lvo_grid *g, vec2 x, int64_t label, int *bad) { /* neighborlist.jl:55-67 */
    int64_t i1, i2;
    if (!findkey(g, x, &i1, &i2)) { *bad = 1; return 0; }
    if (inbounds(g, i1, i2)) {
        int64_t c = (i1 - 1) + g->n1 * (i2 - 1);
        omp_set_lock(&g->locks[c]);
        fv_int_push(&g->cells[c], label);
        omp_unset_lock(&g->locks[c]);
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ polygon.jl */
static void reset_poly(poly_t *p, vec2 rmin, vec2 rmax) { /* polygon.jl:37-47 */
    p->edges.last = 0;
    vec2 A = rmin, C = rmax;
    vec2 B = V(C.x, A.y), D = V(A.x, C.y);
    fv_edge_push(&p->edges, mkedge(B, A, BDARY_DOWN));
    fv_edge_push(&p->edges, mkedge(A, D, BDARY_LEFT));
    fv_edge_push(&p->edges, mkedge(D, C, BDARY_UP));
    fv_edge_push(&p->edges, mkedge(C, B, BDARY_RIGHT));
}

static int voronoicut_poly(poly_t *p, vec2 y, int64_t label) { /* polygon.jl:51-97 */
    vec2 diff = vsub(y, p->x);
    vec2 mid = vscale(0.5, vadd(y, p->x));
    double c = vdot(diff, mid);
    int64_t i = 1;
    vec2 X = V(NAN, NAN), Y = V(NAN, NAN);
    while (i <= p->edges.last) {
        edge_t e = p->edges.data[i - 1];
        double f1 = vdot(diff, e.v1) - c;
        double f2 = vdot(diff, e.v2) - c;
        int s1 = signum(f1), s2 = signum(f2);
        int s12 = s1 + s2;
        if ((0 <= s12 && s12 <= 1) && ((s1 | s2) != 0)) {
            Y = X;
            /* X = 1.0/(f1 - f2)*(f1*e.v2 - f2*e.v1), left-associated */
            double r = 1.0 / (f1 - f2);
            X = V(r * (f1 * e.v2.x - f2 * e.v1.x), r * (f1 * e.v2.y - f2 * e.v1.y));
            if (s1 == 0) X = e.v1;
            if (s2 == 0) X = e.v2;
            edge_t ecut = (s1 == 1) ? mkedge(X, e.v2, e.label) : mkedge(e.v1, X, e.label);
            p->edges.data[i - 1] = ecut;
        }
        if (1 <= s12) fv_edge_deleteat(&p->edges, i);
        else i += 1;
    }
    if (!isnullvector(Y) && !veq(X, Y)) {
        edge_t e = mkedge(X, Y, label);
        if (cross2(vsub(Y, X), vsub(p->x, X)) > 0.0) e = invert(e);
        fv_edge_push(&p->edges, e);
        return 1;
    }
    return 0;
}

static double influence_rr(const poly_t *p) { /* polygon.jl:101-107 */
    double rr = 0.0;
    for (int64_t k = 0; k < p->edges.last; k++) {
        double t = 4.0 * norm_squared(vsub(p->edges.data[k].v1, p->x));
        if (isnan(t) || isnan(rr)) rr = NAN; /* Julia max propagates NaN */
        else rr = rr > t ? rr : t;
    }
    return rr;
}

static double poly_area(const poly_t *p) { /* polygon.jl:114-122 */
    double A = 0.0;
    for (int64_t k = 0; k < p->edges.last; k++) {
        edge_t e = p->edges.data[k];
        A += 0.5 * fabs(cross2(vsub(e.v1, p->x), vsub(e.v2, p->x)));
    }
    return A;
}

static vec2 poly_centroid(const poly_t *p) { /* polygon.jl:210-219 (tri_area :201-203) */
    double A = 0.0;
    vec2 c = V(0.0, 0.0);
    for (int64_t k = 0; k < p->edges.last; k++) {
        edge_t e = p->edges.data[k];
        double dA = 0.5 * fabs(cross2(vsub(e.v1, p->x), vsub(e.v2, p->x)));
        A += dA;
        vec2 s = vadd(vadd(p->x, e.v1), e.v2);
        c = vadd(c, V((dA * s.x) / 3.0, (dA * s.y) / 3.0));
    }
    return V(c.x / A, c.y / A);
}

static inline double lr_ratio(vec2 dx, edge_t e) { /* polygon.jl:228-232 */
    double l2 = norm_squared(vsub(e.v1, e.v2));
    double r2 = norm_squared(dx);
    return sqrt(l2 / r2);
}

static inline vec2 normal_vector(edge_t e) { /* polygon.jl:153-156 */
    vec2 n = V(e.v1.y - e.v2.y, e.v2.x - e.v1.x);
    double nn = vnorm(n);
    return V(n.x / nn, n.y / nn);
}

/* ------------------------------------------------------------------ voronoigrid.jl */
static inline double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

static inline vec2 get_arrow(const lvo_grid *g, vec2 x, vec2 y) { /* voronoigrid.jl:116-125 */
    vec2 v = vsub(x, y);
    if (g->xperiodic && (fabs(v.x) > 0.5 * g->xperiod)) {
        double s = jl_sign(v.x) * g->xperiod;
        v = V(v.x - s * 1.0, v.y - s * 0.0);
    }
    if (g->yperiodic && (fabs(v.y) > 0.5 * g->yperiod)) {
        double s = jl_sign(v.y) * g->yperiod;
        v = V(v.x - s * 0.0, v.y - s * 1.0);
    }
    return v;
}

static void insert_periodic(lvo_grid *g, vec2 x, int64_t label, int *bad) { /* voronoigrid.jl:130-147 */
    double px = g->xperiod, py = g->yperiod;
    cl_insert(g, x, label, bad);
    if (g->xperiodic) {
        cl_insert(g, V(x.x + px * 1.0, x.y + px * 0.0), label, bad);
        cl_insert(g, V(x.x - px * 1.0, x.y - px * 0.0), label, bad);
    }
    if (g->yperiodic) {
        cl_insert(g, V(x.x + py * 0.0, x.y + py * 1.0), label, bad);
        cl_insert(g, V(x.x - py * 0.0, x.y - py * 1.0), label, bad);
    }
    if (g->xperiodic && g->yperiodic) {
        cl_insert(g, V((x.x + px * 1.0) + py * 0.0, (x.y + px * 0.0) + py * 1.0), label, bad);
        cl_insert(g, V((x.x + px * 1.0) - py * 0.0, (x.y + px * 0.0) - py * 1.0), label, bad);
        cl_insert(g, V((x.x - px * 1.0) + py * 0.0, (x.y - px * 0.0) + py * 1.0), label, bad);
        cl_insert(g, V((x.x - px * 1.0) - py * 0.0, (x.y - px * 0.0) - py * 1.0), label, bad);
    }
}

/* returns 0 ok, LVO_EDESTROYED when the reference would throw (voronoigrid.jl:63-65) */
static int voronoicut_grid(const lvo_grid *g, poly_t *poly) { /* voronoigrid.jl:53-81 */
    vec2 x = poly->x;
    double prr = influence_rr(poly);
    int64_t k1, k2;
    if (!findkey(g, x, &k1, &k2)) return LVO_ENAN;
    for (int64_t t = 0; t < g->npath; t++) {
        double rr = g->path[t].rr;
        if (rr > prr) break;
        if (rr > g->rr_max) return LVO_EDESTROYED;
        int64_t c1 = k1 + g->path[t].i1, c2 = k2 + g->path[t].i2;
        if (!inbounds(g, c1, c2)) continue;
        const fv_int *cell = &CELL(g, c1, c2);
        for (int64_t s = 0; s < cell->last; s++) {
            int64_t i = cell->data[s];
            const poly_t *q = g->polygons[i - 1];
            vec2 y = vadd(x, get_arrow(g, q->x, x));
            if (veq(x, y) || (norm_squared(vsub(x, y)) > prr)) continue;
            if (voronoicut_poly(poly, y, i)) prr = influence_rr(poly);
        }
    }
    return LVO_OK;
}

static void sort_edges(poly_t *p) { /* IO.jl:35-48 */
    int64_t n = p->edges.last;
    edge_t *E = p->edges.data;
    for (int64_t i = 1; i <= n; i++) {
        vec2 last_vert = E[i - 1].v2;
        for (int64_t j = i + 1; j <= n; j++) {
            if (veq(last_vert, E[j - 1].v1)) {
                edge_t tmp = E[i];
                E[i] = E[j - 1];
                E[j - 1] = tmp;
                break;
            }
        }
    }
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

int lvo_remesh(lvo_grid *g) { /* voronoigrid.jl:89-108 */
    int64_t nc = g->n1 * g->n2;
    int status = LVO_OK;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nc; c++) g->cells[c].last = 0; /* :91-93 */
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < g->n; i++) { /* :96-100 */
        poly_t *poly = g->polygons[i];
        reset_poly(poly, g->cmin, g->cmax);
        int b = 0;
        insert_periodic(g, poly->x, i + 1, &b);
        bad |= b;
    }
    if (bad) return LVO_ENAN;
    /* restore `julia -t 1` bucket order (ascending label) regardless of thread count */
    if (omp_get_max_threads() > 1) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; c++)
            if (g->cells[c].last > 1) {
                fv_int *cell = &g->cells[c];
                /* insertion sort: buckets hold a handful of labels */
                if (cell->last <= 16) {
                    for (int64_t a = 1; a < cell->last; a++) {
                        int64_t v = cell->data[a], b2 = a - 1;
                        while (b2 >= 0 && cell->data[b2] > v) { cell->data[b2 + 1] = cell->data[b2]; b2--; }
                        cell->data[b2 + 1] = v;
                    }
                } else qsort(cell->data, (size_t)cell->last, sizeof(int64_t), cmp_i64);
            }
    }
#pragma omp parallel for schedule(static) reduction(max : status)
    for (int64_t i = 0; i < g->n; i++) { /* :102-105 */
        int st = voronoicut_grid(g, g->polygons[i]);
        sort_edges(g->polygons[i]);
        if (st > status) status = st;
    }
    return status;
}

int64_t lvo_nnz(const lvo_grid *g) {
    int64_t s = 0;
    for (int64_t i = 0; i < g->n; i++) s += g->polygons[i]->edges.last;
    return s;
}
void lvo_get_mesh(const lvo_grid *g, int64_t *rowptr, lvo_edge *edges) {
    int64_t s = 0;
    for (int64_t i = 0; i < g->n; i++) {
        rowptr[i] = s;
        const fv_edge *E = &g->polygons[i]->edges;
        for (int64_t k = 0; k < E->last; k++) {
            lvo_edge *o = &edges[s + k];
            o->v1[0] = E->data[k].v1.x; o->v1[1] = E->data[k].v1.y;
            o->v2[0] = E->data[k].v2.x; o->v2[1] = E->data[k].v2.y;
            o->label = E->data[k].label;
        }
        s += E->last;
    }
    rowptr[g->n] = s;
}
void lvo_area(const lvo_grid *g, double *area) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) area[i] = poly_area(g->polygons[i]);
}
void lvo_centroid(const lvo_grid *g, double *cxy) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        vec2 c = poly_centroid(g->polygons[i]);
        cxy[2 * i] = c.x; cxy[2 * i + 1] = c.y;
    }
}

/* ------------------------------------------------------------------ iterators.jl */
/* neighbors(p, grid): (q, e, y) for edges with label > 0, y = p.x + get_arrow(q.x, p.x) (iterators.jl:23-33) */
#define FOR_NEIGHBORS(g, p, q, e, y)                                                     \
    for (int64_t _k = 0; _k < (p)->edges.last; _k++)                                     \
        if ((p)->edges.data[_k].label > 0)                                               \
            for (int _o1 = 1; _o1;)                                                      \
                for (const edge_t e = (p)->edges.data[_k]; _o1;)                         \
                    for (const poly_t *q = (g)->polygons[e.label - 1]; _o1;)             \
                        for (const vec2 y = vadd((p)->x, get_arrow((g), q->x, (p)->x)); _o1; _o1 = 0)

/* ------------------------------------------------------------------ pressure.jl */
static void ensure_pressure(lvo_grid *g) { /* pressure.jl:94-100, 150-158 */
    if (g->A_n == g->n && g->A_nb) return;
    free_pressure(g);
    int64_t n = g->n;
    g->A_n = n;
    g->A_nb = (fv_int *)malloc(sizeof(fv_int) * (size_t)(n > 0 ? n : 1));
    g->A_lrr = (fv_f64 *)malloc(sizeof(fv_f64) * (size_t)(n > 0 ? n : 1));
    g->A_diag = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    g->b = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    g->Psol = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    g->GP = (vec2 *)calloc((size_t)(n > 0 ? n : 1), sizeof(vec2));
    for (int64_t i = 0; i < n; i++) {
        fv_int_init(&g->A_nb[i], POLYGON_SIZEHINT);
        fv_f64_init(&g->A_lrr[i], POLYGON_SIZEHINT);
        g->A_diag[i] = 1.0;
    }
}

int lvo_pressure_assemble(lvo_grid *g, double dt) { /* pressure.jl:104-117 */
    ensure_pressure(g);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->A_n; i++) {
        const poly_t *p = g->polygons[i];
        g->A_nb[i].last = 0;
        g->A_lrr[i].last = 0;
        g->A_diag[i] = p->mass / (((p->rho * p->rho) * p->c2) * (dt * dt));
        FOR_NEIGHBORS(g, p, q, e, y) {
            fv_int_push(&g->A_nb[i], e.label);
            fv_f64_push(&g->A_lrr[i], lr_ratio(vsub(p->x, y), e) * (0.5 / p->rho + 0.5 / q->rho));
        }
    }
    return LVO_OK;
}

void lvo_pressure_get_operator(const lvo_grid *g, int64_t *rowptr, int64_t *col, double *w, double *diag) {
    int64_t s = 0;
    for (int64_t i = 0; i < g->A_n; i++) {
        rowptr[i] = s;
        for (int64_t k = 0; k < g->A_nb[i].last; k++) {
            if (col) col[s + k] = g->A_nb[i].data[k];
            if (w) w[s + k] = g->A_lrr[i].data[k];
        }
        s += g->A_nb[i].last;
        if (diag) diag[i] = g->A_diag[i];
    }
    rowptr[g->A_n] = s;
}

void lvo_pressure_matvec(const lvo_grid *g, const double *x, double *y) { /* pressure.jl:119-130 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->A_n; i++) {
        double yi = g->A_diag[i] * x[i];
        const fv_int *nb = &g->A_nb[i];
        const fv_f64 *lr = &g->A_lrr[i];
        for (int64_t k = 0; k < nb->last; k++) {
            int64_t j = nb->data[k];
            double lrr = lr->data[k];
            yi += lrr * (x[i] - x[j - 1]);
        }
        y[i] = yi;
    }
}

static const double ZERO_VBC[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* pressure.jl:205-207 */

/* boundaries(p) (iterators.jl:50-57) over the whole grid, polygon by polygon, edges in storage order: first[i] = number of
 * the first boundary edge of polygon i (first has n+1 entries); returns the total */
static int64_t bdry_first(const lvo_grid *g, int64_t *first) {
    int64_t tot = 0;
    for (int64_t i = 0; i < g->n; i++) {
        first[i] = tot;
        const poly_t *p = g->polygons[i];
        for (int64_t k = 0; k < p->edges.last; k++) tot += (p->edges.data[k].label <= 0);
    }
    first[g->n] = tot;
    return tot;
}
int64_t lvo_boundary_edges(const lvo_grid *g, double *mid, int64_t *label, int64_t *polygon) {
    int64_t o = 0;
    for (int64_t i = 0; i < g->n; i++) {
        const poly_t *p = g->polygons[i];
        for (int64_t k = 0; k < p->edges.last; k++) {
            edge_t e = p->edges.data[k];
            if (!(e.label <= 0)) continue;
            if (mid) { vec2 m = midpoint_e(e); mid[2 * o] = m.x; mid[2 * o + 1] = m.y; }
            if (label) label[o] = e.label;
            if (polygon) polygon[o] = i + 1;
            o++;
        }
    }
    return o;
}
void lvo_set_vbc_edge(lvo_grid *g, const double *v, int64_t n) { /* NULL: back to the per-wall constants */
    free(g->vbc_edge);
    g->vbc_edge = NULL;
    g->n_vbc_edge = 0;
    if (v && n > 0) {
        g->vbc_edge = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        memcpy(g->vbc_edge, v, sizeof(double) * 2 * (size_t)n);
        g->n_vbc_edge = n;
    }
}

static void rhs_refresh(lvo_grid *g, double dt, int gp_step, const double *vbc_wall) { /* pressure.jl:162-203 */
    double *b = g->b, *P = g->Psol;
    vec2 *GP = g->GP;
    if (!vbc_wall) vbc_wall = ZERO_VBC;
    int64_t *bfirst = NULL;
    if (g->vbc_edge) { bfirst = (int64_t *)malloc(sizeof(int64_t) * (size_t)(g->n + 1)); bdry_first(g, bfirst); }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        const poly_t *p = g->polygons[i];
        double A = poly_area(p);
        int64_t be = bfirst ? bfirst[i] : 0;
        double bi = (A * p->P) / ((p->rho * p->c2) * (dt * dt));
        P[i] = p->P;
        vec2 gp = V(0.0, 0.0);
        FOR_NEIGHBORS(g, p, q, e, y) {
            double lrr = lr_ratio(vsub(p->x, y), e);
            vec2 m = midpoint_e(e);
            bi -= (lrr / dt) * vdot(vsub(p->v, q->v), vsub(m, y));
            double s = lrr * (p->P - q->P);
            vec2 d = vsub(m, p->x);
            gp = vsub(gp, V(s * d.x, s * d.y));
        }
        for (int64_t k = 0; k < p->edges.last; k++) { /* boundaries(p): iterators.jl:50-57 */
            edge_t e = p->edges.data[k];
            if (!(e.label <= 0)) continue;
            vec2 dS = V(e.v1.y - e.v2.y, e.v2.x - e.v1.x);
            vec2 vbc = V(0.0, 0.0);
            if (e.label < 0 && e.label >= -4) vbc = V(vbc_wall[2 * (-e.label - 1)], vbc_wall[2 * (-e.label - 1) + 1]);
            if (bfirst) { vbc = V(g->vbc_edge[2 * be], g->vbc_edge[2 * be + 1]); be++; } /* boundary_velocity(midpoint(e), e.label) */
            bi -= vdot(dS, vsub(vbc, p->v)) / dt;
        }
        GP[i] = V(gp.x / p->mass, gp.y / p->mass);
        b[i] = bi;
    }
    free(bfirst);
    if (gp_step) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < g->n; i++) {
            const poly_t *p = g->polygons[i];
            double bi = b[i];
            FOR_NEIGHBORS(g, p, q, e, y) {
                (void)q;
                double lrr = lr_ratio(vsub(p->x, y), e);
                vec2 m = midpoint_e(e);
                vec2 z = midpoint_v(p->x, y);
                int64_t j = e.label;
                bi += lrr * vdot(vsub(GP[i], GP[j - 1]), vsub(m, z));
            }
            b[i] = bi;
        }
    }
}

void lvo_pressure_rhs(lvo_grid *g, double dt, int gp_step, const double *vbc_wall, double *b, double *P0, double *GP) {
    ensure_pressure(g);
    rhs_refresh(g, dt, gp_step, vbc_wall);
    if (b) memcpy(b, g->b, sizeof(double) * (size_t)g->n);
    if (P0) memcpy(P0, g->Psol, sizeof(double) * (size_t)g->n);
    if (GP) memcpy(GP, g->GP, sizeof(vec2) * (size_t)g->n);
}

/* ---- vector kernels.  threadedvec.jl:33-56 thread axpy!/copyto!/fill!/rmul!; dot/norm/axpby!
 * fall back to generic (serial/BLAS) code in the reference.  Here everything is threaded,
 * which can only flatter the CPU baseline. */
static double vec_dot(int64_t n, const double *a, const double *b) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
static void vec_axpy(int64_t n, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] += a * x[i];
}
static void vec_scal(int64_t n, double a, double *x) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) x[i] *= a;
}
static void vec_copy(int64_t n, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) y[i] = x[i];
}
static void vec_fill(int64_t n, double a, double *x) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) x[i] = a;
}

/* MINRES after Paige & Saunders (1975) in the formulation of Krylov.jl 0.9.8 `minres!`
 * (third-party, un-vendored: Manifest.toml:465-469; call site pressure.jl:219).  Restated
 * from the published algorithm -- iterate-level parity UNPINNED.  M = I, lambda = 0,
 * window = 5, etol = sqrt(eps), conlim = 1/sqrt(eps). */
typedef void (*lvo_matvec_fn)(const lvo_grid *g, const double *x, double *y);
static int minres_core(const lvo_grid *g, lvo_matvec_fn matvec, int64_t n, const double *b, double *x0, double rtol, double atol,
                       int itmax, int warm_start, int *solved_out);
int lvo_minres(const lvo_grid *g, const double *b, double *x0, double rtol, double atol, int itmax, int warm_start) {
    return minres_core(g, lvo_pressure_matvec, g->A_n, b, x0, rtol, atol, itmax, warm_start, NULL);
}
static int minres_core(const lvo_grid *g, lvo_matvec_fn matvec, int64_t n, const double *b, double *x0, double rtol, double atol,
                       int itmax, int warm_start, int *solved_out) {
    const double epsM = 2.220446049250313e-16;
    const double etol = sqrt(epsM), ctol = sqrt(epsM); /* conlim = 1/sqrt(eps) -> ctol = 1/conlim */
    enum { WINDOW = 5 };
    double *r1 = (double *)malloc(sizeof(double) * (size_t)n * 7);
    double *r2 = r1 + n, *v = r2 + n, *yv = v + n, *w1 = yv + n, *w2 = w1 + n, *x = w2 + n;
    double err_vec[WINDOW] = {0, 0, 0, 0, 0};
    if (solved_out) *solved_out = 1;
    if (warm_start) { /* r1 = b - A*dx */
        matvec(g, x0, r1);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) r1[i] = b[i] - r1[i];
    } else vec_copy(n, b, r1);
    vec_fill(n, 0.0, x);
    vec_copy(n, r1, r2);
    vec_copy(n, r1, v);
    double beta1 = vec_dot(n, r1, v);
    int iter = 0;
    if (beta1 == 0.0) { /* zero residual: solution is the warm start (or zero) */
        if (!warm_start) vec_fill(n, 0.0, x0);
        free(r1);
        return 0;
    }
    beta1 = sqrt(beta1);
    double beta = beta1, oldbeta = 0.0, deltabar = 0.0, eps_ = 0.0, phibar = beta1;
    double rhs1 = beta1, rhs2 = 0.0, gmax = 0.0, gmin = INFINITY, cs = -1.0, sn = 0.0;
    double ANorm2 = 0.0, ANorm = 0.0, Acond = 0.0, rNorm = beta1, xNorm = 0.0, xENorm2 = 0.0, err_lbnd = 0.0;
    vec_fill(n, 0.0, w1);
    vec_fill(n, 0.0, w2);
    double tol = atol + rtol * beta1;
    if (itmax == 0) itmax = (int)(2 * n);
    int solved = rNorm <= tol, tired = iter >= itmax, ill_cond = 0;
    double *w = w2;
    while (!(solved || tired || ill_cond)) {
        iter += 1;
        matvec(g, v, yv);
        vec_scal(n, 1.0 / beta, yv);
        if (iter >= 2) vec_axpy(n, -beta / oldbeta, r1, yv);
        double alpha = vec_dot(n, v, yv) / beta;
        vec_axpy(n, -alpha / beta, r2, yv);
        double delta = cs * deltabar + sn * alpha;
        if (iter == 1) w = w2;
        else {
            if (iter >= 3) vec_scal(n, -eps_, w1);
            w = w1;
            vec_axpy(n, -delta, w2, w);
        }
        vec_axpy(n, 1.0 / beta, v, w);
        vec_copy(n, r2, r1);
        vec_copy(n, yv, r2);
        vec_copy(n, r2, v);
        oldbeta = beta;
        beta = vec_dot(n, r2, v);
        if (beta < 0.0) break;
        beta = sqrt(beta);
        ANorm2 = ANorm2 + alpha * alpha + oldbeta * oldbeta + beta * beta;
        double gbar = sn * deltabar - cs * alpha;
        eps_ = sn * beta;
        deltabar = -cs * beta;
        double root = sqrt(gbar * gbar + deltabar * deltabar);
        double gamma = sqrt(gbar * gbar + beta * beta);
        gamma = gamma > epsM ? gamma : epsM;
        cs = gbar / gamma;
        sn = beta / gamma;
        double phi = cs * phibar;
        phibar = sn * phibar;
        vec_scal(n, 1.0 / gamma, w);
        vec_axpy(n, phi, w, x);
        if (iter >= 2) { double *t = w1; w1 = w2; w2 = t; }
        err_vec[iter % WINDOW] = phi;
        if (iter >= WINDOW) {
            double s = 0.0;
            for (int k = 0; k < WINDOW; k++) s += err_vec[k] * err_vec[k];
            err_lbnd = sqrt(s);
        }
        gmax = gmax > gamma ? gmax : gamma;
        gmin = gmin < gamma ? gmin : gamma;
        double zeta = rhs1 / gamma;
        rhs1 = rhs2 - delta * zeta;
        rhs2 = -eps_ * zeta;
        ANorm = sqrt(ANorm2);
        xNorm = sqrt(vec_dot(n, x, x));
        Acond = gmax / gmin;
        rNorm = phibar;
        double test1 = rNorm / (ANorm * xNorm);
        double test2 = root / ANorm;
        xENorm2 = xENorm2 + phi * phi;
        int ill_cond_mach = (1.0 + 1.0 / Acond <= 1.0);
        int solved_mach = (1.0 + test2 <= 1.0);
        int zero_resid_mach = (1.0 + test1 <= 1.0);
        int resid_decrease_mach = (rNorm + 1.0 <= 1.0);
        tired = iter >= itmax;
        int ill_cond_lim = (1.0 / Acond <= ctol);
        int solved_lim = (test2 <= tol);
        int fwd_err = (iter >= WINDOW) && (err_lbnd <= etol * sqrt(xENorm2));
        int zero_resid_lim = (test1 <= tol);
        int resid_decrease_lim = (rNorm <= tol);
        int zero_resid = zero_resid_mach | zero_resid_lim;
        int resid_decrease = resid_decrease_mach | resid_decrease_lim;
        ill_cond = ill_cond_mach | ill_cond_lim;
        solved = solved_mach | solved_lim | zero_resid | fwd_err | resid_decrease;
    }
    if (warm_start) vec_axpy(n, 1.0, x0, x); /* x += dx */
    vec_copy(n, x, x0);
    if (solved_out) *solved_out = solved || ill_cond;
    free(r1);
    return iter;
}

/* plain conjugate gradients (Hestenes-Stiefel); stop on ||r|| <= atol + rtol*||r0|| (recursive r) */
int lvo_cg(const lvo_grid *g, const double *b, double *x, double rtol, double atol, int itmax) {
    int64_t n = g->A_n;
    double *r = (double *)malloc(sizeof(double) * (size_t)n * 3);
    double *p = r + n, *Ap = p + n;
    lvo_pressure_matvec(g, x, Ap);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) { r[i] = b[i] - Ap[i]; p[i] = r[i]; }
    double rr = vec_dot(n, r, r);
    double tol = atol + rtol * sqrt(rr);
    int iter = 0;
    while (sqrt(rr) > tol && iter < itmax) {
        lvo_pressure_matvec(g, p, Ap);
        double pAp = vec_dot(n, p, Ap);
        double alpha = rr / pAp;
        vec_axpy(n, alpha, p, x);
        vec_axpy(n, -alpha, Ap, r);
        double rr_new = vec_dot(n, r, r);
        double beta = rr_new / rr;
        rr = rr_new;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
        iter++;
    }
    free(r);
    return iter;
}

int lvo_find_pressure(lvo_grid *g, double dt, int niter, double rtol, double atol, int itmax, int solver,
                      const double *vbc_wall, int32_t *iters_out, double *relres_out) { /* pressure.jl:215-225 */
    lvo_pressure_assemble(g, dt);
    int64_t n = g->n;
    double *tmp = relres_out ? (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1)) : NULL;
    for (int it = 1; it <= niter; it++) {
        rhs_refresh(g, dt, it > 1, vbc_wall);
        int iters = solver == 0 ? lvo_minres(g, g->b, g->Psol, rtol, atol, itmax, 1)
                                : lvo_cg(g, g->b, g->Psol, rtol, atol, itmax);
        if (iters_out) iters_out[it - 1] = iters;
        if (relres_out) {
            lvo_pressure_matvec(g, g->Psol, tmp);
            double rn = 0.0, bn = 0.0;
            for (int64_t i = 0; i < n; i++) { double d = g->b[i] - tmp[i]; rn += d * d; bn += g->b[i] * g->b[i]; }
            relres_out[it - 1] = bn > 0.0 ? sqrt(rn / bn) : sqrt(rn);
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) g->polygons[i]->P = g->Psol[i]; /* :221-223 */
    }
    free(tmp);
    return LVO_OK;
}

/* ------------------------------------------------------------------ populate.jl:149-174 */
int lvo_populate_hex(lvo_grid *g) {
    double a = pow(4.0 / 3.0, 0.25) * g->dr;
    double b = pow(3.0 / 4.0, 0.25) * g->dr;
    int64_t i_min = (int64_t)floor(g->bmin.x / a) - 1, j_min = (int64_t)floor(g->bmin.y / b);
    int64_t i_max = (int64_t)ceil(g->bmax.x / a), j_max = (int64_t)ceil(g->bmax.y / b);
    int64_t cap = (i_max - i_min + 1) * (j_max - j_min + 1);
    double *xy = (double *)malloc(sizeof(double) * 2 * (size_t)(cap > 0 ? cap : 1));
    int64_t n = 0;
    for (int64_t i = i_min; i <= i_max; i++)
        for (int64_t j = j_min; j <= j_max; j++) {
            double x1 = ((double)i + (double)(j % 2) / 2.0) * a;
            double x2 = (double)j * b;
            if ((g->bmin.x <= x1 && x1 <= g->bmax.x) && (g->bmin.y <= x2 && x2 <= g->bmax.y)) { /* geometry.jl:127-129 */
                xy[2 * n] = x1; xy[2 * n + 1] = x2; n++;
            }
        }
    lvo_set_points(g, n, xy);
    free(xy);
    return lvo_remesh(g);
}

/* ------------------------------------------------------------------ populate.jl:18-145 (the other seeding strategies)
 * charfun = everywhere; ic! is the caller's business.  Julia's ranges are evaluated in twice precision (StepRangeLen),
 * i.e. element k of a:s:b is the correctly rounded a + k*s: long double arithmetic reproduces that on x86. */
static int inside_rect(const lvo_grid *g, double x1, double x2) { /* isinside  geometry.jl:127-129 */
    return (g->bmin.x <= x1 && x1 <= g->bmax.x) && (g->bmin.y <= x2 && x2 <= g->bmax.y);
}
static double corner_rmax(const lvo_grid *g, double cx, double cy) { /* maximum(norm(x - center) for x in verts(rect)) */
    double r = 0.0;
    const double xs[2] = {g->bmin.x, g->bmax.x}, ys[2] = {g->bmin.y, g->bmax.y};
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            double dx = xs[a] - cx, dy = ys[b] - cy, d = sqrt(dx * dx + dy * dy);
            if (d > r) r = d;
        }
    return r;
}
typedef struct { double *xy; int64_t n, cap; } ptbuf;
static void pt_push(ptbuf *b, double x1, double x2) {
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 1024; b->xy = (double *)realloc(b->xy, sizeof(double) * 2 * (size_t)b->cap); }
    b->xy[2 * b->n] = x1; b->xy[2 * b->n + 1] = x2; b->n++;
}
int lvo_populate_circ(lvo_grid *g, double cx, double cy) { /* populate.jl:18-35 */
    double r_max = corner_rmax(g, cx, cy);
    ptbuf b = {NULL, 0, 0};
    /* for r in (0.5*dr):dr:r_max -- floor((r_max - 0.5dr)/dr) + 1 elements, the end point included when it is hit */
    int64_t nring = r_max >= 0.5 * g->dr ? (int64_t)floor((r_max - 0.5 * g->dr) / g->dr * (1.0 + 4e-16)) + 1 : 0;
    for (int64_t k = 0; k < nring; k++) {
        double r = (double)(0.5L * (long double)g->dr + (long double)k * (long double)g->dr);
        int64_t k_max = (int64_t)nearbyint(2.0 * M_PI * r / g->dr); /* round(Int, .): ties to even, like Julia */
        for (int64_t q = 1; q <= k_max; q++) {
            double theta = 2.0 * M_PI * (double)q / (double)k_max;
            double x1 = cx + r * cos(theta), x2 = cy + r * sin(theta);
            if (inside_rect(g, x1, x2)) pt_push(&b, x1, x2);
        }
    }
    lvo_set_points(g, b.n, b.xy);
    free(b.xy);
    return lvo_remesh(g);
}
int lvo_populate_rect(lvo_grid *g) { /* populate.jl:46-66 */
    int64_t N = (int64_t)nearbyint((g->bmax.x - g->bmin.x) / g->dr), M = (int64_t)nearbyint((g->bmax.y - g->bmin.y) / g->dr);
    ptbuf b = {NULL, 0, 0};
    for (int64_t i = 0; i < N; i++) {     /* range(x1_min, x1_max, N): N points, both ends included */
        long double t1 = N > 1 ? (long double)i / (long double)(N - 1) : 0.0L;
        double x1 = (double)((long double)g->bmin.x + t1 * ((long double)g->bmax.x - (long double)g->bmin.x));
        for (int64_t j = 0; j < M; j++) {
            long double t2 = M > 1 ? (long double)j / (long double)(M - 1) : 0.0L;
            double x2 = (double)((long double)g->bmin.y + t2 * ((long double)g->bmax.y - (long double)g->bmin.y));
            double p1 = x1 + 0.5 * g->dr, p2 = x2 + 0.5 * g->dr;
            if (inside_rect(g, p1, p2)) pt_push(&b, p1, p2);
        }
    }
    lvo_set_points(g, b.n, b.xy);
    free(b.xy);
    return lvo_remesh(g);
}
int lvo_populate_vogel(lvo_grid *g, double cx, double cy) { /* populate.jl:105-121 */
    double r_max = corner_rmax(g, cx, cy);
    int64_t N = (int64_t)nearbyint(M_PI * r_max * r_max / (g->dr * g->dr));
    ptbuf b = {NULL, 0, 0};
    for (int64_t i = 1; i <= N; i++) {
        double r = r_max * sqrt((double)i / (double)N);
        double theta = 2.39996322972865332 * (double)i;
        double x1 = cx + r * cos(theta), x2 = cy + r * sin(theta);
        if (inside_rect(g, x1, x2)) pt_push(&b, x1, x2);
    }
    lvo_set_points(g, b.n, b.xy);
    free(b.xy);
    return lvo_remesh(g);
}
/* populate_rand!  populate.jl:76-95 with the uniform samples (s1, s2 per point) supplied by the caller instead of Julia's
 * global RNG; ns = number of sample pairs available (>= N = round(area/dr^2)) */
int lvo_populate_rand(lvo_grid *g, const double *s, int64_t ns) {
    int64_t N = (int64_t)nearbyint(fabs(g->bmax.x - g->bmin.x) * fabs(g->bmax.y - g->bmin.y) / (g->dr * g->dr));
    if (N > ns) return LVO_EINVAL;
    ptbuf b = {NULL, 0, 0};
    for (int64_t i = 0; i < N; i++) {
        double s1 = s[2 * i], s2 = s[2 * i + 1];
        double x1 = s1 * g->bmax.x + (1 - s1) * g->bmin.x, x2 = s2 * g->bmax.y + (1 - s2) * g->bmin.y;
        if (inside_rect(g, x1, x2)) pt_push(&b, x1, x2);
    }
    lvo_set_points(g, b.n, b.xy);
    free(b.xy);
    return lvo_remesh(g);
}
/* the relaxation loop of populate_lloyd!  populate.jl:132-145: niter x (remesh!; p.x = centroid(p)), then remesh! */
int lvo_lloyd(lvo_grid *g, int niter) {
    for (int it = 0; it < niter; it++) {
        int st = lvo_remesh(g);
        if (st != LVO_OK) return st;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < g->n; i++) g->polygons[i]->x = poly_centroid(g->polygons[i]);
    }
    return lvo_remesh(g);
}

/* ------------------------------------------------------------------ move.jl */
static double least_positive_residue(double x, double d) { return fmod(fmod(x, d) + d, d); } /* voronoigrid.jl:181-183 */
static vec2 periodic_wrap(const lvo_grid *g, vec2 x) { /* voronoigrid.jl:187-192 */
    double _x = least_positive_residue(x.x - g->bmin.x, g->xperiod) + g->bmin.x;
    double _y = least_positive_residue(x.y - g->bmin.y, g->yperiod) + g->bmin.y;
    /* Bool*Float64: false is a strong zero */
    double d1 = g->xperiodic ? (_x - x.x) : 0.0;
    double d2 = g->yperiodic ? (_y - x.y) : 0.0;
    return V((x.x + d1 * 1.0) + d2 * 0.0, (x.y + d1 * 0.0) + d2 * 1.0);
}
static int try_move(const lvo_grid *g, poly_t *p, double dt, int *nanflag) { /* move.jl:23-33 */
    if (isnan(p->v.x) || isnan(p->v.y)) { *nanflag = 1; return 1; }
    vec2 _x = periodic_wrap(g, V(p->x.x + dt * p->v.x, p->x.y + dt * p->v.y));
    if ((g->bmin.x <= _x.x && _x.x <= g->bmax.x) && (g->bmin.y <= _x.y && _x.y <= g->bmax.y)) {
        p->x = _x;
        return 1;
    }
    return 0;
}
int lvo_move(lvo_grid *g, double dt) { /* move.jl:9-21 */
    int nanflag = 0;
#pragma omp parallel for schedule(static) reduction(| : nanflag)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        int nf = 0;
        if (try_move(g, p, dt, &nf)) { nanflag |= nf; continue; }
        for (int64_t k = 0; k < p->edges.last; k++) {
            edge_t e = p->edges.data[k];
            if (!(e.label <= 0)) continue;
            vec2 nv = normal_vector(e);
            double d = vdot(p->v, nv);
            p->v = V(p->v.x - d * nv.x, p->v.y - d * nv.y);
        }
        if (try_move(g, p, dt, &nf)) { nanflag |= nf; continue; }
        p->v = V(0.0, 0.0);
        try_move(g, p, dt, &nf);
        nanflag |= nf;
    }
    if (nanflag) return LVO_ENAN; /* "Velocity field invalidated." move.jl:24-26 */
    return lvo_remesh(g);
}

/* ------------------------------------------------------------------ pressure.jl:10-82 */
static inline double eint(const poly_t *p) { return p->e - 0.5 * norm_squared(p->v); } /* :32-34 */
void lvo_stiffened_eos(lvo_grid *g, double gamma, double P0) { /* :64-70 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        p->rho = p->mass / poly_area(p);
        p->P = ((gamma - 1.0) * p->rho) * eint(p);
        p->c2 = (gamma * (p->P + P0)) / p->rho;
    }
}
void lvo_ideal_eos(lvo_grid *g, double gamma, double Pmin) { /* :49-55 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        p->rho = p->mass / poly_area(p);
        p->P = ((gamma - 1.0) * p->rho) * eint(p);
        double pm = p->P > Pmin ? p->P : Pmin;
        p->c2 = (gamma * pm) / p->rho;
    }
}
void lvo_pressure_step(lvo_grid *g, double dt) { /* :10-25 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        FOR_NEIGHBORS(g, p, q, e, y) {
            double lrr = lr_ratio(vsub(p->x, y), e);
            vec2 m = midpoint_e(e);
            double s = ((dt / p->mass) * lrr) * (p->P - q->P);
            vec2 d = vsub(m, p->x);
            p->v = V(p->v.x + s * d.x, p->v.y + s * d.y);
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        FOR_NEIGHBORS(g, p, q, e, y) {
            double lrr = lr_ratio(vsub(p->x, y), e);
            vec2 m = midpoint_e(e);
            double a = vdot(vsub(m, p->x), V(p->P * p->v.x, p->P * p->v.y));
            double b = vdot(vsub(m, y), V(q->P * q->v.x, q->P * q->v.y));
            p->e -= ((dt * lrr) / p->mass) * (a - b);
        }
    }
}

/* ------------------------------------------------------------------ diffusion.jl */
void lvo_find_D(lvo_grid *g) { /* diffusion.jl:8-19 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        double D[4] = {0, 0, 0, 0};
        FOR_NEIGHBORS(g, p, q, e, y) {
            vec2 m = midpoint_e(e);
            double lrr = lr_ratio(vsub(p->x, y), e);
            vec2 a = vsub(p->v, q->v), b = vsub(m, y);
            /* outer(x,y) = (x1*y1, x2*y1, x1*y2, x2*y2) geometry.jl:186-188 */
            D[0] += lrr * (a.x * b.x); D[1] += lrr * (a.y * b.x);
            D[2] += lrr * (a.x * b.y); D[3] += lrr * (a.y * b.y);
        }
        double A = poly_area(p);
        for (int k = 0; k < 4; k++) D[k] /= A;
        /* 0.5*(D + D') */
        p->D[0] = 0.5 * (D[0] + D[0]);
        p->D[1] = 0.5 * (D[1] + D[2]);
        p->D[2] = 0.5 * (D[2] + D[1]);
        p->D[3] = 0.5 * (D[3] + D[3]);
    }
}
static void getS(const poly_t *p, double dr, double S[4]) { /* diffusion.jl:22-29 */
    double divv = ((p->D[0] * 1.0 + p->D[1] * 0.0) + p->D[2] * 0.0) + p->D[3] * 1.0;
    double mu = p->mu;
    if (divv < 0.0) mu -= (divv * p->rho) * (dr * dr);
    double t = 2.0 * mu;
    S[0] = t * (p->D[0] - (divv * 1.0) / 3.0);
    S[1] = t * (p->D[1] - (divv * 0.0) / 3.0);
    S[2] = t * (p->D[2] - (divv * 0.0) / 3.0);
    S[3] = t * (p->D[3] - (divv * 1.0) / 3.0);
}
static inline vec2 matvec2(const double M[4], vec2 v) { return V(M[0] * v.x + M[2] * v.y, M[1] * v.x + M[3] * v.y); }
void lvo_viscous_step(lvo_grid *g, double dt, int artificial_viscosity) { /* diffusion.jl:39-53 */
    double avdr = artificial_viscosity ? g->dr : 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        FOR_NEIGHBORS(g, p, q, e, y) {
            vec2 m = midpoint_e(e);
            double Sp[4], Sq[4], M[4];
            getS(p, avdr, Sp); getS(q, avdr, Sq);
            double s = (dt * lr_ratio(vsub(p->x, y), e)) / p->mass;
            for (int k = 0; k < 4; k++) M[k] = s * (Sp[k] - Sq[k]);
            vec2 d = matvec2(M, vsub(m, p->x));
            p->v = vsub(p->v, d);
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        FOR_NEIGHBORS(g, p, q, e, y) {
            vec2 m = midpoint_e(e);
            double Sp[4], Sq[4];
            getS(p, avdr, Sp); getS(q, avdr, Sq);
            double s = (dt * lr_ratio(vsub(p->x, y), e)) / p->mass;
            double a = vdot(vsub(m, p->x), matvec2(Sp, p->v));
            double b = vdot(vsub(m, y), matvec2(Sq, q->v));
            p->e += s * (a - b);
        }
    }
}

/* diffusion.jl:55-80 bdary_friction!: viscous drag of the walls (Dirichlet condition for a tangential wall velocity).
 * vDirichlet is a closure in the reference; here it is the per-wall constant vwall[-label-1] (what examples/cavity.jl:41-44
 * evaluates to), charfun = everywhere. */
void lvo_bdary_friction_ex(lvo_grid *g, double dt, const double *vwall, const unsigned char *wall_on, const double *v_edge,
                           const unsigned char *on_edge);
void lvo_bdary_friction(lvo_grid *g, double dt, const double *vwall) { lvo_bdary_friction_ex(g, dt, vwall, NULL, NULL, NULL); }
/* general form: v_edge[2e], on_edge[e] = vDirichlet(m), charfun(m) evaluated by the caller at the midpoint of boundary edge
 * number e (numbering of lvo_boundary_edges); wall_on[4] switches whole walls off */
void lvo_bdary_friction_ex(lvo_grid *g, double dt, const double *vwall, const unsigned char *wall_on, const double *v_edge,
                           const unsigned char *on_edge) {
    int64_t *bfirst = NULL;
    if (v_edge || on_edge) { bfirst = (int64_t *)malloc(sizeof(int64_t) * (size_t)(g->n + 1)); bdry_first(g, bfirst); }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        double tmp = 1.0;
        int64_t be = bfirst ? bfirst[i] : 0;
        for (int64_t k = 0; k < p->edges.last; k++) { /* boundaries(p): iterators.jl:50-57 */
            edge_t e = p->edges.data[k];
            if (!(e.label <= 0)) continue;
            const int64_t e_no = be++;
            if (wall_on && e.label < 0 && e.label >= -4 && !wall_on[-e.label - 1]) continue; /* if !charfun(m) continue end */
            if (on_edge && !on_edge[e_no]) continue;
            vec2 m = vscale(0.5, vadd(e.v1, e.v2));
            vec2 nv = normal_vector(e);
            double lrr = len_e(e) / fabs(vdot(vsub(m, p->x), nv));
            vec2 vd = V(0.0, 0.0);
            if (vwall && e.label < 0 && e.label >= -4) vd = V(vwall[2 * (-e.label - 1)], vwall[2 * (-e.label - 1) + 1]);
            if (v_edge) vd = V(v_edge[2 * e_no], v_edge[2 * e_no + 1]);
            double c = p->mu * lrr;
            vec2 f = V((c * vd.x) / p->mass, (c * vd.y) / p->mass);
            tmp += ((dt * p->mu) * lrr) / p->mass;
            p->e += dt * vdot(f, p->v);
            p->v = V(p->v.x + dt * f.x, p->v.y + dt * f.y);
        }
        p->v = V(p->v.x / tmp, p->v.y / tmp);
    }
    free(bfirst);
}

/* ------------------------------------------------------------------ relaxation.jl:10-73 */
void lvo_find_dv(lvo_grid *g, double dt, double alpha) { /* relaxation.jl:10-25 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        p->dv = V(0.0, 0.0);
        vec2 c = poly_centroid(p);
        double rmax = 0.0, rmin = INFINITY;
        FOR_NEIGHBORS(g, p, q, e, y) {
            (void)q; (void)e;
            double r = vnorm(vsub(p->x, y));
            rmax = rmax > r ? rmax : r;
            rmin = rmin < r ? rmin : r;
        }
        p->quality = rmin / rmax;
        double nD = sqrt(((p->D[0] * p->D[0] + p->D[1] * p->D[1]) + p->D[2] * p->D[2]) + p->D[3] * p->D[3]);
        double lambda = (alpha * nD) / (p->quality * p->quality);
        double s = lambda / (1.0 + dt * lambda);
        p->dv = V(s * (c.x - p->x.x), s * (c.y - p->x.y));
    }
}
int lvo_relaxation_step(lvo_grid *g, double dt, int rusanov) { /* relaxation.jl:36-73 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        p->momentum = V(p->mass * p->v.x, p->mass * p->v.y);
        p->energy = p->mass * p->e;
        FOR_NEIGHBORS(g, p, q, e, y) {
            if (!(p->phase == q->phase)) continue;
            double lrr = lr_ratio(vsub(p->x, y), e);
            vec2 m = midpoint_e(e);
            vec2 z = midpoint_v(p->x, y);
            vec2 pq = vsub(p->x, y), mz = vsub(m, z);
            double pdvpq = vdot(p->dv, pq), qdvpq = vdot(q->dv, pq);
            double pdvmz = vdot(p->dv, mz), qdvmz = vdot(q->dv, mz);
            double c = dt * lrr;
            p->mass += c * ((pdvmz * p->rho - qdvmz * q->rho) - 0.5 * (pdvpq * p->rho + qdvpq * q->rho));
            double a1 = pdvmz * p->rho, a2 = qdvmz * q->rho, a3 = pdvpq * p->rho, a4 = qdvpq * q->rho;
            p->momentum.x += c * ((a1 * p->v.x - a2 * q->v.x) - 0.5 * (a3 * p->v.x + a4 * q->v.x));
            p->momentum.y += c * ((a1 * p->v.y - a2 * q->v.y) - 0.5 * (a3 * p->v.y + a4 * q->v.y));
            p->energy += c * ((a1 * p->e - a2 * q->e) - 0.5 * (a3 * p->e + a4 * q->e));
            if (rusanov) {
                double na = vnorm(p->dv), nb = vnorm(q->dv);
                double a = na > nb ? na : nb;
                double l = len_e(e);
                double k = ((0.5 * dt) * l) * a;
                p->mass += k * (q->rho - p->rho);
                p->momentum.x += k * (q->rho * q->v.x - p->rho * p->v.x);
                p->momentum.y += k * (q->rho * q->v.y - p->rho * p->v.y);
                p->energy += k * (q->rho * q->e - p->rho * p->e);
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        p->v = V(p->momentum.x / p->mass, p->momentum.y / p->mass);
        p->e = p->energy / p->mass;
        p->x = V(p->x.x + dt * p->dv.x, p->x.y + dt * p->dv.y);
    }
    return lvo_remesh(g);
}

/* ------------------------------------------------------------------ relaxation.jl:75-206 (multiphase projector) */
static void mp_pass2(const lvo_grid *g, const vec2 *t, double *res) { /* second sweep of mul! (:107-121) and refresh! (:162-177) */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        const poly_t *p = g->polygons[i];
        double r = 0.0;
        FOR_NEIGHBORS(g, p, q, e, y) {
            if (p->phase != q->phase) {
                double lrr = lr_ratio(vsub(p->x, y), e);
                vec2 m = midpoint_e(e), z = midpoint_v(p->x, y);
                int64_t j = e.label - 1;
                r -= lrr * (vdot(vsub(t[i], t[j]), vsub(m, z)) - 0.5 * vdot(vadd(t[i], t[j]), vsub(p->x, y)));
            }
        }
        res[i] = r;
    }
}
static vec2 *g_mp_tmp = NULL; /* tmp_vec of the MultiphaseProjector (:83) */
static void mp_matvec(const lvo_grid *g, const double *x, double *res) { /* relaxation.jl:91-123 */
    vec2 *tmp = g_mp_tmp;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        const poly_t *p = g->polygons[i];
        vec2 t = V(0.0, 0.0);
        FOR_NEIGHBORS(g, p, q, e, y) {
            if (p->phase != q->phase) {
                double lrr = lr_ratio(vsub(p->x, y), e);
                vec2 m = midpoint_e(e);
                int64_t j = e.label - 1;
                double s = lrr * (x[i] - x[j]);
                vec2 d = vsub(m, p->x);
                t = vsub(t, V(s * d.x, s * d.y));
            }
        }
        double A = poly_area(p);
        tmp[i] = V(t.x / A, t.y / A);
    }
    mp_pass2(g, tmp, res);
}
/* direct application of the projector and of its right-hand side (tests): y = A x (relaxation.jl:91-123), b from dv (:162-177) */
void lvo_multiphase_apply(lvo_grid *g, const double *x, double *y, double *b) {
    int64_t n = g->n;
    g_mp_tmp = (vec2 *)malloc(sizeof(vec2) * (size_t)(n > 0 ? n : 1));
    if (x && y) mp_matvec(g, x, y);
    if (b) {
        vec2 *dv = (vec2 *)malloc(sizeof(vec2) * (size_t)(n > 0 ? n : 1));
        for (int64_t i = 0; i < n; i++) dv[i] = g->polygons[i]->dv;
        mp_pass2(g, dv, b);
        free(dv);
    }
    free(g_mp_tmp); g_mp_tmp = NULL;
}
int lvo_multiphase_projection(lvo_grid *g, double quality_threshold, double rtol, double atol, int itmax, int *iters, int *solved) {
    int64_t n = g->n;
    double *b = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    double *res = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    vec2 *dv = (vec2 *)malloc(sizeof(vec2) * (size_t)(n > 0 ? n : 1));
    g_mp_tmp = (vec2 *)malloc(sizeof(vec2) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) dv[i] = g->polygons[i]->dv;
    mp_pass2(g, dv, b); /* refresh!  :162-177 */
    int ok = 1;
    int it = minres_core(g, mp_matvec, n, b, res, rtol, atol, itmax, 0, &ok); /* :182 */
    if (iters) *iters = it;
    if (solved) *solved = ok;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) { /* :191-203 */
        poly_t *p = g->polygons[i];
        if (p->quality < quality_threshold) continue;
        double A = poly_area(p);
        vec2 d = p->dv;
        FOR_NEIGHBORS(g, p, q, e, y) {
            if (p->phase != q->phase) {
                int64_t j = e.label - 1;
                vec2 m = midpoint_e(e);
                double s = lr_ratio(vsub(p->x, y), e) * (res[i] - res[j]);
                vec2 w = vsub(m, p->x);
                d = V(d.x + (s * w.x) / A, d.y + (s * w.y) / A);
            }
        }
        dv[i] = d;
    }
    for (int64_t i = 0; i < n; i++) g->polygons[i]->dv = dv[i];
    free(b); free(res); free(dv); free(g_mp_tmp); g_mp_tmp = NULL;
    return LVO_OK;
}
void lvo_gravity_step(lvo_grid *g, double gx, double gy, double dt) { /* pressure.jl:77-82 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < g->n; i++) {
        poly_t *p = g->polygons[i];
        p->v = V(p->v.x + dt * gx, p->v.y + dt * gy);
    }
}
