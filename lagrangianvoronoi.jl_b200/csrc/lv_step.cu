// lv_step.cu -- the callers either side of the hot path, on the device.
//
// With remesh! and find_pressure! on the GPU, a time step of the reference still gathers / scatters
// x, v, rho, mass, c2, P across PCIe around every call.  This file keeps the polygon fields of
// @Euler_vars (celldefs.jl:7-27) resident in HBM, in LABEL order (label order is stable across remeshes),
// and restates the explicit per-cell sweeps of a canonical step! so that a whole step never leaves the device:
//   move!              move.jl:9-33            stiffened_eos! / ideal_eos!   pressure.jl:49-70
//   pressure_step!     pressure.jl:10-25       gravity_step!                 pressure.jl:77-82
//   find_D!            diffusion.jl:8-19       viscous_step!                 diffusion.jl:39-53
//   find_dv!           relaxation.jl:10-25     relaxation_step!              relaxation.jl:36-73
//   bdary_friction!    diffusion.jl:64-80
// Every sweep is one thread per polygon walking its CSR row (slot = prim_of_label[label]); neighbour fields
// are read by label.  Expressions keep the reference's association (compiled with -fmad=false), so results
// agree with the CPU restatement to rounding of the sums' inputs, i.e. bit for bit in practice.
#include "lv_internal.cuh"
#include <cstring>

#define ST_BLOCK 128

struct StepView { // everything a sweep needs, passed by value
    LvGridParams g;
    int64_t n;
    const int *prim;
    const int *rowptr;
    const unsigned char *rdeg;
    const int *col;
    const double2 *v1, *v2;
    const unsigned *ent_label;
    const double *area;
    const double2 *cen;
    double2 *x, *v, *dv, *mom;
    double *rho, *e, *P, *c2, *mass, *energy, *quality, *mu, *phase;
    double *D; // 4 per polygon, column-major D11 D21 D12 D22
};

__device__ __forceinline__ double lr_ratio(double2 dx, double2 a, double2 b) { // polygon.jl:228-232
    const double ex = a.x - b.x, ey = a.y - b.y;
    return sqrt((ex * ex + ey * ey) / (dx.x * dx.x + dx.y * dx.y));
}

// iterate neighbors(p, grid) (iterators.jl:23-33): body gets q (label), the edge end points and y
#define FOR_NEIGHBORS(S, i, xi)                                                                   \
    const int _s = (S).prim[i];                                                                   \
    const int _r0 = _s >= 0 ? (S).rowptr[_s] : 0, _d = _s >= 0 ? (int)(S).rdeg[_s] : 0;           \
    for (int _k = _r0; _k < _r0 + _d; _k++)                                                       \
        if ((S).col[_k] >= 0)                                                                     \
            for (int _once = 1; _once;)                                                           \
                for (const int q = (int)((S).ent_label[(S).col[_k]] & ~LV_IMAGE_BIT); _once;)      \
                    for (const double2 ea = (S).v1[_k], eb = (S).v2[_k]; _once;)                  \
                        for (const double2 y = lv_neighbor_pos((S).g, xi, (S).x[q]); _once; _once = 0)

// ---- move!  move.jl:9-33 ---------------------------------------------------------------------------------
__device__ __forceinline__ double least_positive_residue(double x, double d) { return fmod(fmod(x, d) + d, d); } // voronoigrid.jl:181-183

__device__ __forceinline__ bool try_move(const StepView &S, double bminx, double bminy, double bmaxx, double bmaxy, double2 x, double2 v,
                                         double dt, double2 &out) { // move.jl:23-33 (NaN velocity is checked by the caller)
    const double px = x.x + dt * v.x, py = x.y + dt * v.y;
    // periodic_wrap  voronoigrid.jl:187-192
    const double wx = least_positive_residue(px - bminx, S.g.xperiod) + bminx;
    const double wy = least_positive_residue(py - bminy, S.g.yperiod) + bminy;
    const double d1 = S.g.xper ? (wx - px) : 0.0, d2 = S.g.yper ? (wy - py) : 0.0;
    const double nx = (px + d1 * 1.0) + d2 * 0.0, ny = (py + d1 * 0.0) + d2 * 1.0;
    out = make_double2(nx, ny);
    return (bminx <= nx && nx <= bmaxx) && (bminy <= ny && ny <= bmaxy); // isinside  geometry.jl:127-129
}

__global__ void __launch_bounds__(ST_BLOCK) k_move(StepView S, double dt, double bminx, double bminy, double bmaxx, double bmaxy, int *flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    double2 x = S.x[i], v = S.v[i], nx;
    if (isnan(v.x) || isnan(v.y)) { atomicOr(&flags[LVF_NAN], 1); return; } // "Velocity field invalidated."  move.jl:24-26
    if (try_move(S, bminx, bminy, bmaxx, bmaxy, x, v, dt, nx)) { S.x[i] = nx; return; }
    // project v to the tangent space of every wall edge of the polygon (move.jl:12-15)
    const int s = S.prim[i];
    if (s >= 0) {
        const int r0 = S.rowptr[s], d = S.rdeg[s];
        for (int k = r0; k < r0 + d; k++) {
            if (S.col[k] >= 0) continue;
            const double2 a = S.v1[k], b = S.v2[k];
            double nnx = a.y - b.y, nny = b.x - a.x; // normal_vector  polygon.jl:153-156
            const double nn = sqrt(nnx * nnx + nny * nny);
            nnx /= nn; nny /= nn;
            const double dn = v.x * nnx + v.y * nny;
            v = make_double2(v.x - dn * nnx, v.y - dn * nny);
        }
    }
    if (isnan(v.x) || isnan(v.y)) { atomicOr(&flags[LVF_NAN], 1); S.v[i] = v; return; }
    if (try_move(S, bminx, bminy, bmaxx, bmaxy, x, v, dt, nx)) { S.x[i] = nx; S.v[i] = v; return; }
    v = make_double2(0.0, 0.0);
    if (try_move(S, bminx, bminy, bmaxx, bmaxy, x, v, dt, nx)) S.x[i] = nx;
    S.v[i] = v;
}

// ---- bdary_friction!  diffusion.jl:64-80 ---------------------------------------------------------------------
// vDirichlet is a closure in the reference; the ABI takes the per-wall constants the examples use (cavity.jl:41-44),
// indexed by -label-1 = UP, RIGHT, DOWN, LEFT.  Only the polygon's own fields are touched.
// Per-edge form (lv_step_bdary_friction_ex): v_edge[e] = vDirichlet(m) and on_edge[e] = charfun(m) evaluated by the host at
// the midpoint of boundary edge number e (numbering of lv_boundary_edges); wall_on switches whole walls off.
struct WallVel { double v[8]; unsigned char on[4]; };
__global__ void __launch_bounds__(ST_BLOCK) k_bdary_friction(StepView S, double dt, WallVel w, const int *__restrict__ bptr,
                                                             const double2 *__restrict__ v_edge, const unsigned char *__restrict__ on_edge) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s = S.prim[i];
    if (s < 0) return;
    const double2 x = S.x[i];
    double2 v = S.v[i];
    const double mu = S.mu[i], mass = S.mass[i];
    double e = S.e[i], tmp = 1.0;
    const int r0 = S.rowptr[s], d = S.rdeg[s];
    int be = bptr ? bptr[i] : 0;
    for (int k = r0; k < r0 + d; k++) {
        const int lab = S.col[k];
        if (lab >= 0) continue; // boundaries(p): wall codes are negative
        const int e_no = be++;
        const int wk0 = (lab >= -4) ? -lab - 1 : -1;
        if (wk0 >= 0 && !w.on[wk0]) continue;            // if !charfun(m) continue end  diffusion.jl:69
        if (on_edge && !on_edge[e_no]) continue;
        const double2 a = S.v1[k], b = S.v2[k];
        const double mx = 0.5 * (a.x + b.x), my = 0.5 * (a.y + b.y);
        double nx = a.y - b.y, ny = b.x - a.x; // normal_vector  polygon.jl:153-156
        const double nn = sqrt(nx * nx + ny * ny);
        nx /= nn; ny /= nn;
        const double ex = a.x - b.x, ey = a.y - b.y;
        const double lrr = sqrt(ex * ex + ey * ey) / fabs((mx - x.x) * nx + (my - x.y) * ny);
        const double c = mu * lrr;
        const int wk = (lab >= -4) ? -lab - 1 : -1;
        double vdx = wk >= 0 ? w.v[2 * wk] : 0.0, vdy = wk >= 0 ? w.v[2 * wk + 1] : 0.0;
        if (v_edge) { vdx = v_edge[e_no].x; vdy = v_edge[e_no].y; }
        const double fx = (c * vdx) / mass, fy = (c * vdy) / mass;
        tmp += ((dt * mu) * lrr) / mass;
        e += dt * (fx * v.x + fy * v.y);
        v = make_double2(v.x + dt * fx, v.y + dt * fy);
    }
    S.e[i] = e;
    S.v[i] = make_double2(v.x / tmp, v.y / tmp);
}

// ---- EOS  pressure.jl:32-70 ----------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_eos(StepView S, double gamma, double p0, int stiffened) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s = S.prim[i];
    if (s < 0) return;
    const double2 v = S.v[i];
    const double rho = S.mass[i] / S.area[s];
    const double eint = S.e[i] - 0.5 * (v.x * v.x + v.y * v.y);
    const double P = ((gamma - 1.0) * rho) * eint;
    S.rho[i] = rho;
    S.P[i] = P;
    if (stiffened) S.c2[i] = (gamma * (P + p0)) / rho;              // stiffened_eos!  :64-70
    else S.c2[i] = (gamma * (P > p0 ? P : p0)) / rho;               // ideal_eos!      :49-55 (p0 = Pmin)
}

// ---- pressure_step!  pressure.jl:10-25 -------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_pressure_step_v(StepView S, double dt, double2 *vout) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 x = S.x[i];
    double2 v = S.v[i];
    const double Pi = S.P[i], mi = S.mass[i];
    FOR_NEIGHBORS(S, i, x) {
        const double lrr = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb);
        const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
        const double sc = ((dt / mi) * lrr) * (Pi - S.P[q]);
        v = make_double2(v.x + sc * (mx - x.x), v.y + sc * (my - x.y));
    }
    vout[i] = v; // the reference updates p.v in place; neighbours' v is not read in this sweep
}
__global__ void __launch_bounds__(ST_BLOCK) k_pressure_step_e(StepView S, double dt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 x = S.x[i], v = S.v[i];
    const double Pi = S.P[i], mi = S.mass[i];
    double e = S.e[i];
    FOR_NEIGHBORS(S, i, x) {
        const double lrr = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb);
        const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
        const double2 vq = S.v[q];
        const double Pq = S.P[q];
        const double a = (mx - x.x) * (Pi * v.x) + (my - x.y) * (Pi * v.y);
        const double b = (mx - y.x) * (Pq * vq.x) + (my - y.y) * (Pq * vq.y);
        e -= ((dt * lrr) / mi) * (a - b);
    }
    S.e[i] = e;
}

__global__ void __launch_bounds__(ST_BLOCK) k_gravity(StepView S, double gx, double gy, double dt) { // pressure.jl:77-82
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 v = S.v[i];
    S.v[i] = make_double2(v.x + dt * gx, v.y + dt * gy);
}

// ---- find_D!  diffusion.jl:8-19 -----------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_find_D(StepView S) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s0 = S.prim[i];
    if (s0 < 0) return;
    const double2 x = S.x[i], v = S.v[i];
    double D0 = 0, D1 = 0, D2 = 0, D3 = 0;
    FOR_NEIGHBORS(S, i, x) {
        const double lrr = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb);
        const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
        const double2 vq = S.v[q];
        const double ax = v.x - vq.x, ay = v.y - vq.y, bx = mx - y.x, by = my - y.y;
        D0 += lrr * (ax * bx); D1 += lrr * (ay * bx); D2 += lrr * (ax * by); D3 += lrr * (ay * by); // outer  geometry.jl:186-188
    }
    const double A = S.area[s0];
    D0 /= A; D1 /= A; D2 /= A; D3 /= A;
    S.D[4 * i + 0] = 0.5 * (D0 + D0);
    S.D[4 * i + 1] = 0.5 * (D1 + D2);
    S.D[4 * i + 2] = 0.5 * (D2 + D1);
    S.D[4 * i + 3] = 0.5 * (D3 + D3);
}

// getS  diffusion.jl:22-29
__device__ __forceinline__ void getS(const StepView &S, int i, double dr, double out[4]) {
    const double D0 = S.D[4 * i], D1 = S.D[4 * i + 1], D2 = S.D[4 * i + 2], D3 = S.D[4 * i + 3];
    const double divv = ((D0 * 1.0 + D1 * 0.0) + D2 * 0.0) + D3 * 1.0;
    double mu = S.mu[i];
    if (divv < 0.0) mu -= (divv * S.rho[i]) * (dr * dr);
    const double t = 2.0 * mu;
    out[0] = t * (D0 - (divv * 1.0) / 3.0);
    out[1] = t * (D1 - (divv * 0.0) / 3.0);
    out[2] = t * (D2 - (divv * 0.0) / 3.0);
    out[3] = t * (D3 - (divv * 1.0) / 3.0);
}

// ---- viscous_step!  diffusion.jl:39-53 ------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_viscous_v(StepView S, double dt, double avdr, double2 *vout) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 x = S.x[i];
    double2 v = S.v[i];
    const double mi = S.mass[i];
    double Sp[4];
    getS(S, (int)i, avdr, Sp);
    FOR_NEIGHBORS(S, i, x) {
        double Sq[4];
        getS(S, q, avdr, Sq);
        const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
        const double sc = (dt * lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb)) / mi;
        const double M0 = sc * (Sp[0] - Sq[0]), M1 = sc * (Sp[1] - Sq[1]), M2 = sc * (Sp[2] - Sq[2]), M3 = sc * (Sp[3] - Sq[3]);
        const double wx = mx - x.x, wy = my - x.y;
        v = make_double2(v.x - (M0 * wx + M2 * wy), v.y - (M1 * wx + M3 * wy));
    }
    vout[i] = v;
}
__global__ void __launch_bounds__(ST_BLOCK) k_viscous_e(StepView S, double dt, double avdr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 x = S.x[i], v = S.v[i];
    const double mi = S.mass[i];
    double e = S.e[i];
    double Sp[4];
    getS(S, (int)i, avdr, Sp);
    FOR_NEIGHBORS(S, i, x) {
        double Sq[4];
        getS(S, q, avdr, Sq);
        const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
        const double sc = (dt * lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb)) / mi;
        const double2 vq = S.v[q];
        const double a = (mx - x.x) * (Sp[0] * v.x + Sp[2] * v.y) + (my - x.y) * (Sp[1] * v.x + Sp[3] * v.y);
        const double b = (mx - y.x) * (Sq[0] * vq.x + Sq[2] * vq.y) + (my - y.y) * (Sq[1] * vq.x + Sq[3] * vq.y);
        e += sc * (a - b);
    }
    S.e[i] = e;
}

// ---- find_dv!  relaxation.jl:10-25 ------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_find_dv(StepView S, double dt, double alpha) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s0 = S.prim[i];
    if (s0 < 0) return;
    const double2 x = S.x[i];
    const double2 c = S.cen[s0];
    double rmax = 0.0, rmin = __longlong_as_double(0x7ff0000000000000ll);
    FOR_NEIGHBORS(S, i, x) {
        (void)q; (void)ea; (void)eb;
        const double dx = x.x - y.x, dy = x.y - y.y;
        const double r = sqrt(dx * dx + dy * dy);
        rmax = rmax > r ? rmax : r;
        rmin = rmin < r ? rmin : r;
    }
    const double quality = rmin / rmax;
    const double D0 = S.D[4 * i], D1 = S.D[4 * i + 1], D2 = S.D[4 * i + 2], D3 = S.D[4 * i + 3];
    const double nD = sqrt(((D0 * D0 + D1 * D1) + D2 * D2) + D3 * D3);
    const double lambda = (alpha * nD) / (quality * quality);
    const double sc = lambda / (1.0 + dt * lambda);
    S.quality[i] = quality;
    S.dv[i] = make_double2(sc * (c.x - x.x), sc * (c.y - x.y));
}

// ---- relaxation_step!  relaxation.jl:36-73 ---------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_relax_flux(StepView S, double dt, int rusanov, double *mass_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 x = S.x[i], v = S.v[i], dvp = S.dv[i];
    const double rho = S.rho[i], e = S.e[i], ph = S.phase[i];
    double mass = S.mass[i];
    double2 mom = make_double2(mass * v.x, mass * v.y);
    double energy = mass * e;
    const double ndvp = sqrt(dvp.x * dvp.x + dvp.y * dvp.y);
    FOR_NEIGHBORS(S, i, x) {
        if (!(ph == S.phase[q])) continue;
        const double lrr = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb);
        const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
        const double zx = 0.5 * (x.x + y.x), zy = 0.5 * (x.y + y.y);
        const double pqx = x.x - y.x, pqy = x.y - y.y, mzx = mx - zx, mzy = my - zy;
        const double2 dvq = S.dv[q], vq = S.v[q];
        const double rq = S.rho[q], eq = S.e[q];
        const double pdvpq = dvp.x * pqx + dvp.y * pqy, qdvpq = dvq.x * pqx + dvq.y * pqy;
        const double pdvmz = dvp.x * mzx + dvp.y * mzy, qdvmz = dvq.x * mzx + dvq.y * mzy;
        const double cc = dt * lrr;
        mass += cc * ((pdvmz * rho - qdvmz * rq) - 0.5 * (pdvpq * rho + qdvpq * rq));
        const double a1 = pdvmz * rho, a2 = qdvmz * rq, a3 = pdvpq * rho, a4 = qdvpq * rq;
        mom.x += cc * ((a1 * v.x - a2 * vq.x) - 0.5 * (a3 * v.x + a4 * vq.x));
        mom.y += cc * ((a1 * v.y - a2 * vq.y) - 0.5 * (a3 * v.y + a4 * vq.y));
        energy += cc * ((a1 * e - a2 * eq) - 0.5 * (a3 * e + a4 * eq));
        if (rusanov) {
            const double ndvq = sqrt(dvq.x * dvq.x + dvq.y * dvq.y);
            const double a = ndvp > ndvq ? ndvp : ndvq;
            const double ex = ea.x - eb.x, ey = ea.y - eb.y;
            const double l = sqrt(ex * ex + ey * ey); // len(e)  geometry.jl:136-138
            const double kk = ((0.5 * dt) * l) * a;
            mass += kk * (rq - rho);
            mom.x += kk * (rq * vq.x - rho * v.x);
            mom.y += kk * (rq * vq.y - rho * v.y);
            energy += kk * (rq * eq - rho * e);
        }
    }
    mass_out[i] = mass; // p.mass is only read by its own thread in the reference's first loop
    S.mom[i] = mom;
    S.energy[i] = energy;
}
__global__ void __launch_bounds__(ST_BLOCK) k_relax_apply(StepView S, double dt, const double *mass_new) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double m = mass_new[i];
    const double2 mom = S.mom[i], dv = S.dv[i], x = S.x[i];
    S.mass[i] = m;
    S.v[i] = make_double2(mom.x / m, mom.y / m);
    S.e[i] = S.energy[i] / m;
    S.x[i] = make_double2(x.x + dt * dv.x, x.y + dt * dv.y);
}

__global__ void __launch_bounds__(ST_BLOCK) k_to_centroid(StepView S) { // p.x = centroid(p)  populate.jl:136-138
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s = S.prim[i];
    if (s >= 0) S.x[i] = S.cen[s];
}

// ---- multiphase projector  relaxation.jl:75-206 -------------------------------------------------------------------
// mul!(res, A::MultiphaseProjector, x)  relaxation.jl:91-123: two sweeps, only edges between different phases contribute
__global__ void __launch_bounds__(ST_BLOCK) k_mp_tmp(StepView S, const double *__restrict__ xv, double2 *__restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s0 = S.prim[i];
    const double2 x = S.x[i];
    const double ph = S.phase[i], xi = xv[i];
    double tx = 0.0, ty = 0.0;
    FOR_NEIGHBORS(S, i, x) {
        if (ph != S.phase[q]) {
            const double lrr = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb);
            const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
            const double sc = lrr * (xi - xv[q]);
            tx -= sc * (mx - x.x);
            ty -= sc * (my - x.y);
        }
    }
    const double A = s0 >= 0 ? S.area[s0] : 1.0;
    tmp[i] = make_double2(tx / A, ty / A);
}
// second sweep of mul! (tvec = tmp_vec) and, with tvec = dv and sign = +1, the right-hand side of refresh! (:162-177)
__global__ void __launch_bounds__(ST_BLOCK) k_mp_apply(StepView S, const double2 *__restrict__ tvec, double *__restrict__ res) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double2 x = S.x[i], ti = tvec[i];
    const double ph = S.phase[i];
    double r = 0.0;
    FOR_NEIGHBORS(S, i, x) {
        if (ph != S.phase[q]) {
            const double lrr = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb);
            const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
            const double zx = 0.5 * (x.x + y.x), zy = 0.5 * (x.y + y.y);
            const double2 tj = tvec[q];
            const double a = (ti.x - tj.x) * (mx - zx) + (ti.y - tj.y) * (my - zy);
            const double b = (ti.x + tj.x) * (x.x - y.x) + (ti.y + tj.y) * (x.y - y.y);
            r -= lrr * (a - 0.5 * b);
        }
    }
    res[i] = r;
}
// p.dv += lr_ratio*(res_i - res_j)*(m - p.x)/area(p) over interface edges, cells of poor quality skipped (:191-203)
__global__ void __launch_bounds__(ST_BLOCK) k_mp_update(StepView S, const double *__restrict__ res, double quality_threshold, double2 *__restrict__ dv_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int s0 = S.prim[i];
    const double2 x = S.x[i];
    double2 dv = S.dv[i];
    const double ph = S.phase[i], ri = res[i];
    if (!(S.quality[i] < quality_threshold) && s0 >= 0) {
        const double A = S.area[s0];
        FOR_NEIGHBORS(S, i, x) {
            if (ph != S.phase[q]) {
                const double mx = 0.5 * (ea.x + eb.x), my = 0.5 * (ea.y + eb.y);
                const double sc = lr_ratio(make_double2(x.x - y.x, x.y - y.y), ea, eb) * (ri - res[q]);
                dv = make_double2(dv.x + (sc * (mx - x.x)) / A, dv.y + (sc * (my - x.y)) / A);
            }
        }
    }
    dv_out[i] = dv;
}

// ---- state management --------------------------------------------------------------------------------------------
static const char *const FIELD_NAMES[] = {"x", "v", "dv", "momentum", "rho", "e", "P", "c2", "mass", "energy", "quality", "mu", "phase", "D"};
static const int FIELD_NC[] = {2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 4};
enum { F_X = 0, F_V, F_DV, F_MOM, F_RHO, F_E, F_P, F_C2, F_MASS, F_ENERGY, F_QUALITY, F_MU, F_PHASE, F_D, F_COUNT };

static int field_index(const char *name) {
    for (int k = 0; k < F_COUNT; k++)
        if (!strcmp(name, FIELD_NAMES[k])) return k;
    return -1;
}

// Strip mode (lv_strip.cu): the state arrays cover the LOCAL generator list -- owned generators first, at [0, n_own), then the
// ghosts -- with the fixed capacity of the strip; "x" is the strip's own position array (the ghost exchange of every remesh
// writes the ghosts' positions there); sweeps run over the owned generators and read ghosts through label-order halos.
static inline bool st_strip(const LvContext *c) { return c->strip.on; }
static int state_ensure(LvContext *c, int64_t n) {
    if (st_strip(c)) {
        if (n > c->strip.cap_loc) return lv_set_error(c, LV_ECAPACITY, "strip: %lld generators exceed the local capacity", (long long)n);
        if (!c->st_field[F_V] || c->st_cap < c->strip.cap_loc) {
            LV_CUDA(c, cudaStreamSynchronize(c->stream));
            const int64_t cap = c->strip.cap_loc;
            for (int k = 0; k < F_COUNT; k++) {
                if (k == F_X) continue;
                if (c->st_field[k]) lv_free(c, c->st_field[k], sizeof(double) * (size_t)FIELD_NC[k] * (size_t)c->st_cap);
                LV_TRY(lv_alloc(c, (void **)&c->st_field[k], sizeof(double) * (size_t)FIELD_NC[k] * (size_t)cap));
                LV_CUDA(c, cudaMemsetAsync(c->st_field[k], 0, sizeof(double) * (size_t)FIELD_NC[k] * (size_t)cap, c->stream));
            }
            if (c->st_tmp) lv_free(c, c->st_tmp, sizeof(double) * 2 * (size_t)c->st_cap);
            LV_TRY(lv_alloc(c, (void **)&c->st_tmp, sizeof(double) * 2 * (size_t)cap));
            c->st_cap = cap;
        }
        c->st_field[F_X] = (double *)c->strip.loc_xy;
        c->st_x_alias = true;
        c->st_n = n;
        return LV_OK;
    }
    if (c->st_cap >= n && c->st_field[0]) { c->st_n = n; return LV_OK; }
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    const int64_t cap = n + n / 16 + 64;
    for (int k = 0; k < F_COUNT; k++) {
        if (c->st_field[k]) lv_free(c, c->st_field[k], sizeof(double) * (size_t)FIELD_NC[k] * (size_t)c->st_cap);
        LV_TRY(lv_alloc(c, (void **)&c->st_field[k], sizeof(double) * (size_t)FIELD_NC[k] * (size_t)cap));
        LV_CUDA(c, cudaMemsetAsync(c->st_field[k], 0, sizeof(double) * (size_t)FIELD_NC[k] * (size_t)cap, c->stream));
    }
    if (c->st_tmp) lv_free(c, c->st_tmp, sizeof(double) * 2 * (size_t)c->st_cap);
    LV_TRY(lv_alloc(c, (void **)&c->st_tmp, sizeof(double) * 2 * (size_t)cap));
    c->st_cap = cap;
    c->st_n = n;
    return LV_OK;
}

static StepView make_view(LvContext *c) {
    StepView S;
    S.g = c->gp;
    S.n = c->st_n;
    S.prim = c->d_prim_of_label; S.rowptr = c->d_rowptr; S.rdeg = c->d_deg; S.col = c->d_col;
    S.v1 = c->d_v1; S.v2 = c->d_v2; S.ent_label = c->d_ent_label; S.area = c->d_area; S.cen = c->d_cen;
    S.x = (double2 *)c->st_field[F_X]; S.v = (double2 *)c->st_field[F_V]; S.dv = (double2 *)c->st_field[F_DV];
    S.mom = (double2 *)c->st_field[F_MOM];
    S.rho = c->st_field[F_RHO]; S.e = c->st_field[F_E]; S.P = c->st_field[F_P]; S.c2 = c->st_field[F_C2];
    S.mass = c->st_field[F_MASS]; S.energy = c->st_field[F_ENERGY]; S.quality = c->st_field[F_QUALITY];
    S.mu = c->st_field[F_MU]; S.phase = c->st_field[F_PHASE]; S.D = c->st_field[F_D];
    return S;
}

static int need_mesh(LvContext *c) {
    if (!c->st_field[0]) return lv_set_error(c, LV_EINVAL, "no device state: call lv_state_set(\"x\", ...) first");
    const int64_t n_mesh = st_strip(c) ? c->strip.n_loc : c->st_n;
    if (!c->mesh_valid || c->n != n_mesh || (st_strip(c) && c->st_n != c->strip.n_own))
        return lv_set_error(c, LV_EINVAL, "no mesh for the device state: call lv_state_remesh first");
    return LV_OK;
}

// ghosts of label-ordered fields from their owners (no-op on one GPU)
static int st_halo(LvContext *c, std::initializer_list<int> fields) {
    if (!st_strip(c)) return LV_OK;
    for (int k : fields) {
        if (FIELD_NC[k] <= 4) LV_TRY(lv_strip_halo_state(c, c->st_field[k], FIELD_NC[k]));
    }
    return LV_OK;
}

extern "C" int32_t lv_strip_remesh(LvHandle c, int64_t *counts_out);
static int state_remesh(LvContext *c) {
    if (!c->st_field[0]) return lv_set_error(c, LV_EINVAL, "no device state");
    if (st_strip(c)) {
        // generators that left the strip take their fields to the new owner, then the strip remesh (ghost exchange + clipping)
        double *fields[F_COUNT];
        int nc[F_COUNT], nf = 0;
        for (int k = 0; k < F_COUNT; k++) {
            if (k == F_X) continue;
            fields[nf] = c->st_field[k]; nc[nf] = FIELD_NC[k]; nf++;
        }
        LV_TRY(lv_strip_migrate(c, nf, fields, nc));
        c->st_n = c->strip.n_own;
        return lv_strip_remesh(c, nullptr);
    }
    return lv_remesh_dev(c, c->st_n, c->st_field[F_X]);
}

#define GRID(n) (int)(((n) + ST_BLOCK - 1) / ST_BLOCK), ST_BLOCK, 0, c->stream

extern "C" {

// polygon fields in label order: x v dv momentum (2 each), rho e P c2 mass energy quality mu phase, D (4)
int32_t lv_state_set(LvHandle c, const char *name, const double *host, int64_t n) {
    if (!c || !name || (!host && n > 0)) return LV_EINVAL;
    LV_ENTER(c);
    const int k = field_index(name);
    if (k < 0) return lv_set_error(c, LV_EINVAL, "unknown field '%s'", name);
    if (k == F_X) {
        LV_TRY(state_ensure(c, n));
        if (st_strip(c)) c->strip.n_own = n; // the owned generators; their global labels come from lv_strip_set_owned
    } else if (!c->st_field[0] || n != c->st_n) return lv_set_error(c, LV_EINVAL, "set \"x\" first; field length must equal the number of polygons");
    if (n > 0) LV_CUDA(c, cudaMemcpyAsync(c->st_field[k], host, sizeof(double) * (size_t)FIELD_NC[k] * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    if (k == F_X) c->mesh_valid = false;
    return LV_OK;
}
int32_t lv_state_get(LvHandle c, const char *name, double *host) {
    if (!c || !name || !host) return LV_EINVAL;
    LV_ENTER(c);
    const int k = field_index(name);
    if (k < 0 || !c->st_field[0]) return lv_set_error(c, LV_EINVAL, "unknown field or no device state");
    if (c->st_n > 0) LV_CUDA(c, cudaMemcpyAsync(host, c->st_field[k], sizeof(double) * (size_t)FIELD_NC[k] * (size_t)c->st_n, cudaMemcpyDeviceToHost, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}
int32_t lv_state_ptr(LvHandle c, const char *name, void **dev_ptr, int64_t *n) { // device pointer of a field (label order)
    if (!c || !name || !dev_ptr) return LV_EINVAL;
    const int k = field_index(name);
    if (k < 0 || !c->st_field[0]) return lv_set_error(c, LV_EINVAL, "unknown field or no device state");
    *dev_ptr = c->st_field[k];
    if (n) *n = c->st_n;
    return LV_OK;
}

// Strip mode: make the generators this rank owns (lv_strip_set_owned) the resident state -- "x" is the strip's own position
// array, every other field is allocated for the local list (owned + ghosts) and zeroed; fill them with lv_state_set (n_own
// values each).  From then on move! / relaxation_step! migrate generators between ranks together with their fields.
int32_t lv_state_attach_strip(LvHandle c) {
    if (!c || !c->strip.on) return lv_set_error(c, LV_EINVAL, "lv_state_attach_strip: not in strip mode");
    LV_ENTER(c);
    return state_ensure(c, c->strip.n_own);
}

int32_t lv_state_remesh(LvHandle c) { // remesh!(grid) on the resident positions
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    return state_remesh(c);
}

int32_t lv_step_move(LvHandle c, double dt) { // move!(grid, dt)  move.jl:9-21
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    LV_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 8, c->stream));
    if (S.n > 0) { k_move<<<GRID(S.n)>>>(S, dt, c->bmin[0], c->bmin[1], c->bmax[0], c->bmax[1], c->d_flags); c->launches++; }
    LV_TRY(lv_publish_flags(c, nullptr));
    if (c->h_flags[LVF_NAN]) return lv_set_error(c, LV_ENAN, "Velocity field invalidated.");
    return state_remesh(c);
}

int32_t lv_step_eos(LvHandle c, double gamma, double p0, int32_t stiffened) { // stiffened_eos!(grid, gamma, P0) / ideal_eos!(grid, gamma; Pmin)
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    if (S.n > 0) { k_eos<<<GRID(S.n)>>>(S, gamma, p0, stiffened); c->launches++; }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int32_t lv_step_pressure_step(LvHandle c, double dt) { // pressure_step!(grid, dt)  pressure.jl:10-25
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    LV_TRY(st_halo(c, {F_P}));
    if (S.n > 0) {
        k_pressure_step_v<<<GRID(S.n)>>>(S, dt, (double2 *)c->st_tmp);
        LV_CUDA(c, cudaMemcpyAsync(S.v, c->st_tmp, sizeof(double2) * (size_t)S.n, cudaMemcpyDeviceToDevice, c->stream));
    }
    LV_TRY(st_halo(c, {F_V})); // the energy sweep reads the neighbours' NEW velocities
    if (S.n > 0) {
        k_pressure_step_e<<<GRID(S.n)>>>(S, dt);
        c->launches += 2;
    }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int32_t lv_step_gravity(LvHandle c, double gx, double gy, double dt) { // gravity_step!(grid, g, dt)  pressure.jl:77-82
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->st_field[0]) return lv_set_error(c, LV_EINVAL, "no device state");
    StepView S = make_view(c);
    if (S.n > 0) { k_gravity<<<GRID(S.n)>>>(S, gx, gy, dt); c->launches++; }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int32_t lv_step_find_D(LvHandle c) { // find_D!(grid)  diffusion.jl:8-19
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    LV_TRY(st_halo(c, {F_V}));
    if (S.n > 0) { k_find_D<<<GRID(S.n)>>>(S); c->launches++; }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int32_t lv_step_viscous_step(LvHandle c, double dt, int32_t artificial_viscosity) { // viscous_step!  diffusion.jl:39-53
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    const double avdr = artificial_viscosity ? c->dr : 0.0;
    LV_TRY(st_halo(c, {F_D, F_MU, F_RHO})); // getS of the neighbours  diffusion.jl:22-29
    if (S.n > 0) {
        k_viscous_v<<<GRID(S.n)>>>(S, dt, avdr, (double2 *)c->st_tmp);
        LV_CUDA(c, cudaMemcpyAsync(S.v, c->st_tmp, sizeof(double2) * (size_t)S.n, cudaMemcpyDeviceToDevice, c->stream));
    }
    LV_TRY(st_halo(c, {F_V}));
    if (S.n > 0) {
        k_viscous_e<<<GRID(S.n)>>>(S, dt, avdr);
        c->launches += 2;
    }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int32_t lv_step_bdary_friction_ex(LvHandle c, double dt, const double *vwall, const uint8_t *wall_on, const double *v_edge,
                                  const uint8_t *on_edge, int64_t n_edge) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    WallVel w;
    for (int k = 0; k < 8; k++) w.v[k] = vwall ? vwall[k] : 0.0;
    for (int k = 0; k < 4; k++) w.on[k] = wall_on ? (wall_on[k] != 0) : 1;
    const bool per_edge = v_edge || on_edge;
    if (per_edge) {
        LV_TRY(lv_bdry_index(c));
        if (n_edge != c->n_bedge) return lv_set_error(c, LV_EINVAL, "wall data for %lld edges, the mesh has %lld boundary edges", (long long)n_edge, (long long)c->n_bedge);
        if (n_edge > c->cap_bedge || !c->d_vbc_edge) {
            LV_CUDA(c, cudaStreamSynchronize(c->stream));
            if (c->d_vbc_edge) cudaFree(c->d_vbc_edge);
            if (c->d_bf_on) cudaFree(c->d_bf_on);
            c->cap_bedge = n_edge + n_edge / 8 + 64;
            LV_CUDA(c, cudaMalloc((void **)&c->d_vbc_edge, sizeof(double2) * (size_t)c->cap_bedge));
            LV_CUDA(c, cudaMalloc((void **)&c->d_bf_on, (size_t)c->cap_bedge));
            c->vbc_edge_on = false;
            c->bvel_valid = false;
        }
        // the per-edge velocity buffer is shared with lv_set_boundary_velocity: a later right-hand side must be given its own
        if (v_edge && n_edge > 0) { LV_CUDA(c, cudaMemcpyAsync(c->d_vbc_edge, v_edge, sizeof(double2) * (size_t)n_edge, cudaMemcpyHostToDevice, c->stream)); c->vbc_edge_on = false; c->bvel_valid = false; }
        if (on_edge && n_edge > 0) LV_CUDA(c, cudaMemcpyAsync(c->d_bf_on, on_edge, (size_t)n_edge, cudaMemcpyHostToDevice, c->stream));
    }
    if (S.n > 0) {
        k_bdary_friction<<<GRID(S.n)>>>(S, dt, w, per_edge ? c->d_bdry_ptr : nullptr, v_edge ? c->d_vbc_edge : nullptr, on_edge ? c->d_bf_on : nullptr);
        c->launches++;
    }
    LV_CUDA(c, cudaGetLastError());
    if (per_edge) LV_CUDA(c, cudaStreamSynchronize(c->stream)); // the host arrays may be reused
    return LV_OK;
}
int32_t lv_step_bdary_friction(LvHandle c, double dt, const double *vwall) { // per-wall constants, charfun = everywhere
    return lv_step_bdary_friction_ex(c, dt, vwall, nullptr, nullptr, nullptr, 0);
}

int32_t lv_step_find_dv(LvHandle c, double dt, double alpha) { // find_dv!(grid, dt, alpha)  relaxation.jl:10-25
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    if (S.n > 0) { k_find_dv<<<GRID(S.n)>>>(S, dt, alpha); c->launches++; }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int32_t lv_step_relaxation_step(LvHandle c, double dt, int32_t rusanov) { // relaxation_step!  relaxation.jl:36-73
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    StepView S = make_view(c);
    LV_TRY(st_halo(c, {F_DV, F_V, F_RHO, F_E, F_PHASE}));
    if (S.n > 0) {
        k_relax_flux<<<GRID(S.n)>>>(S, dt, rusanov, c->st_tmp);
        k_relax_apply<<<GRID(S.n)>>>(S, dt, c->st_tmp);
        c->launches += 2;
    }
    LV_CUDA(c, cudaGetLastError());
    return state_remesh(c);
}

// populate_lloyd!'s relaxation loop (populate.jl:132-145): niter x (remesh!; p.x = centroid(p)), then a final remesh!
int32_t lv_step_lloyd(LvHandle c, int32_t niter) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    if (!c->st_field[0]) return lv_set_error(c, LV_EINVAL, "no device state");
    for (int it = 0; it < niter; it++) {
        LV_TRY(state_remesh(c));
        StepView S = make_view(c);
        if (S.n > 0) { k_to_centroid<<<GRID(S.n)>>>(S); c->launches++; }
    }
    return state_remesh(c);
}

// multiphase_projection!(solver)  relaxation.jl:179-206: MINRES (atol = rtol = 1e-4, itmax = 200 in the reference) on the
// matrix-free projector, then the correction of dv.  *solved = 0 reproduces the reference's @warn (it carries on).
int32_t lv_step_multiphase_projection(LvHandle c, double quality_threshold, double rtol, double atol, int32_t itmax, int32_t *iters,
                                      int32_t *solved) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    if (c->comm && !lv_strip_peer_mode(c)) return lv_set_error(c, LV_EINVAL, "the multiphase projector on several GPUs needs the peer-memory strip exchange");
    LV_TRY(lv_pr_ensure(c));
    StepView S = make_view(c);
    const int n = (int)S.n;
    if (n == 0 && !c->comm) return LV_OK;
    // On strips the vectors are label-ordered over the local generator list: rows [0, n_own) are this rank's unknowns, the
    // ghost entries behind them are refreshed from their owners before every sweep that reads neighbours; dot products
    // run over the owned rows and are summed over the ranks by lv_minres_apply.
    double *b = c->d_vec[6], *sol = c->d_vec[5];
    double2 *tmp = (double2 *)c->st_tmp;
    LV_TRY(st_halo(c, {F_PHASE, F_DV}));
    if (n > 0) { k_mp_apply<<<GRID(S.n)>>>(S, S.dv, b); c->launches++; } // refresh!: b_i = -sum lrr (dot(dv_i - dv_j, m-z) - 0.5 dot(dv_i + dv_j, x-y))
    auto apply = [&](const double *in, double *out) -> int {
        if (st_strip(c)) LV_TRY(lv_strip_halo_state(c, (double *)in, 1));
        if (n > 0) k_mp_tmp<<<GRID(S.n)>>>(S, in, tmp);
        if (st_strip(c)) LV_TRY(lv_strip_halo_state(c, (double *)tmp, 2));
        if (n > 0) k_mp_apply<<<GRID(S.n)>>>(S, tmp, out);
        c->launches += 2;
        return LV_OK;
    };
    int it = 0, ok = 1;
    LV_TRY(lv_minres_apply(c, n, apply, b, sol, rtol, atol, itmax, &it, &ok));
    if (iters) *iters = it;
    if (solved) *solved = ok;
    if (st_strip(c)) LV_TRY(lv_strip_halo_state(c, sol, 1));
    if (n > 0) {
        k_mp_update<<<GRID(S.n)>>>(S, sol, quality_threshold, tmp);
        LV_CUDA(c, cudaMemcpyAsync(S.dv, tmp, sizeof(double2) * (size_t)S.n, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->launches++;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// mul!(res, A::MultiphaseProjector, x) (relaxation.jl:91-123) and the right-hand side of refresh! (:162-177) on host
// vectors in label order: the operator the projection solves with, exposed for direct parity checks.  NULL skips.
int32_t lv_step_multiphase_apply(LvHandle c, const double *x, double *y, double *b) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    LV_TRY(lv_pr_ensure(c));
    StepView S = make_view(c);
    if (S.n == 0) return LV_OK;
    const size_t bytes = sizeof(double) * (size_t)S.n;
    double *dx = c->d_vec[5], *dy = c->d_vec[6];
    double2 *tmp = (double2 *)c->st_tmp;
    LV_TRY(st_halo(c, {F_PHASE, F_DV}));
    if (x && y) {
        LV_CUDA(c, cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, c->stream));
        if (st_strip(c)) LV_TRY(lv_strip_halo_state(c, dx, 1));
        k_mp_tmp<<<GRID(S.n)>>>(S, dx, tmp);
        if (st_strip(c)) LV_TRY(lv_strip_halo_state(c, (double *)tmp, 2));
        k_mp_apply<<<GRID(S.n)>>>(S, tmp, dy);
        c->launches += 2;
        LV_CUDA(c, cudaMemcpyAsync(y, dy, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    if (b) {
        k_mp_apply<<<GRID(S.n)>>>(S, S.dv, dy);
        c->launches++;
        LV_CUDA(c, cudaMemcpyAsync(b, dy, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}

// find_pressure!(solver, dt, niter; boundary_velocity) on the resident state  pressure.jl:215-225
int32_t lv_step_find_pressure(LvHandle c, double dt, int32_t niter, double rtol, double atol, int32_t itmax, int32_t solver,
                              const double *vbc_wall, int32_t *iters_out, double *relres_out) {
    if (!c) return LV_EINVAL;
    LV_ENTER(c);
    LV_TRY(need_mesh(c));
    LV_TRY(lv_fields_upload_dev(c, c->st_field[F_MASS], c->st_field[F_RHO], c->st_field[F_C2], c->st_field[F_P], c->st_field[F_V]));
    LV_TRY(lv_pr_find_pressure(c, dt, niter, rtol, atol, itmax, solver, vbc_wall, iters_out, relres_out));
    return lv_scatter_to_labels(c, c->d_P, c->st_field[F_P], 1);
}

} // extern "C"
