/* clip_trace.c -- development tool (NOT product, NOT a test): replays voronoicut!(grid, poly) of the CPU restatement for
 * every polygon of a periodic jittered lattice and writes, per polygon, the sequence of neighbour-walk events -- which
 * path node, which candidate, its distance, the influence radius before / after, whether it cut.  oracle/experiments/
 * clip_sim.py replays these traces under different warp-scheduling policies of the GPU clipping kernel (K2) with a cost
 * model calibrated on LV_CLIP_STATS, so that policies can be compared without a GPU.
 *
 *   gcc -O2 -ffp-contract=off -fopenmp -o /tmp/clip_trace oracle/experiments/clip_trace.c -lm
 *   /tmp/clip_trace 128 0 /tmp/trace.bin        (lattice side, seed, output)
 */
#include "../lv_oracle.c"
#include <stdio.h>

static uint64_t mix64(uint64_t z) { /* lagrangianvoronoi.jl_b200/synthetic.py */
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct { int32_t t; int32_t kind; double d2, prr_before, prr_after; int32_t nv, pad; } ev_t;
/* kind: 0 node entered (d2 = node.rr, nv = bucket population, -1 out of bounds), 1 self, 2 fails the distance filter,
 *       3 passes, no cut, 4 cut, 5 walk ended at node t (rr > prr), 6 path exhausted */

int main(int argc, char **argv) {
    int M = argc > 1 ? atoi(argv[1]) : 128;
    int seed = argc > 2 ? atoi(argv[2]) : 0;
    const char *out = argc > 3 ? argv[3] : "/tmp/trace.bin";
    double dr = 1.0 / M;
    double bmin[2] = {0, 0}, bmax[2] = {1, 1};
    lvo_grid *g = lvo_grid_create(bmin, bmax, dr, 2.0 * dr, 10.0 * dr, 1, 1, 0);
    int64_t n = (int64_t)M * M;
    double *xy = (double *)malloc(sizeof(double) * 2 * n);
    for (int64_t i = 0; i < n; i++) { /* same generator as synthetic.jittered_lattice: read from stdin instead when given "-" */
        xy[2 * i] = 0; xy[2 * i + 1] = 0;
    }
    if (fread(xy, sizeof(double), 2 * n, stdin) != (size_t)(2 * n)) { fprintf(stderr, "expected %lld points on stdin\n", (long long)n); return 1; }
    (void)seed; (void)mix64;
    lvo_set_points(g, n, xy);
    omp_set_num_threads(1);
    if (lvo_remesh(g) != 0) { fprintf(stderr, "remesh failed\n"); return 1; }
    FILE *f = fopen(out, "wb");
    int64_t hdr[4] = {n, g->n1, g->n2, g->npath};
    fwrite(hdr, sizeof(hdr), 1, f);
    ev_t *buf = (ev_t *)malloc(sizeof(ev_t) * 100000);
    for (int64_t ip = 0; ip < n; ip++) {
        poly_t *poly = g->polygons[ip];
        reset_poly(poly, g->cmin, g->cmax);
        vec2 x = poly->x;
        double prr = influence_rr(poly);
        int64_t k1, k2;
        findkey(g, x, &k1, &k2);
        int ne = 0;
        int64_t t;
        for (t = 0; t < g->npath; t++) {
            double rr = g->path[t].rr;
            if (rr > prr) { buf[ne++] = (ev_t){(int32_t)t, 5, rr, prr, prr, 0, 0}; break; }
            int64_t c1 = k1 + g->path[t].i1, c2 = k2 + g->path[t].i2;
            if (!inbounds(g, c1, c2)) { buf[ne++] = (ev_t){(int32_t)t, 0, rr, prr, prr, -1, 0}; continue; }
            const fv_int *cell = &CELL(g, c1, c2);
            buf[ne++] = (ev_t){(int32_t)t, 0, rr, prr, prr, (int32_t)cell->last, 0};
            for (int64_t s = 0; s < cell->last; s++) {
                int64_t i = cell->data[s];
                const poly_t *q = g->polygons[i - 1];
                vec2 y = vadd(x, get_arrow(g, q->x, x));
                double d2 = norm_squared(vsub(x, y));
                int nv = (int)poly->edges.last;
                if (veq(x, y)) { buf[ne++] = (ev_t){(int32_t)t, 1, d2, prr, prr, nv, 0}; continue; }
                if (d2 > prr) { buf[ne++] = (ev_t){(int32_t)t, 2, d2, prr, prr, nv, 0}; continue; }
                double before = prr;
                /* would an axis-aligned bounding box of the vertices (relative to x) already rule the cut out?
                 * cut <=> some vertex v has d.(v - x) > |d|^2 / 2 (+ eps); the box gives an upper bound of d.(v - x) */
                int boxrej = 0;
                {
                    double lx = 1e300, hx = -1e300, ly = 1e300, hy = -1e300;
                    for (int64_t k = 0; k < poly->edges.last; k++) {
                        double vx = poly->edges.data[k].v1.x - x.x, vy = poly->edges.data[k].v1.y - x.y;
                        if (vx < lx) lx = vx; if (vx > hx) hx = vx; if (vy < ly) ly = vy; if (vy > hy) hy = vy;
                    }
                    double dx = y.x - x.x, dy = y.y - x.y;
                    double ub = dx * (dx > 0 ? hx : lx) + dy * (dy > 0 ? hy : ly);
                    boxrej = ub <= 0.5 * (dx * dx + dy * dy) * (1.0 - 1e-9);
                }
                if (voronoicut_poly(poly, y, i)) { prr = influence_rr(poly); buf[ne++] = (ev_t){(int32_t)t, 4, d2, before, prr, nv, boxrej}; }
                else buf[ne++] = (ev_t){(int32_t)t, 3, d2, before, prr, nv, boxrej};
            }
        }
        if (t == g->npath) buf[ne++] = (ev_t){(int32_t)t, 6, 0, prr, prr, 0, 0};
        int64_t rec[4] = {ip, (k1 - 1) + g->n1 * (k2 - 1), ne, (int64_t)poly->edges.last};
        fwrite(rec, sizeof(rec), 1, f);
        fwrite(buf, sizeof(ev_t), (size_t)ne, f);
    }
    fclose(f);
    fprintf(stderr, "wrote %s: %lld polygons\n", out, (long long)n);
    return 0;
}
