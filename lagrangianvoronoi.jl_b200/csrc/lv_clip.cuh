// lv_clip.cuh -- declarations shared by the two clipping kernels (lv_clip.cu: reference-order
// edge-list kernel, lv_clip_fast.cu: linked-slot kernel).
#pragma once
#include "lv_internal.cuh"

#define SIGNUM_EPS 4.440892098500626e-16 // 2*eps(Float64)  polygon.jl:2
#define BD_UP (-1)                       // polygon.jl:4-7
#define BD_RIGHT (-2)
#define BD_DOWN (-3)
#define BD_LEFT (-4)

struct ClipArgs {
    LvGridParams g;
    const LvPathNode *path;
    const int *cell_start;
    const unsigned *ent_label;
    const double2 *ent_xy;
    const int *prim_of_label;
    const unsigned char *own; // [nslot] slot is the primary slot of an owned generator
    int nslot;
    int *rowptr;          // row start
    unsigned char *rdeg;  // row degree
    int *col;
    double2 *v1, *v2;
    double *area;
    double2 *cen;
    unsigned long long *tile_state;
    int *flags;
    long long cap_nnz;
    // lane-refill kernel (lv_clip_fast.cu): parked rings, [ring slot][mesh slot]
    double2 *park_v;
    int *park_l;
    unsigned long long *park_nxt;
    int *park_hdr;
    int force_anomaly; // test hook (LV_CLIP_FORCE_ANOMALY=1): pretend some polygons are not generic
    // phase A of the linked-slot kernel: generators further than wreach from the lines x = wlo / whi (the domain edges)
    // cannot meet a candidate across the periodic seam
    double2 wlo, whi;
    double wreach;
};


#define TS_AGG (1ull << 62)
#define TS_INC (2ull << 62)
#define TS_MASK ((1ull << 62) - 1)

// flags[LVF_OVERFLOW] bits
#define OVF_POLY 1   // a polygon outgrew the kernel's edge capacity
#define OVF_NNZ 2    // the CSR edge buffers are too small
#define OVF_ANOMALY 4 // the fast kernel met a case it does not handle exactly (rerun with the edge-list kernel)

// CSR column of a candidate slot l: always the PRIMARY slot of its generator (an image entry only lends its position
// to the clipping).  Ghost generators (multi-GPU) have a primary slot on this rank as well, and that is the one slot per
// ghost the halo exchange fills.
__device__ __forceinline__ int lv_col_of(const ClipArgs &a, int l) {
    const unsigned e = a.ent_label[l];
    if (!(e & LV_IMAGE_BIT)) return l;
    const int p = a.prim_of_label[e & ~LV_IMAGE_BIT];
    return p >= 0 ? p : l;
}

int lv_clip_launch_fast(LvContext *c, const ClipArgs &a, int level); // lv_clip_fast.cu
