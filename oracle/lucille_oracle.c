/*
 * lucille_oracle.c -- TEST INFRASTRUCTURE, not product code.  See lucille_oracle.h.
 *
 * CPU restatement (plain C) of the lucille hot path.  Written from the behaviour of the
 * reference, function by function, with the reference file:line each piece follows.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).
 */
#include <assert.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lucille_oracle.h"

#define ORC_STACK 128          /* reference reserves 101 entries, bvh.c:80,126 */

typedef struct {
    double   bmin[3], bmax[3];
    uint64_t index;            /* tri_bbox_t, bvh.c:107-114 */
} tbox_t;

struct orc_tree {
    int         empty;
    uint64_t    ntris;
    double      bmin[3], bmax[3];
    orc_node_t *nodes;
    int64_t     nnodes, cap;
    int         max_depth;
    double     *tri_xyz;       /* post-build order, [ntris][9] v0 v1 v2 */
    double     *tri_nrm;       /* post-build order, [ntris][9] n0 n1 n2, or NULL */
    double     *tri_col;       /* post-build order, [ntris][9] c0 c1 c2, or NULL */
    double     *tri_st;        /* post-build order, [ntris][6] st0 st1 st2, or NULL */
    uint8_t    *tri_flags;     /* post-build order: bit0 has colour, bit1 has st, bit2 inside; or NULL */
    uint32_t   *orig;          /* post-build position -> input triangle */
    /* per-precision views, built once after construction */
    double     *lbox64, *rbox64, *tri64;
    float      *lbox32, *rbox32, *tri32;
    float       smin32[3], smax32[3];
};

/* ------------------------------------------------------------------ build (double, as the reference) */

/* bvh.c:1697-1731 bbox_add_margin */
static void add_margin(double bmin[3], double bmax[3])
{
    int i;
    double margin[3];
    const double eps = ORC_EPS;
    for (i = 0; i < 3; i++) {
        double scale = bmax[i] - bmin[i];
        margin[i] = (scale < eps) ? eps : eps * scale;
    }
    for (i = 0; i < 3; i++) { bmin[i] -= margin[i]; bmax[i] += margin[i]; }
}

/* bvh.c:1870-1895 calc_bbox_of_triangles (vmin/vmax are `(a<b)?a:b` selects, vector.h) */
static void bbox_of_range(double bmin[3], double bmax[3], const tbox_t *b, uint64_t n)
{
    uint64_t i; int k;
    for (k = 0; k < 3; k++) { bmin[k] = b[0].bmin[k]; bmax[k] = b[0].bmax[k]; }
    for (i = 1; i < n; i++) {
        for (k = 0; k < 3; k++) {
            bmin[k] = (bmin[k] < b[i].bmin[k]) ? bmin[k] : b[i].bmin[k];
            bmax[k] = (bmax[k] > b[i].bmax[k]) ? bmax[k] : b[i].bmax[k];
        }
    }
}

/* bvh.c:1190-1208 calc_surface_area */
static double surface_area(const double bmin[3], const double bmax[3])
{
    double sa = (bmax[0] - bmin[0]) * (bmax[1] - bmin[1]) +
                (bmax[1] - bmin[1]) * (bmax[2] - bmin[2]) +
                (bmax[2] - bmin[2]) * (bmax[0] - bmin[0]);
    sa *= 2.0;
    return sa;
}

/* bvh.c:1210-1228 SAH: evaluated in double, *stored in a float* before it is compared */
static double sah_cost(int ns1, double left_area, int ns2, double right_area, double s)
{
    const float Taabb = 0.2f;
    const float Ttri  = 0.8f;
    float T;
    T = 2.0f * Taabb
      + (left_area / s) * (double)ns1 * Ttri
      + (right_area / s) * (double)ns2 * Ttri;
    return T;
}

typedef struct { uint32_t bin[2][3][ORC_BINS]; } binbuf_t;

/* bvh.c:1571-1692 bin_triangle_edge */
static void bin_edges(binbuf_t *bb, const double bmin[3], const double bmax[3], const tbox_t *b, uint64_t n)
{
    uint64_t i; int k;
    const double eps = 1.0e-14;
    const double binsize = (double)ORC_BINS;
    double size[3], invsize[3];
    for (k = 0; k < 3; k++) {
        size[k] = bmax[k] - bmin[k];
        invsize[k] = (size[k] > eps) ? binsize / size[k] : 0.0;
    }
    memset(bb, 0, sizeof(*bb));
    for (i = 0; i < n; i++) {
        for (k = 0; k < 3; k++) {
            double qmin = (b[i].bmin[k] - bmin[k]) * invsize[k];
            double qmax = (b[i].bmax[k] - bmin[k]) * invsize[k];
            uint32_t imin = (uint32_t)qmin;
            uint32_t imax = (uint32_t)qmax;
            if (imin >= ORC_BINS) imin = ORC_BINS - 1;
            if (imax >= ORC_BINS) imax = ORC_BINS - 1;
            bb->bin[0][k][imin]++;
            bb->bin[1][k][imax]++;
        }
    }
}

/* bvh.c:1230-1326 find_cut_from_bin */
static void find_cut(double *cut_pos, int *cut_axis, const binbuf_t *bb,
                     const double bmin[3], const double bmax[3], uint64_t ntris)
{
    int i, j, k;
    double min_cost = ORC_INFINITY, min_pos = 0.0;
    int    min_axis = 0;
    double bstep[3], sa_total;
    for (k = 0; k < 3; k++) bstep[k] = (bmax[k] - bmin[k]) / (double)ORC_BINS;
    sa_total = surface_area(bmin, bmax);

    for (j = 0; j < 3; j++) {
        uint64_t left = 0, right = ntris;
        double lmin[3], lmax[3], rmin[3], rmax[3];
        for (k = 0; k < 3; k++) { lmin[k] = rmin[k] = bmin[k]; lmax[k] = rmax[k] = bmax[k]; }
        for (i = 0; i < ORC_BINS - 1; i++) {
            double pos, cost;
            left  += bb->bin[0][j][i];
            right -= bb->bin[1][j][i];
            pos = bmin[j] + (i + 1) * bstep[j];
            lmax[j] = pos;
            rmin[j] = pos;
            cost = sah_cost((int)left, surface_area(lmin, lmax), (int)right, surface_area(rmin, rmax), sa_total);
            if (cost < min_cost) { min_cost = cost; min_axis = j; min_pos = pos; }
        }
    }
    *cut_axis = min_axis;
    *cut_pos  = min_pos;
}

typedef struct {
    orc_tree *T;
    tbox_t   *boxes, *boxes_buf;
    const double *tri_in;      /* input order */
} build_ctx;

static int64_t new_node(orc_tree *T)
{
    if (T->nnodes == T->cap) {
        T->cap = T->cap ? 2 * T->cap : 1024;
        T->nodes = (orc_node_t *)realloc(T->nodes, sizeof(orc_node_t) * (size_t)T->cap);
    }
    memset(&T->nodes[T->nnodes], 0, sizeof(orc_node_t));
    return T->nnodes++;
}

/* bvh.c:1328-1564 bvh_construct; returns the node's index in DFS preorder */
static int64_t construct(build_ctx *B, const double bmin[3], const double bmax[3],
                         uint64_t left, uint64_t right, int depth)
{
    orc_tree *T = B->T;
    uint64_t n = right - left, i;
    int64_t me = new_node(T);
    if (depth > T->max_depth) T->max_depth = depth;

    if (n <= ORC_LEAF_TRIS) {                                 /* bvh.c:1352-1403 */
        for (i = 0; i < n; i++) {                             /* gather_triangles, bvh.c:1897-1917 */
            uint64_t src = B->boxes[left + i].index;
            memcpy(T->tri_xyz + 9 * (left + i), B->tri_in + 9 * src, sizeof(double) * 9);
            T->orig[left + i] = (uint32_t)src;
        }
        T->nodes[me].is_leaf = 1;
        T->nodes[me].tri_start = (int64_t)left;
        T->nodes[me].ntris = (int64_t)n;
        T->nodes[me].child0 = T->nodes[me].child1 = -1;
        return me;
    }

    {
        binbuf_t bb;
        double cut_pos; int cut_axis;
        uint64_t nl = 0, nr = n - 1;
        double lmin[3], lmax[3], rmin[3], rmax[3];
        int64_t c0, c1;
        int k;

        bin_edges(&bb, bmin, bmax, B->boxes + left, n);
        find_cut(&cut_pos, &cut_axis, &bb, bmin, bmax, n);

        memcpy(B->boxes_buf, B->boxes + left, sizeof(tbox_t) * n);        /* bvh.c:1438 */
        for (i = 0; i < n; i++) {
            if (B->boxes_buf[i].bmax[cut_axis] < cut_pos) {               /* bvh.c:1442 */
                B->boxes[left + nl] = B->boxes_buf[i];
                nl++;
            } else {
                B->boxes[left + nr] = B->boxes_buf[i];                    /* right part fills from the back */
                nr--;
            }
        }
        if (nl == 0 || nl == n) nl = n / 2;                               /* bvh.c:1471-1478 */

        T->nodes[me].axis = cut_axis;

        bbox_of_range(lmin, lmax, B->boxes + left, nl);
        add_margin(lmin, lmax);
        for (k = 0; k < 3; k++) { T->nodes[me].lbox[k] = lmin[k]; T->nodes[me].lbox[3 + k] = lmax[k]; }
        c0 = construct(B, lmin, lmax, left, left + nl, depth + 1);
        T->nodes[me].child0 = c0;

        bbox_of_range(rmin, rmax, B->boxes + left + nl, n - nl);
        add_margin(rmin, rmax);
        for (k = 0; k < 3; k++) { T->nodes[me].rbox[k] = rmin[k]; T->nodes[me].rbox[3 + k] = rmax[k]; }
        c1 = construct(B, rmin, rmax, left + nl, right, depth + 1);
        T->nodes[me].child1 = c1;
    }
    return me;
}

static float f32_down(double x) { float f = (float)x; if ((double)f > x) f = nextafterf(f, -INFINITY); return f; }
static float f32_up(double x)   { float f = (float)x; if ((double)f < x) f = nextafterf(f,  INFINITY); return f; }

/* Per-precision views.  double: the reference's own numbers.  float: boxes rounded OUTWARD
 * (min down, max up) so fp32 culling stays conservative w.r.t. the double tree; vertices rounded
 * to nearest; e1/e2 subtracted in fp32 (the same subtraction triangle_isect does per test). */
static void make_views(orc_tree *T)
{
    int64_t i; uint64_t p; int k;
    T->lbox64 = (double *)malloc(sizeof(double) * 6 * (size_t)T->nnodes);
    T->rbox64 = (double *)malloc(sizeof(double) * 6 * (size_t)T->nnodes);
    T->lbox32 = (float *)malloc(sizeof(float) * 6 * (size_t)T->nnodes);
    T->rbox32 = (float *)malloc(sizeof(float) * 6 * (size_t)T->nnodes);
    T->tri64  = (double *)malloc(sizeof(double) * 9 * (size_t)T->ntris);
    T->tri32  = (float *)malloc(sizeof(float) * 9 * (size_t)T->ntris);
    for (i = 0; i < T->nnodes; i++) {
        const orc_node_t *n = &T->nodes[i];
        for (k = 0; k < 6; k++) { T->lbox64[6 * i + k] = n->lbox[k]; T->rbox64[6 * i + k] = n->rbox[k]; }
        for (k = 0; k < 3; k++) {
            T->lbox32[6 * i + k]     = f32_down(n->lbox[k]);
            T->lbox32[6 * i + 3 + k] = f32_up(n->lbox[3 + k]);
            T->rbox32[6 * i + k]     = f32_down(n->rbox[k]);
            T->rbox32[6 * i + 3 + k] = f32_up(n->rbox[3 + k]);
        }
    }
    for (p = 0; p < T->ntris; p++) {
        const double *v = T->tri_xyz + 9 * p;
        double *d = T->tri64 + 9 * p;
        float  *f = T->tri32 + 9 * p;
        float v0[3], v1[3], v2[3];
        for (k = 0; k < 3; k++) {
            d[k] = v[k]; d[3 + k] = v[3 + k] - v[k]; d[6 + k] = v[6 + k] - v[k];
            v0[k] = (float)v[k]; v1[k] = (float)v[3 + k]; v2[k] = (float)v[6 + k];
            f[k] = v0[k]; f[3 + k] = v1[k] - v0[k]; f[6 + k] = v2[k] - v0[k];
        }
    }
    for (k = 0; k < 3; k++) { T->smin32[k] = f32_down(T->bmin[k]); T->smax32[k] = f32_up(T->bmax[k]); }
}

/* bvh.c:276-379 ri_bvh_build (+ create_triangle_list 1736-1826, calc_scene_bbox 1829-1849) */
orc_tree *orc_build(const double *tri_xyz, uint64_t ntris)
{
    orc_tree *T = (orc_tree *)calloc(1, sizeof(orc_tree));
    build_ctx B;
    uint64_t i; int k, c;

    T->ntris = ntris;
    if (ntris == 0) { T->empty = 1; return T; }               /* bvh.c:311-315 */

    B.T = T; B.tri_in = tri_xyz;
    B.boxes     = (tbox_t *)malloc(sizeof(tbox_t) * ntris);
    B.boxes_buf = (tbox_t *)malloc(sizeof(tbox_t) * ntris);
    T->tri_xyz  = (double *)malloc(sizeof(double) * 9 * ntris);
    T->orig     = (uint32_t *)malloc(sizeof(uint32_t) * ntris);

    for (i = 0; i < ntris; i++) {                             /* get_bbox_of_triangle, bvh.c:1852-1868 */
        const double *v = tri_xyz + 9 * i;
        for (k = 0; k < 3; k++) {
            double mn = v[k], mx = v[k];
            for (c = 1; c < 3; c++) {
                mn = (mn < v[3 * c + k]) ? mn : v[3 * c + k];
                mx = (mx > v[3 * c + k]) ? mx : v[3 * c + k];
            }
            B.boxes[i].bmin[k] = mn; B.boxes[i].bmax[k] = mx;
        }
        B.boxes[i].index = i;
    }
    bbox_of_range(T->bmin, T->bmax, B.boxes, ntris);          /* calc_scene_bbox */
    add_margin(T->bmin, T->bmax);                             /* bvh.c:330 */

    construct(&B, T->bmin, T->bmax, 0, ntris, 0);

    free(B.boxes); free(B.boxes_buf);
    make_views(T);
    return T;
}

void orc_free(orc_tree *T)
{
    if (!T) return;
    free(T->nodes); free(T->tri_xyz); free(T->orig); free(T->tri_nrm); free(T->tri_col); free(T->tri_st); free(T->tri_flags);
    free(T->lbox64); free(T->rbox64); free(T->tri64);
    free(T->lbox32); free(T->rbox32); free(T->tri32);
    free(T);
}

void orc_set_normals(orc_tree *T, const double *tri_normals)
{
    uint64_t p;
    free(T->tri_nrm); T->tri_nrm = NULL;
    if (!tri_normals || T->empty) return;
    T->tri_nrm = (double *)malloc(sizeof(double) * 9 * T->ntris);
    for (p = 0; p < T->ntris; p++) memcpy(T->tri_nrm + 9 * p, tri_normals + 9 * (size_t)T->orig[p], sizeof(double) * 9);
}

int     orc_is_empty(const orc_tree *T)  { return T->empty; }
int64_t orc_num_nodes(const orc_tree *T) { return T->nnodes; }
int     orc_max_depth(const orc_tree *T) { return T->max_depth; }
int64_t orc_get_nodes(const orc_tree *T, orc_node_t *out)
{
    if (T->nnodes) memcpy(out, T->nodes, sizeof(orc_node_t) * (size_t)T->nnodes);
    return T->nnodes;
}
void orc_get_triorder(const orc_tree *T, uint32_t *orig)
{
    if (T->ntris && !T->empty) memcpy(orig, T->orig, sizeof(uint32_t) * T->ntris);
}
void orc_scene_bbox(const orc_tree *T, double *bmin, double *bmax)
{
    int k; for (k = 0; k < 3; k++) { bmin[k] = T->bmin[k]; bmax[k] = T->bmax[k]; }
}

/* ------------------------------------------------------------------ traversal, both precisions */

#define REAL double
#define SFX(x) x##_f64
#define R(x) x
#define REAL_MAX DBL_MAX
#define RFABS fabs
#define RSQRT sqrt
#include "oracle_trav.inc"
#undef REAL
#undef SFX
#undef R
#undef REAL_MAX
#undef RFABS
#undef RSQRT

#define REAL float
#define SFX(x) x##_f32
#define R(x) x##f
#define REAL_MAX FLT_MAX
#define RFABS fabsf
#define RSQRT sqrtf
#include "oracle_trav.inc"
#undef REAL
#undef SFX
#undef R
#undef REAL_MAX
#undef RFABS
#undef RSQRT

static void view64(const orc_tree *T, view_t_f64 *V)
{
    int k;
    V->lbox = T->lbox64; V->rbox = T->rbox64; V->tri = T->tri64;
    for (k = 0; k < 3; k++) { V->smin[k] = T->bmin[k]; V->smax[k] = T->bmax[k]; }
}
static void view32(const orc_tree *T, view_t_f32 *V)
{
    int k;
    V->lbox = T->lbox32; V->rbox = T->rbox32; V->tri = T->tri32;
    for (k = 0; k < 3; k++) { V->smin[k] = T->smin32[k]; V->smax[k] = T->smax32[k]; }
}

void orc_intersect_f64(const orc_tree *T, const double *rays, uint64_t n, orc_hit_f64 *out, orc_counters_t *c)
{
    view_t_f64 V = {0}; uint64_t i;
    if (!T->empty) view64(T, &V);
    for (i = 0; i < n; i++) {
        double t = 1.0e38, u = 0, v = 0; uint32_t prim = ORC_MISS_PRIM;
        int hit = trace_f64(T, &V, rays + 6 * i, rays + 6 * i + 3, 0, &t, &u, &v, &prim, c);
        out[i].t = hit ? t : 1.0e38; out[i].u = hit ? u : 0.0; out[i].v = hit ? v : 0.0;
        out[i].prim = hit ? prim : ORC_MISS_PRIM; out[i].hit = (uint32_t)hit;
    }
}

void orc_intersect_f32(const orc_tree *T, const float *rays, uint64_t n, orc_hit_f32 *out, orc_counters_t *c)
{
    view_t_f32 V = {0}; uint64_t i;
    if (!T->empty) view32(T, &V);
    for (i = 0; i < n; i++) {
        float t = 1.0e38f, u = 0, v = 0; uint32_t prim = ORC_MISS_PRIM;
        int hit = trace_f32(T, &V, rays + 8 * i, rays + 8 * i + 4, 0, &t, &u, &v, &prim, c);
        out[i].t = hit ? t : 1.0e38f; out[i].u = hit ? u : 0.0f; out[i].v = hit ? v : 0.0f;
        out[i].prim = hit ? prim : ORC_MISS_PRIM;
    }
}

void orc_occluded_f64(const orc_tree *T, const double *rays, uint64_t n, uint8_t *out, orc_counters_t *c)
{
    view_t_f64 V = {0}; uint64_t i;
    if (!T->empty) view64(T, &V);
    for (i = 0; i < n; i++) {
        double t, u, v; uint32_t prim;
        out[i] = (uint8_t)trace_f64(T, &V, rays + 6 * i, rays + 6 * i + 3, 1, &t, &u, &v, &prim, c);
    }
}

void orc_occluded_f32(const orc_tree *T, const float *rays, uint64_t n, uint8_t *out, orc_counters_t *c)
{
    view_t_f32 V = {0}; uint64_t i;
    if (!T->empty) view32(T, &V);
    for (i = 0; i < n; i++) {
        float t, u, v; uint32_t prim;
        out[i] = (uint8_t)trace_f32(T, &V, rays + 8 * i, rays + 8 * i + 4, 1, &t, &u, &v, &prim, c);
    }
}

/* ------------------------------------------------------------------ hit state (double) */

/* intersection_state.c:99-248 for geometry carrying only "P": Ns = Ng, tangent/binormal from ri_ortho_basis(Ng).
 * Ng = normalize((v1-v0) x (v2-v0)), base/geometric.c:20-33.  No face-forwarding. */
static void state_build_uv(const orc_tree *T, const double org[3], const double dir[3], double t, double bu, double bv, uint32_t prim,
                           orc_state_f64 *s)
{
    const double *v = T->tri_xyz + 9 * (size_t)prim;
    double v01[3], v02[3], basis[3][3];
    int k;
    for (k = 0; k < 3; k++) s->P[k] = org[k] + dir[k] * t;
    for (k = 0; k < 3; k++) { v01[k] = v[3 + k] - v[k]; v02[k] = v[6 + k] - v[k]; }
    cross_f64(s->Ng, v01, v02);
    normalize_f64(s->Ng);
    /* a triangle whose nine normal components are all zero belongs to a geom without normals (geom->normals == NULL) */
    const double *nn = T->tri_nrm ? T->tri_nrm + 9 * (size_t)prim : NULL;
    int has_n = 0;
    if (nn) for (k = 0; k < 9; k++) if (nn[k] != 0.0) has_n = 1;
    if (has_n) {                                              /* ri_lerp_vector, geometric.c:40-62 */
        const double *n = nn;
        const double w0 = 1.0 - bu - bv;
        for (k = 0; k < 3; k++) {
            const double a = n[k] * w0, b = n[3 + k] * bu, c = n[6 + k] * bv;
            s->Ns[k] = (a + b) + c;
        }
    } else {
        for (k = 0; k < 3; k++) s->Ns[k] = s->Ng[k];
    }
    ortho_basis_f64(basis, s->Ng);
    for (k = 0; k < 3; k++) { s->tangent[k] = basis[0][k]; s->binormal[k] = basis[1][k]; }
}

static void state_build(const orc_tree *T, const double org[3], const double dir[3], double t, uint32_t prim, orc_state_f64 *s)
{
    double *saved = ((orc_tree *)T)->tri_nrm;          /* callers without (u, v): geometric normal only (path tracer uses Ng) */
    ((orc_tree *)T)->tri_nrm = NULL;
    state_build_uv(T, org, dir, t, 0.0, 0.0, prim, s);
    ((orc_tree *)T)->tri_nrm = saved;
}

void orc_set_attributes(orc_tree *T, const double *colors, const uint8_t *has_color, const double *st, const uint8_t *has_st,
                        const uint8_t *inside)
{
    uint64_t p;
    free(T->tri_col); free(T->tri_st); free(T->tri_flags);
    T->tri_col = NULL; T->tri_st = NULL; T->tri_flags = NULL;
    if (T->empty) return;
    T->tri_col = (double *)calloc(9 * T->ntris, sizeof(double));
    T->tri_st = (double *)calloc(6 * T->ntris, sizeof(double));
    T->tri_flags = (uint8_t *)calloc(T->ntris, 1);
    for (p = 0; p < T->ntris; p++) {
        const size_t o = (size_t)T->orig[p];
        if (colors && has_color && has_color[o]) { memcpy(T->tri_col + 9 * p, colors + 9 * o, sizeof(double) * 9); T->tri_flags[p] |= 1; }
        if (st && has_st && has_st[o]) { memcpy(T->tri_st + 6 * p, st + 6 * o, sizeof(double) * 6); T->tri_flags[p] |= 2; }
        if (inside && inside[o]) T->tri_flags[p] |= 4;
    }
}

void orc_state_ext_build_f64(const orc_tree *T, const double *rays, const orc_hit_f64 *hits, uint64_t n, orc_state_ext_f64 *out)
{
    uint64_t i;
    int k;
    for (i = 0; i < n; i++) {
        const double *r = rays + 6 * i;
        orc_state_ext_f64 *o = &out[i];
        memset(o, 0, sizeof(*o));
        o->hit = (int32_t)hits[i].hit;
        if (!hits[i].hit) continue;
        {
            const uint32_t p = hits[i].prim;
            const double u = hits[i].u, v = hits[i].v;
            const uint8_t fl = T->tri_flags ? T->tri_flags[p] : 0;
            double d[3] = { r[3], r[4], r[5] };
            for (k = 0; k < 3; k++) o->E[k] = r[k];                          /* intersection_state.c:133 */
            normalize_f64(d);                                                /* :130-131 */
            for (k = 0; k < 3; k++) o->I[k] = d[k];
            if (fl & 1) {                                                    /* ri_lerp_vector, geometric.c:40-62 */
                const double *c = T->tri_col + 9 * (size_t)p;
                const double w0 = 1.0 - u - v;
                for (k = 0; k < 3; k++) { const double a = c[k] * w0, b = c[3 + k] * u, cc = c[6 + k] * v; o->color[k] = (a + b) + cc; }
            } else {
                for (k = 0; k < 3; k++) o->color[k] = 1.0;                   /* :204-207 */
            }
            if (fl & 2) {                                                    /* lerp_uv, :266-280 */
                const double *t = T->tri_st + 6 * (size_t)p;
                o->st[0] = (1 - u - v) * t[0] + u * t[2] + v * t[4];
                o->st[1] = (1 - u - v) * t[1] + u * t[3] + v * t[5];
            }
            o->t = hits[i].t;
            o->inside = (fl & 4) ? 1 : 0;                                    /* :233-246 */
        }
    }
}

void orc_state_build_f64(const orc_tree *T, const double *rays, const orc_hit_f64 *hits, uint64_t n, orc_state_f64 *out)
{
    uint64_t i;
    for (i = 0; i < n; i++) {
        if (hits[i].hit) state_build_uv(T, rays + 6 * i, rays + 6 * i + 3, hits[i].t, hits[i].u, hits[i].v, hits[i].prim, &out[i]);
        else memset(&out[i], 0, sizeof(out[i]));
    }
}

/* ------------------------------------------------------------------ MT19937 (random.c) */

typedef struct { uint32_t mt[624]; int mti; } mt_t;

/* random.c:98-112 seedMT2: mt[i] = 69069 * mt[i-1] (1998 seeding) */
static void mt_seed(mt_t *m, uint32_t seed)
{
    int i;
    m->mt[0] = seed;
    for (i = 1; i < 624; i++) m->mt[i] = 69069u * m->mt[i - 1];
    m->mti = 624;
}

/* random.c:211-247 randomMT2 (integer part) */
static uint32_t mt_next_u32(mt_t *m)
{
    static const uint32_t mag01[2] = { 0x0u, 0x9908b0dfu };
    uint32_t y;
    if (m->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (m->mt[kk] & 0x80000000u) | (m->mt[kk + 1] & 0x7fffffffu);
            m->mt[kk] = m->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1];
        }
        for (; kk < 623; kk++) {
            y = (m->mt[kk] & 0x80000000u) | (m->mt[kk + 1] & 0x7fffffffu);
            m->mt[kk] = m->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1];
        }
        y = (m->mt[623] & 0x80000000u) | (m->mt[0] & 0x7fffffffu);
        m->mt[623] = m->mt[396] ^ (y >> 1) ^ mag01[y & 1];
        m->mti = 0;
    }
    y = m->mt[m->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_next(mt_t *m) { return (double)mt_next_u32(m) * 2.3283064365386963e-10; }

void orc_mt_stream(uint32_t seed, uint64_t n, double *out)
{
    mt_t m; uint64_t i;
    mt_seed(&m, seed);
    for (i = 0; i < n; i++) out[i] = mt_next(&m);
}

void orc_mt_stream_u32(uint32_t seed, uint64_t n, uint32_t *out)
{
    mt_t m; uint64_t i;
    mt_seed(&m, seed);
    for (i = 0; i < n; i++) out[i] = mt_next_u32(&m);
}

/* ------------------------------------------------------------------ pixel loop */

/* spiral.c:97-140 NthBucketSpiral */
static void nth_bucket_spiral(int n, int nxb, int nyb, int *bx, int *by)
{
    int nx, ny, nxny, minnxny, x, y;
    int minnb = (nxb < nyb) ? nxb : nyb;
    int center = (minnb - 1) / 2;
    nx = nxb; ny = nyb;
    while (n < nx * ny) { nx = nx - 1; ny = ny - 1; }
    nxny = nx * ny;
    minnxny = (nx < ny) ? nx : ny;
    if (minnxny % 2 == 1) {
        if (n <= (nxny + ny)) { x = nx - minnxny / 2; y = -minnxny / 2 + n - nxny; }
        else { x = nx - minnxny / 2 - (n - (nxny + ny)); y = ny - minnxny / 2; }
    } else {
        if (n <= (nxny + ny)) { x = -minnxny / 2; y = ny - minnxny / 2 - (n - nxny); }
        else { x = -minnxny / 2 + (n - (nxny + ny)); y = -minnxny / 2; }
    }
    *bx = x + center; *by = y + center;
}

/* render.c:582-710 create_bucket_list with BUCKET_ORDER_SPIRAL */
int orc_bucket_list(int width, int height, int bucket_size, int32_t *out, int max_buckets)
{
    int nxb = width / bucket_size + ((width % bucket_size) ? 1 : 0);
    int nyb = height / bucket_size + ((height % bucket_size) ? 1 : 0);
    int wrem = width % bucket_size, hrem = height % bucket_size;
    int n, nb = nxb * nyb;
    if (nb > max_buckets) return -nb;
    for (n = 0; n < nb; n++) {
        int bx, by, w, h;
        nth_bucket_spiral(n, nxb, nyb, &bx, &by);
        w = (bx == nxb - 1 && wrem) ? wrem : bucket_size;
        h = (by == nyb - 1 && hrem) ? hrem : bucket_size;
        out[4 * n + 0] = bx * bucket_size; out[4 * n + 1] = by * bucket_size;
        out[4 * n + 2] = w; out[4 * n + 3] = h;
    }
    return nb;
}

/* render.c:870-917 init_sigma (one table) */
static void sigma_table(unsigned int period, unsigned int *sigma)
{
    unsigned int i, inverse, digit, bits;
    for (i = 0; i < period; i++) {
        digit = period; inverse = 0;
        for (bits = i; bits; bits >>= 1) {
            digit >>= 1;
            if (bits & 1) inverse += digit;
        }
        sigma[i] = inverse;
    }
}

/* render.c:830-861 sample_subpixel (note: periodx masks both indices) */
void orc_subpixel_jitter(int xs, int ys, int xsamples, int ysamples, double *jx, double *jy)
{
    unsigned int sx[256], sy[256];
    unsigned int periodx = (unsigned int)xsamples, periody = (unsigned int)ysamples;
    unsigned int j, k;
    double a, b;
    sigma_table(periodx, sx);
    sigma_table(periody, sy);
    j = (unsigned int)xs & (periodx - 1);
    k = (unsigned int)ys & (periodx - 1);
    a = (double)xs + (double)sx[k] / (double)periodx;
    b = (double)ys + (double)sy[j] / (double)periody;
    a /= (double)xsamples;
    b /= (double)ysamples;
    a += 0.5 / (xsamples * xsamples);
    b += 0.5 / (ysamples * ysamples);
    *jx = a; *jy = b;
}

/* camera.c:248-352 (perspective, flength > 0) + render.c:770-781 */
void orc_camera_ray(const orc_frame_t *f, double x, double y, double *org, double *dir)
{
    double v[4], pos[4], dirpos[4];
    double w = f->width, h = f->height;
    float sign = f->is_rh ? -1.0 : 1.0;
    int i, j;
    v[0] = (2.0f * x - w) / w;
    v[1] = (2.0f * y - h) / h;
    v[2] = sign * f->flength;
    v[3] = 1.0;
    for (j = 0; j < 4; j++) {                  /* ri_vector_transform(pos, (0,0,0,1), c2w), vector.h:182-210 */
        double o[4] = { 0.0, 0.0, 0.0, 1.0 };
        pos[j] = 0.0;
        for (i = 0; i < 4; i++) pos[j] += o[i] * f->c2w[4 * i + j];
    }
    for (j = 0; j < 4; j++) {
        dirpos[j] = 0.0;
        for (i = 0; i < 4; i++) dirpos[j] += v[i] * f->c2w[4 * i + j];
    }
    for (j = 0; j < 3; j++) { org[j] = pos[j]; dir[j] = dirpos[j] - pos[j]; }
    normalize_f64(dir);
}

/* ambientocclusion.c:42-151 calculate_occlusion */
static double ao_radiance(const orc_tree *T, const view_t_f64 *V, const orc_state_f64 *s,
                          int ntheta, int nphi, mt_t *rng, uint64_t *nrays)
{
    double basis[3][3], org[3], dirl[3], dir[3];
    const double eps = 1.0e-6;
    double occlusion = 0.0, nsamples;
    uint32_t i, j; int k;

    ortho_basis_f64(basis, s->Ns);
    for (k = 0; k < 3; k++) org[k] = s->P[k];
    for (k = 0; k < 3; k++) org[k] += s->Ns[k] * eps;

    for (j = 0; j < (uint32_t)nphi; j++) {
        for (i = 0; i < (uint32_t)ntheta; i++) {
            double z0 = (i + mt_next(rng)) / (double)ntheta;
            double z1 = (j + mt_next(rng)) / (double)nphi;
            double cos_theta = sqrt(z0);
            double phi = 2.0 * M_PI * z1;
            double t, u, v; uint32_t prim;
            dirl[0] = cos(phi) * cos_theta;
            dirl[1] = sin(phi) * cos_theta;
            dirl[2] = sqrt(1.0 - cos_theta * cos_theta);
            for (k = 0; k < 3; k++)
                dir[k] = dirl[0] * basis[0][k] + dirl[1] * basis[1][k] + dirl[2] * basis[2][k];
            (*nrays)++;
            if (trace_f64(T, V, org, dir, 0, &t, &u, &v, &prim, NULL)) occlusion += 1.0;
        }
    }
    nsamples = ntheta * nphi;
    return 1.0 * (nsamples - occlusion) / nsamples;
}

/* render.c:1107-1146 render_bucket, 715-823 subsample, 919-979 bucket_write;
 * ambientocclusion.c:332-415 ri_transport_ambientocclusion (no sun-sky, no texture). */
void orc_render_ao(const orc_tree *T, const orc_frame_t *f, float *rgb, uint64_t *nrays_out)
{
    int nb_max = (f->width / f->bucket_size + 1) * (f->height / f->bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(f->width, f->height, f->bucket_size, buckets, nb_max);
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t nrays = 0;
    int b;

    if (!T->empty) view64(T, &V);
    mt_seed(&rng, 4357);                                       /* random.c:221 */

    for (b = 0; b < nb; b++) {
        int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        int u, v;
        for (v = by; v < by + bh; v++) {
            for (u = bx; u < bx + bw; u++) {
                double accum = 0.0, px;
                int xs, ys;
                for (ys = 0; ys < f->ysamples; ys++) {
                    for (xs = 0; xs < f->xsamples; xs++) {
                        double jx, jy, org[3], dir[3], t, uu, vv, rad = 0.0;
                        uint32_t prim;
                        orc_subpixel_jitter(xs, ys, f->xsamples, f->ysamples, &jx, &jy);
                        orc_camera_ray(f, (double)(u + jx), (double)(v + jy), org, dir);
                        nrays++;
                        if (trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
                            orc_state_f64 s;
                            state_build_uv(T, org, dir, t, uu, vv, prim, &s);
                            rad = ao_radiance(T, &V, &s, f->ntheta, f->nphi, &rng, &nrays);
                        }
                        accum = accum + rad;
                    }
                }
                px = accum * ((double)1.0 / (f->xsamples * f->ysamples));
                {
                    float *dst = rgb + 3 * ((size_t)(f->height - v - 1) * f->width + u);
                    dst[0] = dst[1] = dst[2] = (float)px;
                }
            }
        }
    }
    free(buckets);
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ material texture (row a11, ambientocclusion.c:393-401) */

/* render/texture.c:86-236 with USE_ZORDER 0 */
static void tex_fetch(const float *data, int width, int height, double u, double v, double color_out[4])
{
    int i, idx, x, y;
    double sx, sy, w[4], texel[4][4], px, py, dx, dy;
    sx = floor(u); sy = floor(v);
    u = u - sx; v = v - sy;
    if (u < 0.0) u = 0.0;
    if (u >= 1.0) u = 1.0;
    if (v < 0.0) v = 0.0;
    if (v >= 1.0) v = 1.0;
    px = u * (width - 1);
    py = v * (height - 1);
    x = (int)px; y = (int)py;
    dx = px - x; dy = py - y;
    w[0] = (1.0 - dx) * (1.0 - dy);
    w[1] = (1.0 - dx) * dy;
    w[2] = dx * (1.0 - dy);
    w[3] = dx * dy;
    idx = y * width + x;
    for (i = 0; i < 4; i++) texel[0][i] = data[4 * idx + i];
    for (i = 0; i < 4; i++) { texel[1][i] = 0.0; texel[2][i] = 0.0; texel[3][i] = 0.0; }
    if (y < height - 1 && x < width - 1) {
        idx = (y + 1) * width + x;     for (i = 0; i < 4; i++) texel[1][i] = data[4 * idx + i];
        idx = y * width + x + 1;       for (i = 0; i < 4; i++) texel[2][i] = data[4 * idx + i];
        idx = (y + 1) * width + x + 1; for (i = 0; i < 4; i++) texel[3][i] = data[4 * idx + i];
    } else if (y < height - 1) {
        idx = (y + 1) * width + x;     for (i = 0; i < 4; i++) texel[1][i] = data[4 * idx + i];
    } else if (x < width - 1) {
        idx = y * width + x + 1;       for (i = 0; i < 4; i++) texel[2][i] = data[4 * idx + i];
    }
    for (i = 0; i < 4; i++)
        color_out[i] = (double)(w[0] * texel[0][i] + w[1] * texel[1][i] + w[2] * texel[2][i] + w[3] * texel[3][i]);
}

void orc_texture_fetch(const float *rgba, int width, int height, const double *uv, uint64_t n, double *out4)
{
    uint64_t i;
    for (i = 0; i < n; i++) tex_fetch(rgba, width, height, uv[2 * i], uv[2 * i + 1], out4 + 4 * i);
}

void orc_render_ao_textured(const orc_tree *T, const orc_frame_t *f, const float *rgba, int tw, int th, float *rgb, uint64_t *nrays_out)
{
    int nb_max = (f->width / f->bucket_size + 1) * (f->height / f->bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(f->width, f->height, f->bucket_size, buckets, nb_max);
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t nrays = 0;
    int b, k;

    if (!T->empty) view64(T, &V);
    mt_seed(&rng, 4357);
    for (b = 0; b < nb; b++) {
        int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        int u, v;
        for (v = by; v < by + bh; v++) {
            for (u = bx; u < bx + bw; u++) {
                double accum[3] = {0.0, 0.0, 0.0};
                int xs, ys;
                float *dst = rgb + 3 * ((size_t)(f->height - v - 1) * f->width + u);
                for (ys = 0; ys < f->ysamples; ys++) {
                    for (xs = 0; xs < f->xsamples; xs++) {
                        double jx, jy, org[3], dir[3], t, uu, vv, rad[3] = {0.0, 0.0, 0.0};
                        uint32_t prim;
                        orc_subpixel_jitter(xs, ys, f->xsamples, f->ysamples, &jx, &jy);
                        orc_camera_ray(f, (double)(u + jx), (double)(v + jy), org, dir);
                        nrays++;
                        if (trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
                            orc_state_f64 st;
                            double lo, s0 = 0.0, s1 = 0.0, texcol[4];
                            state_build_uv(T, org, dir, t, uu, vv, prim, &st);
                            lo = ao_radiance(T, &V, &st, f->ntheta, f->nphi, &rng, &nrays);
                            if (T->tri_flags && (T->tri_flags[prim] & 2)) {                  /* lerp_uv, intersection_state.c:266-280 */
                                const double *c = T->tri_st + 6 * (size_t)prim;
                                s0 = (1 - uu - vv) * c[0] + uu * c[2] + vv * c[4];
                                s1 = (1 - uu - vv) * c[1] + uu * c[3] + vv * c[5];
                            }
                            tex_fetch(rgba, tw, th, s0, s1, texcol);
                            for (k = 0; k < 3; k++) { rad[k] = lo; rad[k] *= texcol[k]; }     /* ambientocclusion.c:398-400 */
                        }
                        for (k = 0; k < 3; k++) accum[k] = accum[k] + rad[k];
                    }
                }
                for (k = 0; k < 3; k++) dst[k] = (float)(accum[k] * ((double)1.0 / (f->xsamples * f->ysamples)));
            }
        }
    }
    free(buckets);
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ dirt-map transport */

/* dirtmap.c:84-221 calculate_dirt(4, 4) */
static double dirt_radiance(const orc_tree *T, const view_t_f64 *V, const orc_state_f64 *s, mt_t *rng, uint64_t *nrays)
{
    const uint32_t ntheta = 4, nphi = 4;
    const double eps = 1.0e-5, dirt_gain = 1.0f, near_clip = 0.1, far_clip = 0.5;
    const double dirt_color = 0.0, base_color = 1.0;           /* black / white on all three channels */
    double basis[3][3], org[3], dirl[3], dir[3], sum_color = 0.0, nsamples;
    uint32_t i, j; int k;

    ortho_basis_f64(basis, s->Ns);
    for (k = 0; k < 3; k++) org[k] = s->P[k];
    for (k = 0; k < 3; k++) org[k] += s->Ns[k] * eps;
    for (j = 0; j < nphi; j++) {
        for (i = 0; i < ntheta; i++) {
            double z0 = (i + mt_next(rng)) / (double)ntheta;
            double z1 = (j + mt_next(rng)) / (double)nphi;
            double cos_theta = sqrt(z0);
            double phi = 2.0 * M_PI * z1;
            double t, u, v; uint32_t prim;
            dirl[0] = cos(phi) * cos_theta;
            dirl[1] = sin(phi) * cos_theta;
            dirl[2] = sqrt(1.0 - cos_theta * cos_theta);
            for (k = 0; k < 3; k++)
                dir[k] = dirl[0] * basis[0][k] + dirl[1] * basis[1][k] + dirl[2] * basis[2][k];
            (*nrays)++;
            if (trace_f64(T, V, org, dir, 0, &t, &u, &v, &prim, NULL)) {
                if (t <= near_clip) {
                    sum_color = sum_color + dirt_color;
                } else if (t >= far_clip) {
                    sum_color = sum_color + base_color;
                } else {                                       /* mix_color, dirtmap.c:70-82 */
                    double p = pow(1.0 - ((t - near_clip) / (far_clip - near_clip)), 1.0f / dirt_gain);
                    double col;
                    if (p < 0.0) p = 0.0;
                    if (p > 1.0) p = 1.0;
                    col = (1.0 - p) * base_color - p * dirt_color;
                    sum_color = sum_color + col;
                }
            } else {
                sum_color = sum_color + base_color;
            }
        }
    }
    nsamples = ntheta * nphi;
    return sum_color / nsamples;
}

void orc_transport_batch(const orc_tree *T, int which, int ntheta, int nphi, const double *rays, uint64_t n, double *radiance3)
{
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t i, nrays = 0;
    if (!T->empty) view64(T, &V);
    mt_seed(&rng, 4357);
    for (i = 0; i < n; i++) {
        const double *r = rays + 6 * i;
        double org[3] = { r[0], r[1], r[2] }, dir[3] = { r[3], r[4], r[5] }, t, uu, vv, rad = 0.0;
        uint32_t prim;
        if (trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
            orc_state_f64 st;
            state_build_uv(T, org, dir, t, uu, vv, prim, &st);
            if (which == 3) rad = 1.0;                                       /* transport.c:150: white */
            else rad = which == 1 ? dirt_radiance(T, &V, &st, &rng, &nrays) : ao_radiance(T, &V, &st, ntheta, nphi, &rng, &nrays);
        }
        radiance3[3 * i] = radiance3[3 * i + 1] = radiance3[3 * i + 2] = rad;
    }
}

void orc_render_dirtmap(const orc_tree *T, const orc_frame_t *f, float *rgb, uint64_t *nrays_out)
{
    int nb_max = (f->width / f->bucket_size + 1) * (f->height / f->bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(f->width, f->height, f->bucket_size, buckets, nb_max);
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t nrays = 0;
    int b;

    if (!T->empty) view64(T, &V);
    mt_seed(&rng, 4357);
    for (b = 0; b < nb; b++) {
        int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        int u, v;
        for (v = by; v < by + bh; v++) {
            for (u = bx; u < bx + bw; u++) {
                double accum = 0.0, px;
                int xs, ys;
                for (ys = 0; ys < f->ysamples; ys++) {
                    for (xs = 0; xs < f->xsamples; xs++) {
                        double jx, jy, org[3], dir[3], t, uu, vv, rad = 0.0;
                        uint32_t prim;
                        orc_subpixel_jitter(xs, ys, f->xsamples, f->ysamples, &jx, &jy);
                        orc_camera_ray(f, (double)(u + jx), (double)(v + jy), org, dir);
                        nrays++;
                        if (trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
                            orc_state_f64 st;
                            state_build_uv(T, org, dir, t, uu, vv, prim, &st);
                            rad = dirt_radiance(T, &V, &st, &rng, &nrays);
                        }
                        accum = accum + rad;
                    }
                }
                px = accum * ((double)1.0 / (f->xsamples * f->ysamples));
                {
                    float *dst = rgb + 3 * ((size_t)(f->height - v - 1) * f->width + u);
                    dst[0] = dst[1] = dst[2] = (float)px;
                }
            }
        }
    }
    free(buckets);
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ Whitted transport */

/* reflection.c:25-49: the dot product is stored in a float */
static void reflect_f64(double out[3], const double in[3], const double n[3])
{
    float dot = in[0] * n[0] + in[1] * n[1] + in[2] * n[2];
    int k;
    for (k = 0; k < 3; k++) { const double nd = n[k] * (2 * dot); out[k] = in[k] - nd; }
}

/* reflection.c:69-127 */
static void refract_f64(double out[3], const double in[3], const double n[3], double eta)
{
    double cos1, coeff, N[3], e = 1.0 / eta;
    cos1 = in[0] * n[0] + in[1] * n[1] + in[2] * n[2];
    if (cos1 < 0.0) {
        cos1 = -cos1;
        N[0] = n[0]; N[1] = n[1]; N[2] = n[2];
    } else {
        e = eta;
        N[0] = -(n[0]); N[1] = -(n[1]); N[2] = -(n[2]);
    }
    coeff = 1.0 - (e * e) * (1.0 - cos1 * cos1);
    if (coeff <= 0.0) {                                       /* total internal reflection */
        reflect_f64(out, in, n);
        normalize_f64(out);
        return;
    }
    coeff = e * cos1 - sqrt(coeff);
    out[0] = coeff * N[0] + e * in[0];
    out[1] = coeff * N[1] + e * in[1];
    out[2] = coeff * N[2] + e * in[2];
    normalize_f64(out);
}

/* texture.c:238-277 */
static void ibl_fetch(const float *env, int w, int h, const double dir[3], double out[4])
{
    const double pi = 3.1415926535;
    double u, v, r, norm2, ndir[3] = { dir[0], dir[1], dir[2] };
    normalize_f64(ndir);
    if (ndir[2] >= -1.0 && ndir[2] < 1.0) r = (1.0 / pi) * acos(ndir[2]);
    else r = 0.0;
    norm2 = ndir[0] * ndir[0] + ndir[1] * ndir[1];
    if (norm2 > 1.0e-6) r /= sqrt(norm2);
    u = ndir[0] * r;
    v = ndir[1] * r;
    u = 0.5 * u + 0.5;
    v = 0.5 - 0.5 * v;
    tex_fetch(env, w, h, u, v, out);
}

/* whitted.c:92-151 with trace_whitted (:31-83) unrolled into a loop: the recursion only passes the newest hit down */
static void whitted_radiance(const orc_tree *T, const view_t_f64 *V, const float *env, int ew, int eh, const double eye_org[3],
                             const double eye_dir[3], double rad[3], uint64_t *nrays)
{
    const double eps = 1.0e-7, eta = 1.33;
    double org[3] = { eye_org[0], eye_org[1], eye_org[2] }, dir[3] = { eye_dir[0], eye_dir[1], eye_dir[2] };
    double t, uu, vv, texel[4];
    uint32_t prim;
    int depth, k;
    rad[0] = rad[1] = rad[2] = 0.0;
    (*nrays)++;
    if (!trace_f64(T, V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
        if (env) { ibl_fetch(env, ew, eh, dir, texel); for (k = 0; k < 3; k++) rad[k] = texel[k]; }
        return;
    }
    for (depth = 1; depth <= 8; depth++) {                    /* MAX_TRACE_DEPTH 8, whitted.c:24,48 */
        orc_state_f64 st;
        double I[3] = { dir[0], dir[1], dir[2] }, Rd[3];
        state_build_uv(T, org, dir, t, uu, vv, prim, &st);
        normalize_f64(I);                                     /* intersection_state.c:130-131 */
        refract_f64(Rd, I, st.Ns, eta);
        for (k = 0; k < 3; k++) { org[k] = st.P[k] + eps * Rd[k]; dir[k] = Rd[k]; }
        (*nrays)++;
        if (!trace_f64(T, V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
            if (env) { ibl_fetch(env, ew, eh, Rd, texel); for (k = 0; k < 3; k++) rad[k] = texel[k]; }
            return;
        }
    }
}

void orc_transport_whitted(const orc_tree *T, const float *env, int ew, int eh, const double *rays, uint64_t n, double *radiance3,
                           uint64_t *nrays_out)
{
    view_t_f64 V = {0};
    uint64_t i, nrays = 0;
    if (!T->empty) view64(T, &V);
    for (i = 0; i < n; i++) whitted_radiance(T, &V, env, ew, eh, rays + 6 * i, rays + 6 * i + 3, radiance3 + 3 * i, &nrays);
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ socket display driver stream
 * display/sockdrv.c:118-262 + sockdrv_defs.h: [COMMAND_NEW=0][8][width][height], then one [COMMAND_PIXEL=2][24*1024] message per
 * MAXPACKETS = 1024 pixels {int x, int y, float r, g, b, 1.0f} in the order bucket_write hands them over (render.c:919-979: buckets in
 * spiral order, rows inside a bucket, y = H-1-row), then [COMMAND_FINISH=1].  Pixels left over when the frame is not a multiple of
 * 1024 are never sent (sock_dd_close does not flush gpackets) -- reproduced.  rgb = [height][width][3] floats in display order. */
uint64_t orc_sockdrv_encode(const float *rgb, int width, int height, int bucket_size, unsigned char *out, uint64_t cap)
{
    int nb_max = (width / bucket_size + 1) * (height / bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(width, height, bucket_size, buckets, nb_max), b, sx, sy;
    const uint64_t npix = (uint64_t)width * height, ngroups = npix / 1024;
    const uint64_t need = 16 + ngroups * (8 + 24 * 1024) + 4;
    uint64_t count = 0;
    unsigned char *w = out;
    if (!out || cap < need) { free(buckets); return need; }
    { int32_t h[4] = { 0, 8, width, height }; memcpy(w, h, 16); w += 16; }
    for (b = 0; b < nb; b++) {
        const int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        for (sy = 0; sy < bh; sy++) {
            for (sx = 0; sx < bw; sx++) {
                const int x = sx + bx, y = height - (sy + by) - 1;
                const float *px = rgb + 3 * ((size_t)y * width + x);
                int32_t xy[2] = { x, y };
                float col[4] = { px[0], px[1], px[2], 1.0f };
                if (count / 1024 >= ngroups) continue;                         /* the unsent remainder */
                if (count % 1024 == 0) { int32_t h[2] = { 2, 24 * 1024 }; memcpy(w, h, 8); w += 8; }
                memcpy(w, xy, 8); memcpy(w + 8, col, 16); w += 24;
                count++;
            }
        }
    }
    { int32_t fin = 1; memcpy(w, &fin, 4); w += 4; }
    free(buckets);
    return (uint64_t)(w - out);
}

/* ------------------------------------------------------------------ hemisphere gathers at shading points
 * The per-point loops of three more ri_raytrace callers, Monte Carlo branches (Option use_qmc defaults to 0, option.c:139):
 *   kind 0  occlusion() shadeop            shader.c:680-768   coverage / nsamples (float)
 *   kind 1  ri_ibl_sample_cosweight        ibl.c:53-228       pi * sum(Le / pi) / (ntheta * nphi), Le = angular-map lookup on a miss
 *   kind 2  ri_domelight_sample            ibl.c:231-389      pi * sum(col * intensity / pi) / nsamples on misses
 * points = [n][6] (P, N); one MT19937 stream (randomMT / randomMT2, same generator and seed, random.c:163-247) over the points in
 * order, 2 draws per ray whether it hits or not. */
/* qmc.c:182-260 faure_permutation, one base: p_2 = (0,1); odd b: p_{b-1} with values >= (b-1)/2 raised by one and (b-1)/2 put in
 * the middle; even b: (2 p_{b/2}, 2 p_{b/2} + 1). */
static void faure_perm(int base, int *out)
{
    int tmp[128], j;
    if (base == 2) { out[0] = 0; out[1] = 1; return; }
    if (base % 2 != 0) {
        const int c = (base - 1) / 2;
        faure_perm(base - 1, tmp);
        for (j = 0; j < c; j++) out[j] = (2 * tmp[j] >= base - 1) ? tmp[j] + 1 : tmp[j];
        out[c] = c;
        for (j = c + 1; j < base; j++) out[j] = (2 * tmp[j - 1] >= base - 1) ? tmp[j - 1] + 1 : tmp[j - 1];
    } else {
        faure_perm(base / 2, tmp);
        for (j = 0; j < base / 2; j++) out[j] = 2 * tmp[j];
        for (j = base / 2; j < base; j++) out[j] = out[j - base / 2] + 1;
    }
}

static const int qmc_primes[25] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97 };

/* qmc.c:329-349 generalized_vdC */
static double generalized_vdc(int i, int base, const int *perm)
{
    double h = 0.0, f, factor;
    f = factor = 1.0 / (double)base;
    while (i > 0) {
        h += (double)perm[i % base] * factor;
        i /= base;
        factor *= f;
    }
    return h;
}

static double mod_1(double x) { return x - floor(x); }         /* qmc.c:523-532 */

/* The quasi-Monte Carlo branches (Option "use_qmc"): ri_ibl_sample_cosweight (ibl.c:107-151, gen_qmc_samples :540-581: scrambled
 * Halton in bases 3 and 5 at index i + inray->i; brdf is the float constant 0.31831f and there is no pi in the result) and
 * ri_domelight_sample (ibl.c:266-320: Hammersley (i/n, base 3) shifted by ONE scrambled-Halton value u of (inray->i, inray->d) --
 * both coordinates by u, the v it also computes is never used).  nsamples rays per point, no random numbers.  instance = inray->i
 * per point (NULL: 0), dim = inray->d (< 1 counts as 1; primes[dim] must stay below 100, the size of lucille's permutation table). */
void orc_point_gather_qmc(const orc_tree *T, int kind, int nsamples, const double *points, uint64_t n, const int32_t *instance, int dim,
                          const float *env, int ew, int eh, const double *col3, double intensity, double *out3, uint64_t *nrays_out)
{
    view_t_f64 V = {0};
    uint64_t p, nrays = 0;
    int perm3[3], perm5[5], permd[128], primd, i, k;
    if (!T->empty) view64(T, &V);
    if (dim < 1) dim = 1;
    if (dim > 24) dim = 24;
    primd = qmc_primes[dim];
    faure_perm(3, perm3); faure_perm(5, perm5); faure_perm(primd, permd);
    for (p = 0; p < n; p++) {
        const double *P = points + 6 * p, *N = P + 3;
        const int inst = instance ? instance[p] : 0;
        double basis[3][3], dpower[3] = { 0.0, 0.0, 0.0 }, org[3], dirl[3], dir[3], theta, phi, t, uu, vv, s0, s1, u = 0.0;
        uint32_t prim;
        ortho_basis_f64(basis, N);
        if (kind == 2) u = generalized_vdc(inst, primd, permd);                 /* ibl.c:286-289 */
        for (i = 0; i < nsamples; i++) {
            if (kind == 1) {
                s0 = mod_1(0.0 + generalized_vdc(i + inst, 3, perm3));          /* ibl.c:560-575: dims 1 and 2 */
                s1 = mod_1(0.0 + generalized_vdc(i + inst, 5, perm5));
            } else {
                const int j = (i > nsamples) ? i % nsamples : i;                /* qmc.c:439-442 */
                s0 = mod_1(u + (double)i / (double)nsamples);                   /* hammersley dim 1 */
                s1 = mod_1(u + generalized_vdc(j, 3, perm3));                   /* hammersley dim 2: primes[1] */
            }
            theta = sqrt(s0);
            phi = 2.0 * M_PI * s1;
            dirl[0] = cos(phi) * theta;
            dirl[1] = sin(phi) * theta;
            dirl[2] = sqrt(1.0 - theta * theta);
            for (k = 0; k < 3; k++)
                dir[k] = dirl[0] * basis[0][k] + dirl[1] * basis[1][k] + dirl[2] * basis[2][k];
            normalize_f64(dir);
            for (k = 0; k < 3; k++) org[k] = P[k];
            if (kind == 1) for (k = 0; k < 3; k++) org[k] += N[k] * 0.0001;
            nrays++;
            if (!(T->empty ? 0 : trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL))) {
                double rad[4], brdf = (kind == 1) ? (double)0.31831f : (double)1.0 / M_PI;
                if (kind == 1) ibl_fetch(env, ew, eh, dir, rad);
                else for (k = 0; k < 3; k++) rad[k] = col3[k] * (double)intensity;
                for (k = 0; k < 3; k++) dpower[k] += rad[k] * brdf;
            }
        }
        for (k = 0; k < 3; k++)
            out3[3 * p + k] = (kind == 1) ? dpower[k] / (double)nsamples : M_PI * dpower[k] / (double)nsamples;
    }
    if (nrays_out) *nrays_out = nrays;
}

void orc_point_gather(const orc_tree *T, int kind, int nsamples, uint32_t seed, const double *points, uint64_t n,
                      const float *env, int ew, int eh, const double *col3, double intensity, double *out3, uint64_t *nrays_out)
{
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t p, nrays = 0;
    int ntheta, nphi, i, j, k;
    if (!T->empty) view64(T, &V);
    mt_seed(&rng, seed);
    if (kind == 0) ntheta = (int)((float)nsamples / 3.0);      /* shader.c:710 (nsamples is a float there) */
    else ntheta = (int)(nsamples / 3.0);                       /* ibl.c:160, 326 */
    ntheta = (int)sqrt((double)ntheta);
    if (ntheta < 1) ntheta = 1;
    if (kind != 0 && ntheta > 128) ntheta = 128;               /* MAX_HEMISAMPLE, ibl.h:20 */
    nphi = 3 * ntheta;
    for (p = 0; p < n; p++) {
        const double *P = points + 6 * p, *N = P + 3;
        double basis[3][3], dpower[3] = { 0.0, 0.0, 0.0 }, org[3], dirl[3], dir[3], theta, phi, t, uu, vv;
        uint32_t prim;
        int coverage = 0, hit;
        ortho_basis_f64(basis, N);
        for (j = 0; j < nphi; j++) {
            for (i = 0; i < ntheta; i++) {
                if (kind == 0) theta = sqrt((double)i + mt_next(&rng)) / (double)ntheta;           /* shader.c:731 */
                else theta = sqrt(((double)i + mt_next(&rng)) / ntheta);                           /* ibl.c:174, 337 */
                phi = 2.0 * M_PI * ((double)j + mt_next(&rng)) / (double)nphi;
                dirl[0] = cos(phi) * theta;
                dirl[1] = sin(phi) * theta;
                dirl[2] = sqrt(1.0 - theta * theta);
                for (k = 0; k < 3; k++)
                    dir[k] = dirl[0] * basis[0][k] + dirl[1] * basis[1][k] + dirl[2] * basis[2][k];
                normalize_f64(dir);
                for (k = 0; k < 3; k++) org[k] = P[k];
                if (kind == 0) for (k = 0; k < 3; k++) org[k] += 0.0001 * dir[k];                 /* shader.c:750-752 */
                if (kind == 1) for (k = 0; k < 3; k++) org[k] += N[k] * 0.0001;                   /* ibl.c:92-94 */
                nrays++;
                hit = T->empty ? 0 : trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL);
                if (kind == 0) {
                    if (hit) coverage++;
                } else if (!hit) {
                    double rad[4], brdf = 1.0 / M_PI;
                    if (kind == 1) ibl_fetch(env, ew, eh, dir, rad);
                    else for (k = 0; k < 3; k++) rad[k] = col3[k] * intensity;
                    for (k = 0; k < 3; k++) dpower[k] += rad[k] * brdf;
                }
            }
        }
        if (kind == 0) {
            if (coverage > nsamples) coverage = nsamples;
            out3[3 * p] = out3[3 * p + 1] = out3[3 * p + 2] = (double)(float)((double)coverage / (double)(float)nsamples);
        } else if (kind == 1) {
            for (k = 0; k < 3; k++) out3[3 * p + k] = M_PI * dpower[k] / (double)(ntheta * nphi);
        } else {
            for (k = 0; k < 3; k++) out3[3 * p + k] = M_PI * dpower[k] / (double)nsamples;
        }
    }
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ shading-language callers of ri_raytrace
 *
 * trace() shadeop (render/shader.c:895-976) up to the call of the hit geometry's shader procedure, for n (P, R) pairs:
 * ray.org = P + 0.0001 * R component by component, ray.dir = R (not normalised); on a miss dst = ri_texture_ibl_fetch(light->texture,
 * ray.dir) when the scene's first light is an IBL / sun-sky light (use_env), else zero; on a hit the shader's input block is the hit
 * state: Cs = state.color, P, N = Ns, Ng, dPdu = tangent, dPdv = binormal, I = normalize(state.P - P), s = u, t = v (shader.c:951-968).
 * rays_out [n][6] receives the rays that were traced (the offset origins). */
void orc_shade_trace(const orc_tree *T, const double *pr, uint64_t n, const float *env, int ew, int eh, int use_env,
                     double *rays_out, orc_hit_f64 *hits, orc_state_f64 *states, orc_state_ext_f64 *exts, double *eye3, double *miss_rgb3)
{
    uint64_t i;
    int k;
    for (i = 0; i < n; i++) {
        const double *P = pr + 6 * i, *R = P + 3;
        double *r = rays_out + 6 * i;
        for (k = 0; k < 3; k++) { r[3 + k] = R[k]; r[k] = P[k]; r[k] += 0.0001 * r[3 + k]; }
    }
    orc_intersect_f64(T, rays_out, n, hits, NULL);
    orc_state_build_f64(T, rays_out, hits, n, states);
    orc_state_ext_build_f64(T, rays_out, hits, n, exts);
    for (i = 0; i < n; i++) {
        const double *P = pr + 6 * i;
        double *e = eye3 + 3 * i, *m = miss_rgb3 + 3 * i;
        e[0] = e[1] = e[2] = 0.0; m[0] = m[1] = m[2] = 0.0;
        if (hits[i].hit) {
            for (k = 0; k < 3; k++) e[k] = states[i].P[k] - P[k];
            normalize_f64(e);
        } else if (use_env && env) {
            double texel[4];
            ibl_fetch(env, ew, eh, rays_out + 6 * i + 3, texel);
            for (k = 0; k < 3; k++) m[k] = texel[k];
        }
    }
}

/* next_lightsource() + init_lightsource() (render/shader.c:1116-1186, 1236-1310): the light samples an illuminance loop visits at
 * each of n shading points (P, N).  Per point: m = ntheta * 3 ntheta stratified cosine directions about N from the randomMT stream
 * (two words per sample, all drawn up front), L = normalize(direction), Cl = ri_texture_ibl_fetch(env, L) * (1 / m); the loop then
 * returns, in order, every sample with dot(L, N) > 0, acos(dot) < angle and no occluder along normalize(L) from P + 0.0001 * N --
 * except the LAST sample of the set, which the reference can never return (after the break `sample_index >= nsamples` is true and
 * the function returns NULL, shader.c:1170-1177).  visible[p][j] = 1 for the samples the loop returns.  Returns m. */
int orc_light_samples(const orc_tree *T, int nsamples, double angle, uint32_t seed, const double *points, uint64_t n,
                      const float *env, int ew, int eh, double *L_out, double *Cl_out, uint8_t *visible, uint64_t *nrays_out)
{
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t p, nrays = 0;
    int ntheta, nphi, m, i, j, k;
    if (!T->empty) view64(T, &V);
    mt_seed(&rng, seed);
    ntheta = (int)(nsamples / 3.0);                            /* shader.c:1263-1266 */
    ntheta = (int)sqrt((double)ntheta);
    if (ntheta < 1) ntheta = 1;
    nphi = 3 * ntheta;
    m = ntheta * nphi;
    for (p = 0; p < n; p++) {
        const double *P = points + 6 * p, *N = P + 3;
        double basis[3][3];
        double *L = L_out + 3 * (size_t)m * p, *Cl = Cl_out + 3 * (size_t)m * p;
        uint8_t *vis = visible + (size_t)m * p;
        int count = 0;
        ortho_basis_f64(basis, N);
        for (j = 0; j < nphi; j++) {
            for (i = 0; i < ntheta; i++) {
                double dirl[3], ldir[3], texel[4] = {0.0, 0.0, 0.0, 0.0};
                const double theta = sqrt(((double)i + mt_next(&rng)) / (double)ntheta);
                const double phi = 2.0 * M_PI * ((double)j + mt_next(&rng)) / (double)nphi;
                dirl[0] = cos(phi) * theta;
                dirl[1] = sin(phi) * theta;
                dirl[2] = sqrt(1.0 - theta * theta);
                for (k = 0; k < 3; k++) ldir[k] = dirl[0] * basis[0][k] + dirl[1] * basis[1][k] + dirl[2] * basis[2][k];
                normalize_f64(ldir);
                if (env) ibl_fetch(env, ew, eh, ldir, texel);
                for (k = 0; k < 3; k++) { L[3 * count + k] = ldir[k]; Cl[3 * count + k] = texel[k] * (double)(1.0f / (double)m); }
                count++;
            }
        }
        for (j = 0; j < m; j++) {
            const double *l = L + 3 * j;
            const double ndotl = l[0] * N[0] + l[1] * N[1] + l[2] * N[2];
            double org[3], dir[3], t, uu, vv;
            uint32_t prim;
            vis[j] = 0;
            if (ndotl <= 0.0) continue;
            if (acos(ndotl) >= angle) continue;
            for (k = 0; k < 3; k++) { org[k] = P[k]; org[k] += N[k] * 0.0001; dir[k] = l[k]; }
            normalize_f64(dir);
            nrays++;
            if (!T->empty && trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) continue;
            if (j + 1 < m) vis[j] = 1;                         /* the last sample is dropped: shader.c:1170-1177 */
        }
    }
    if (nrays_out) *nrays_out = nrays;
    return m;
}

void orc_render_whitted(const orc_tree *T, const orc_frame_t *f, const float *env, int ew, int eh, float *rgb, uint64_t *nrays_out)
{
    int nb_max = (f->width / f->bucket_size + 1) * (f->height / f->bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(f->width, f->height, f->bucket_size, buckets, nb_max);
    view_t_f64 V = {0};
    uint64_t nrays = 0;
    int b, k;
    if (!T->empty) view64(T, &V);
    for (b = 0; b < nb; b++) {
        int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        int u, v;
        for (v = by; v < by + bh; v++) {
            for (u = bx; u < bx + bw; u++) {
                double accum[3] = {0.0, 0.0, 0.0};
                int xs, ys;
                float *dst = rgb + 3 * ((size_t)(f->height - v - 1) * f->width + u);
                for (ys = 0; ys < f->ysamples; ys++) {
                    for (xs = 0; xs < f->xsamples; xs++) {
                        double jx, jy, org[3], dir[3], rad[3];
                        orc_subpixel_jitter(xs, ys, f->xsamples, f->ysamples, &jx, &jy);
                        orc_camera_ray(f, (double)(u + jx), (double)(v + jy), org, dir);
                        whitted_radiance(T, &V, env, ew, eh, org, dir, rad, &nrays);
                        for (k = 0; k < 3; k++) accum[k] = accum[k] + rad[k];
                    }
                }
                for (k = 0; k < 3; k++) dst[k] = (float)(accum[k] * ((double)1.0 / (f->xsamples * f->ysamples)));
            }
        }
    }
    free(buckets);
    if (nrays_out) *nrays_out = nrays;
}

/* transport.c:50-173 ri_transport_sample -> trace_path: radiance = white on a hit (light geometry aside), black on a miss */
void orc_render_hitmask(const orc_tree *T, const orc_frame_t *f, float *rgb, uint64_t *nrays_out)
{
    int nb_max = (f->width / f->bucket_size + 1) * (f->height / f->bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(f->width, f->height, f->bucket_size, buckets, nb_max);
    view_t_f64 V = {0};
    uint64_t nrays = 0;
    int b;
    if (!T->empty) view64(T, &V);
    for (b = 0; b < nb; b++) {
        int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        int u, v;
        for (v = by; v < by + bh; v++) {
            for (u = bx; u < bx + bw; u++) {
                double accum = 0.0, px;
                int xs, ys;
                float *dst = rgb + 3 * ((size_t)(f->height - v - 1) * f->width + u);
                for (ys = 0; ys < f->ysamples; ys++) {
                    for (xs = 0; xs < f->xsamples; xs++) {
                        double jx, jy, org[3], dir[3], t, uu, vv;
                        uint32_t prim;
                        orc_subpixel_jitter(xs, ys, f->xsamples, f->ysamples, &jx, &jy);
                        orc_camera_ray(f, (double)(u + jx), (double)(v + jy), org, dir);
                        nrays++;
                        accum = accum + (trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL) ? 1.0 : 0.0);
                    }
                }
                px = accum * ((double)1.0 / (f->xsamples * f->ysamples));
                dst[0] = dst[1] = dst[2] = (float)px;
            }
        }
    }
    free(buckets);
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ sun-sky gather (row a12) */

/* sunsky.c:24-38.  All variables are float, the libm calls are the double ones: every `sin(x)` promotes its float argument
 * and every assignment rounds the double expression once. */
static float sky_angle_between(float v_theta, float v_phi, float s_theta, float s_phi)
{
    float cospi = sin(v_theta) * sin(s_theta) * cos(s_phi - v_phi) + cos(v_theta) * cos(s_theta);
    if (cospi > 1.0) return 0.0;
    if (cospi < -1.0) return M_PI;
    return acos(cospi);
}

/* sunsky.c:136-152 */
static float sky_perez(const float *lam, float theta, float gamma, float lvz, float sun_theta)
{
    float den, num;
    den = ((1.0 + lam[0] * exp(lam[1])) *
           (1.0 + lam[2] * exp(lam[3] * sun_theta) + lam[4] * cos(sun_theta) * cos(sun_theta)));
    num = ((1.0 + lam[0] * exp(lam[1] / cos(theta))) *
           (1.0 + lam[2] * exp(lam[3] * gamma) + lam[4] * cos(gamma) * cos(gamma)));
    return lvz * num / den;
}

/* specrend.c:366-440 (the float* overload): 81 five-nanometre bins, each 10 nm sample used twice, float accumulation */
static void sky_spectrum_to_xyz(const orc_sunsky_t *s, const float *spectrum, float *x, float *y, float *z)
{
    int i;
    float lambda, X = 0, Y = 0, Z = 0;
    for (i = 0, lambda = 380; lambda < 780.1; i++, lambda += 5) {
        X += spectrum[i / 2] * s->cie[i][0];
        Y += spectrum[i / 2] * s->cie[i][1];
        Z += spectrum[i / 2] * s->cie[i][2];
    }
    *x = X; *y = Y; *z = Z;
}

/* specrend.c:127-172 */
static void sky_xyz_to_rgb(const float *cs, float xc, float yc, float zc, float *r, float *g, float *b)
{
    float xr, yr, zr, xg, yg, zg, xb, yb, zb, xw, yw, zw;
    float rx, ry, rz, gx, gy, gz, bx, by, bz, rw, gw, bw;
    xr = cs[0]; yr = cs[1]; zr = 1 - (xr + yr);
    xg = cs[2]; yg = cs[3]; zg = 1 - (xg + yg);
    xb = cs[4]; yb = cs[5]; zb = 1 - (xb + yb);
    xw = cs[6]; yw = cs[7]; zw = 1 - (xw + yw);
    rx = (yg * zb) - (yb * zg); ry = (xb * zg) - (xg * zb); rz = (xg * yb) - (xb * yg);
    gx = (yb * zr) - (yr * zb); gy = (xr * zb) - (xb * zr); gz = (xb * yr) - (xr * yb);
    bx = (yr * zg) - (yg * zr); by = (xg * zr) - (xr * zg); bz = (xr * yg) - (xg * yr);
    if (fabs(yw) > 1.0e-48) {
        rw = ((rx * xw) + (ry * yw) + (rz * zw)) / yw;
        gw = ((gx * xw) + (gy * yw) + (gz * zw)) / yw;
        bw = ((bx * xw) + (by * yw) + (bz * zw)) / yw;
    } else {
        rw = 1.0; gw = 1.0; bw = 1.0;
    }
    rx = rx / rw; ry = ry / rw; rz = rz / rw;
    gx = gx / gw; gy = gy / gw; gz = gz / gw;
    bx = bx / bw; by = by / bw; bz = bz / bw;
    *r = (rx * xc) + (ry * yc) + (rz * zc);
    *g = (gx * xc) + (gy * yc) + (gz * zc);
    *b = (bx * xc) + (by * yc) + (bz * zc);
}

/* sunsky.c:322-408: ri_sunsky_get_sky_spectrum + ri_sunsky_get_sky_rgb (y and z swapped on entry, :337-339) */
static void sky_rgb(const orc_sunsky_t *s, const float v[3], float rgb[3])
{
    float spec[41];
    float theta, phi, vlen, gamma, x, y, Y, lx, ly, lz, t[3], M1, M2, X, Yc, Z;
    int i;
    t[0] = v[0]; t[1] = v[2]; t[2] = v[1];
    if (t[2] < 0.0) {                                         /* under the horizon: zero spectrum */
        for (i = 0; i < 41; i++) spec[i] = 0.0f;
    } else {
        if (t[2] < 0.001) {
            t[2] = 0.001;
            vlen = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            t[0] /= vlen; t[1] /= vlen; t[2] /= vlen;
        }
        theta = acos(t[2]);
        if (fabs(theta) < 1.0e-6) phi = 0.0; else phi = atan2(t[1], t[0]);
        gamma = sky_angle_between(theta, phi, s->sun_theta, s->sun_phi);
        x = sky_perez(s->perez_x, theta, gamma, s->zenith_x, s->sun_theta);
        y = sky_perez(s->perez_y, theta, gamma, s->zenith_y, s->sun_theta);
        Y = sky_perez(s->perez_Y, theta, gamma, s->zenith_Y, s->sun_theta);
        /* chromaticity_to_spectrum, sunsky.c:297-314 */
        M1 = (-1.3515 - 1.7703 * x + 5.9114 * y) / (0.0241 + 0.2562 * x - 0.7341 * y);
        M2 = (0.03 - 31.4424 * x + 30.0717 * y) / (0.0241 + 0.2562 * x - 0.7341 * y);
        for (i = 0; i < 41; i++) spec[i] = s->S0[i] + M1 * s->S1[i] + M2 * s->S2[i];
        sky_spectrum_to_xyz(s, spec, &lx, &ly, &lz);
        if (fabs(ly) < 1.0e-48) ly = 1.0;
        for (i = 0; i < 41; i++) spec[i] = Y * spec[i] / ly;
    }
    sky_spectrum_to_xyz(s, spec, &X, &Yc, &Z);
    sky_xyz_to_rgb(s->cs, X, Yc, Z, &rgb[0], &rgb[1], &rgb[2]);
}

void orc_sunsky_sky_rgb(const orc_sunsky_t *s, const float *dirs, uint64_t n, float *rgb)
{
    uint64_t i;
    for (i = 0; i < n; i++) sky_rgb(s, dirs + 3 * i, rgb + 3 * i);
}

/* ambientocclusion.c:206-324 gather_sunsky(8, 8) + 153-199 contribution_from_sunlight */
static void sunsky_radiance(const orc_tree *T, const view_t_f64 *V, const orc_state_f64 *st, const orc_sunsky_t *s,
                            mt_t *rng, uint64_t *nrays, double Lo[3])
{
    const uint32_t ntheta = 8, nphi = 8;
    const double eps = 1.0e-5;
    double basis[3][3], org[3], dirl[3], dir[3], col[3] = {0.0, 0.0, 0.0};
    double t, u, v, nsamples, m;
    uint32_t prim, i, j;
    int k, l;

    ortho_basis_f64(basis, st->Ns);
    for (k = 0; k < 3; k++) org[k] = st->P[k];
    for (k = 0; k < 3; k++) org[k] += st->Ns[k] * eps;
    for (j = 0; j < nphi; j++) {
        for (i = 0; i < ntheta; i++) {
            double z0 = (i + mt_next(rng)) / (double)ntheta;
            double z1 = (j + mt_next(rng)) / (double)nphi;
            double cos_theta = sqrt(z0);
            double phi = 2.0 * M_PI * z1;
            dirl[0] = cos(phi) * cos_theta;
            dirl[1] = sin(phi) * cos_theta;
            dirl[2] = sqrt(1.0 - cos_theta * cos_theta);
            for (k = 0; k < 3; k++)
                dir[k] = dirl[0] * basis[0][k] + dirl[1] * basis[1][k] + dirl[2] * basis[2][k];
            (*nrays)++;
            if (!trace_f64(T, V, org, dir, 0, &t, &u, &v, &prim, NULL)) {
                float vf[3], c[3];
                vf[0] = dir[0]; vf[1] = dir[1]; vf[2] = dir[2];
                sky_rgb(s, vf, c);
                col[0] += c[0]; col[1] += c[1]; col[2] += c[2];
            }
        }
    }
    for (l = 0; l < s->nsun; l++) {                           /* one shadow ray per LIGHTTYPE_SUNLIGHT, same offset origin */
        for (k = 0; k < 3; k++) dir[k] = s->sun_dir[l][k];
        (*nrays)++;
        if (!trace_f64(T, V, org, dir, 0, &t, &u, &v, &prim, NULL))
            for (k = 0; k < 3; k++) col[k] += s->sun_col[l][k];
    }
    nsamples = ntheta * nphi;
    m = (1.0 / M_PI);
    for (k = 0; k < 3; k++) Lo[k] = m * col[k] / nsamples;
}

/* the pixel loop of orc_render_ao with the sun-sky transport (ambientocclusion.c:369-376) and three channels (render.c:805,820) */
void orc_render_sunsky(const orc_tree *T, const orc_frame_t *f, const orc_sunsky_t *s, float *rgb, uint64_t *nrays_out)
{
    int nb_max = (f->width / f->bucket_size + 1) * (f->height / f->bucket_size + 1);
    int32_t *buckets = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nb_max);
    int nb = orc_bucket_list(f->width, f->height, f->bucket_size, buckets, nb_max);
    view_t_f64 V = {0};
    mt_t rng;
    uint64_t nrays = 0;
    int b, k;

    if (!T->empty) view64(T, &V);
    mt_seed(&rng, 4357);
    for (b = 0; b < nb; b++) {
        int bx = buckets[4 * b], by = buckets[4 * b + 1], bw = buckets[4 * b + 2], bh = buckets[4 * b + 3];
        int u, v;
        for (v = by; v < by + bh; v++) {
            for (u = bx; u < bx + bw; u++) {
                double accum[3] = {0.0, 0.0, 0.0};
                int xs, ys;
                float *dst = rgb + 3 * ((size_t)(f->height - v - 1) * f->width + u);
                for (ys = 0; ys < f->ysamples; ys++) {
                    for (xs = 0; xs < f->xsamples; xs++) {
                        double jx, jy, org[3], dir[3], t, uu, vv, rad[3] = {0.0, 0.0, 0.0};
                        uint32_t prim;
                        orc_subpixel_jitter(xs, ys, f->xsamples, f->ysamples, &jx, &jy);
                        orc_camera_ray(f, (double)(u + jx), (double)(v + jy), org, dir);
                        nrays++;
                        if (trace_f64(T, &V, org, dir, 0, &t, &uu, &vv, &prim, NULL)) {
                            orc_state_f64 st;
                            state_build_uv(T, org, dir, t, uu, vv, prim, &st);
                            sunsky_radiance(T, &V, &st, s, &rng, &nrays, rad);
                        }
                        for (k = 0; k < 3; k++) accum[k] = accum[k] + rad[k];
                    }
                }
                for (k = 0; k < 3; k++) dst[k] = (float)(accum[k] * ((double)1.0 / (f->xsamples * f->ysamples)));
            }
        }
    }
    free(buckets);
    if (nrays_out) *nrays_out = nrays;
}

/* ------------------------------------------------------------------ .hdr output */

typedef struct { uint8_t *out; uint64_t cap, n; } sink_t;
static void sink_put(sink_t *s, const void *p, uint64_t len)
{
    if (s->out && s->n + len <= s->cap) memcpy(s->out + s->n, p, (size_t)len);
    s->n += len;
}

/* rgbe.c:78-96 */
static void hdr_float2rgbe(unsigned char rgbe[4], float red, float green, float blue)
{
    float v;
    int e;
    v = red;
    if (green > v) v = green;
    if (blue > v) v = blue;
    if (v < 1e-32) {
        rgbe[0] = rgbe[1] = rgbe[2] = rgbe[3] = 0;
    } else {
        v = frexp(v, &e) * 256.0 / v;
        rgbe[0] = (unsigned char)(red * v);
        rgbe[1] = (unsigned char)(green * v);
        rgbe[2] = (unsigned char)(blue * v);
        rgbe[3] = (unsigned char)(e + 128);
    }
}

/* rgbe.c:244-294 */
static void hdr_rle(sink_t *s, const unsigned char *data, int numbytes)
{
    const int MINRUNLENGTH = 4;
    int cur, beg_run, run_count, old_run_count, nonrun_count;
    unsigned char buf[2];
    cur = 0;
    while (cur < numbytes) {
        beg_run = cur;
        run_count = old_run_count = 0;
        while ((run_count < MINRUNLENGTH) && (beg_run < numbytes)) {
            beg_run += run_count;
            old_run_count = run_count;
            run_count = 1;
            while ((beg_run + run_count < numbytes) && (run_count < 127) && (data[beg_run] == data[beg_run + run_count]))
                run_count++;
        }
        if ((old_run_count > 1) && (old_run_count == beg_run - cur)) {
            buf[0] = 128 + old_run_count;
            buf[1] = data[cur];
            sink_put(s, buf, 2);
            cur = beg_run;
        }
        while (cur < beg_run) {
            nonrun_count = beg_run - cur;
            if (nonrun_count > 128) nonrun_count = 128;
            buf[0] = nonrun_count;
            sink_put(s, buf, 1);
            sink_put(s, &data[cur], (uint64_t)nonrun_count);
            cur += nonrun_count;
        }
        if (run_count >= MINRUNLENGTH) {
            buf[0] = 128 + run_count;
            buf[1] = data[beg_run];
            sink_put(s, buf, 2);
            cur += run_count;
        }
    }
}

uint64_t orc_hdr_encode(const float *rgb, int width, int height, uint8_t *out, uint64_t cap)
{
    sink_t s = { out, cap, 0 };
    char hdr[128];
    unsigned char rgbe[4], *buffer;
    int x, y, i, len;
    /* RGBE_WriteHeader(fp, width, height, NULL), rgbe.c:117-139 */
    len = snprintf(hdr, sizeof(hdr), "#?%s\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", "RGBE", height, width);
    sink_put(&s, hdr, (uint64_t)len);
    buffer = (unsigned char *)malloc((size_t)4 * (size_t)(width > 0 ? width : 1));
    for (y = 0; y < height; y++) {
        const float *row = rgb + 3 * (size_t)y * (size_t)width;
        const int flat = (width < 8) || (width > 0x7fff);                      /* rgbe.c:303-305 */
        if (!flat) {
            rgbe[0] = 2; rgbe[1] = 2; rgbe[2] = width >> 8; rgbe[3] = width & 0xFF;
            sink_put(&s, rgbe, 4);
        }
        for (x = 0; x < width; x++) {
            float col[3], acc[3];
            for (i = 0; i < 3; i++) {                                          /* hdr_dd_write, hdrdrv.c:72-85 */
                col[i] = row[3 * x + i];
                if (col[i] < 0.0) col[i] = 0.0;
                acc[i] = 0.0f;
                acc[i] += col[i];
            }
            hdr_float2rgbe(rgbe, acc[0], acc[1], acc[2]);
            if (flat) sink_put(&s, rgbe, 4);                                   /* RGBE_WritePixels, rgbe.c:206-222 */
            else { buffer[x] = rgbe[0]; buffer[x + width] = rgbe[1]; buffer[x + 2 * width] = rgbe[2]; buffer[x + 3 * width] = rgbe[3]; }
        }
        if (!flat)
            for (i = 0; i < 4; i++) hdr_rle(&s, &buffer[i * width], width);
    }
    free(buffer);
    return s.n;
}

/* ------------------------------------------------------------------ beam visibility (row a10) */

typedef struct {
    double org[3], dir[4][3], normal[4][3], t_max;
    int    dominant_axis, dirsign[3];
} beam_t;

/* beam.c:332-466 ri_beam_set (the parts the visibility query reads) */
static int beam_set(beam_t *b, const double org[3], const double dir[4][3])
{
    int i, j, k;
    double maxval;
    const double d = 1024.0;
    b->t_max = ORC_INFINITY;
    for (i = 0; i < 3; i++) {                                  /* all four directions in one octant? */
        int zeros = 0, mask = 0;
        for (j = 0; j < 4; j++) {
            if (fabs(dir[j][i]) < ORC_EPS) zeros++;
            else mask += (dir[j][i] < 0.0) ? 1 : -1;
        }
        if ((mask != -(4 - zeros)) && (mask != (4 - zeros))) return -1;
    }
    for (k = 0; k < 3; k++) b->org[k] = org[k];
    maxval = fabs(dir[0][0]);
    b->dominant_axis = 0;
    if (maxval < fabs(dir[0][1])) { maxval = fabs(dir[0][0]); b->dominant_axis = 1; }     /* sic: beam.c:389-392 keeps |x| */
    if (maxval < fabs(dir[0][2])) { maxval = fabs(dir[0][2]); b->dominant_axis = 2; }
    for (k = 0; k < 3; k++) b->dirsign[k] = (dir[0][k] < 0.0) ? 1 : 0;
    {
        double normal[3] = { 0.0, 0.0, 0.0 };
        normal[b->dominant_axis] = 1.0;
        if (b->dirsign[b->dominant_axis]) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
        for (i = 0; i < 4; i++) {
            const double t = dir[i][0] * normal[0] + dir[i][1] * normal[1] + dir[i][2] * normal[2];
            const double kk = (fabs(t) > ORC_EPS) ? d / t : 1.0;
            for (k = 0; k < 3; k++) b->dir[i][k] = kk * dir[i][k];
        }
    }
    cross_f64(b->normal[0], b->dir[1], b->dir[0]);
    cross_f64(b->normal[1], b->dir[2], b->dir[1]);
    cross_f64(b->normal[2], b->dir[3], b->dir[2]);
    cross_f64(b->normal[3], b->dir[0], b->dir[3]);
    return 0;
}

/* bvh.c:2012-2089 test_beam_aabb / test_beam_aabb_misses / get_n_point: 1 = may hit */
static int beam_aabb(const double *box /* min xyz, max xyz */, const beam_t *b)
{
    int i, k;
    for (i = 0; i < 4; i++) {
        const double *n = b->normal[i];
        double np[3], no[3], d;
        for (k = 0; k < 3; k++) np[k] = (n[k] > 0.0) ? box[k] : box[3 + k];
        for (k = 0; k < 3; k++) no[k] = np[k] - b->org[k];
        d = no[0] * n[0] + no[1] * n[1] + no[2] * n[2];
        if (d > 0.0) return 0;
    }
    return 1;
}

/* bvh.c:2139-2281 test_beam_triangle */
static int beam_triangle(const double *v /* v0 v1 v2 */, const beam_t *b)
{
    double e1[3], e2[3], u[4], vv[4], t[4];
    int i, k, mask = 0, cnt;
    for (k = 0; k < 3; k++) { e1[k] = v[3 + k] - v[k]; e2[k] = v[6 + k] - v[k]; }
    for (i = 0; i < 4; i++) {
        double p[3], q[3], s[3], a, inva;
        cross_f64(p, b->dir[i], e2);
        a = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
        inva = (fabs(a) > ORC_EPS) ? 1.0 / a : 0.0;
        for (k = 0; k < 3; k++) s[k] = b->org[k] - v[k];
        cross_f64(q, s, e1);
        u[i]  = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inva;
        vv[i] = (q[0] * b->dir[i][0] + q[1] * b->dir[i][1] + q[2] * b->dir[i][2]) * inva;
        t[i]  = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inva;
        if ((u[i] < 0.0) || (u[i] > 1.0)) continue;
        if ((vv[i] < 0.0) || ((u[i] + vv[i]) > 1.0)) continue;
        if ((t[i] < 0.0) || (t[i] > b->t_max)) continue;
        mask |= (1 << i);
    }
    if (mask == 0) {
        cnt = 0; for (i = 0; i < 4; i++) if (t[i] < 0.0) cnt++;
        if (cnt == 4) return ORC_BEAM_MISS_COMPLETELY;
        cnt = 0; for (i = 0; i < 4; i++) if (u[i] < 0.0) cnt++;
        if ((cnt != 0) && (cnt != 4)) return ORC_BEAM_HIT_PARTIALLY;
        cnt = 0; for (i = 0; i < 4; i++) if (u[i] > 1.0) cnt++;
        if ((cnt != 0) && (cnt != 4)) return ORC_BEAM_HIT_PARTIALLY;
        cnt = 0; for (i = 0; i < 4; i++) if (vv[i] < 0.0) cnt++;
        if ((cnt != 0) && (cnt != 4)) return ORC_BEAM_HIT_PARTIALLY;
        cnt = 0; for (i = 0; i < 4; i++) if ((u[i] + vv[i]) >= 1.0) cnt++;
        if ((cnt != 0) && (cnt != 4)) return ORC_BEAM_HIT_PARTIALLY;
        return ORC_BEAM_MISS_COMPLETELY;
    }
    return (mask == 0xf) ? ORC_BEAM_HIT_COMPLETELY : ORC_BEAM_HIT_PARTIALLY;
}

/* bvh.c:612-667 + 2648-2746 + 2435-2542 */
static int beam_query(const orc_tree *T, const beam_t *b)
{
    int64_t stack[ORC_STACK], node = 0, i;
    int depth = 0;
    double sb[6];
    int k;
    if (T->empty) return 0;
    for (k = 0; k < 3; k++) { sb[k] = T->bmin[k]; sb[3 + k] = T->bmax[k]; }
    if (!beam_aabb(sb, b)) return ORC_BEAM_MISS_COMPLETELY;
    for (;;) {
        const orc_node_t *n = &T->nodes[node];
        if (n->is_leaf) {
            for (i = 0; i < n->ntris; i++) {
                const int r = beam_triangle(T->tri_xyz + 9 * (n->tri_start + i), b);
                if (r == ORC_BEAM_HIT_COMPLETELY || r == ORC_BEAM_HIT_PARTIALLY) return r;
            }
            if (depth < 1) return ORC_BEAM_MISS_COMPLETELY;
            node = stack[--depth];
        } else {
            const int hl = beam_aabb(n->lbox, b), hr = beam_aabb(n->rbox, b);
            const int ret = hl | (hr << 1);
            if (ret == 0) {
                if (depth < 1) return ORC_BEAM_MISS_COMPLETELY;
                node = stack[--depth];
            } else if (ret == 1) node = n->child0;
            else if (ret == 2) node = n->child1;
            else {
                const int order = b->dirsign[b->dominant_axis];
                stack[depth++] = order ? n->child0 : n->child1;
                node = order ? n->child1 : n->child0;
            }
        }
    }
}

void orc_beam_visibility(const orc_tree *T, const double *beams, uint64_t n, int32_t *out)
{
    uint64_t i;
    for (i = 0; i < n; i++) {
        const double *p = beams + 15 * i;
        double dir[4][3];
        beam_t b;
        int j, k;
        for (j = 0; j < 4; j++) for (k = 0; k < 3; k++) dir[j][k] = p[3 + 3 * j + k];
        if (beam_set(&b, p, dir) != 0) { out[i] = ORC_BEAM_INVALID; continue; }
        out[i] = beam_query(T, &b);
    }
}

/* ------------------------------------------------------------------ path trace (row P) */

/* sin/cos of 2*pi*r, r in [0,1): quadrant reduction + Taylor polynomials evaluated with plain IEEE multiplies and adds in a
 * fixed order (no libm), so the CUDA kernel reproduces it bit for bit.  |error| < 1e-15. */
void orc_det_sincos2pi(double r, double *s_out, double *c_out)
{
    const double q = floor(4.0 * r + 0.5);
    const double y = r - 0.25 * q;                       /* |y| <= 1/8 */
    const double th = 6.283185307179586476925286766559 * y;
    const double z = th * th;
    double sp = -7.6471637318198164759e-13;              /* -1/15! */
    sp = sp * z + 1.6059043836821614599e-10;             /*  1/13! */
    sp = sp * z + -2.5052108385441718775e-08;            /* -1/11! */
    sp = sp * z + 2.7557319223985890653e-06;             /*  1/9!  */
    sp = sp * z + -1.9841269841269841270e-04;            /* -1/7!  */
    sp = sp * z + 8.3333333333333333333e-03;             /*  1/5!  */
    sp = sp * z + -1.6666666666666666667e-01;            /* -1/3!  */
    const double sn = th + th * (z * sp);
    double cp = 4.7794773323873852974e-14;               /*  1/16! */
    cp = cp * z + -1.1470745597729724714e-11;            /* -1/14! */
    cp = cp * z + 2.0876756987868098979e-09;             /*  1/12! */
    cp = cp * z + -2.7557319223985890653e-07;            /* -1/10! */
    cp = cp * z + 2.4801587301587301587e-05;             /*  1/8!  */
    cp = cp * z + -1.3888888888888888889e-03;            /* -1/6!  */
    cp = cp * z + 4.1666666666666666667e-02;             /*  1/4!  */
    cp = cp * z + -0.5;
    const double cs = 1.0 + z * cp;
    const int qi = ((int)q) & 3;
    if (qi == 0)      { *s_out = sn;  *c_out = cs; }
    else if (qi == 1) { *s_out = cs;  *c_out = -sn; }
    else if (qi == 2) { *s_out = -sn; *c_out = -cs; }
    else              { *s_out = -cs; *c_out = sn; }
}

static double path_uniform(uint32_t seed, uint64_t sample_id, uint32_t k)
{
    const uint64_t idx = sample_id * 64u + k;
    return (double)(orc_splitmix64((uint64_t)seed + idx * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
}

/* pathtrace.c:480-508 sample_cosweight */
static void path_cosweight(double out[3], const double n[3], double r0, double r1)
{
    double basis[3][3], sn, cs;
    const double cost = sqrt(r0), sint = sqrt(1.0 - r0);
    int i;
    ortho_basis_f64(basis, n);
    orc_det_sincos2pi(r1, &sn, &cs);
    {
        const double v0 = (double)(float)(cs * sint), v1 = (double)(float)(sn * sint), v2 = (double)(float)cost;
        for (i = 0; i < 3; i++) out[i] = v0 * basis[0][i] + v1 * basis[1][i] + v2 * basis[2][i];
    }
}

void orc_render_pathtrace(const orc_tree *T, const orc_path_frame_t *f, float *rgb, uint64_t *nrays_out)
{
    view_t_f64 V = {0};
    orc_frame_t cam;
    uint64_t nrays = 0;
    int x, y, s;
    const double inv_pi = 1.0 / 3.14159265358979323846;
    if (!T->empty) view64(T, &V);
    memset(&cam, 0, sizeof(cam));
    memcpy(cam.c2w, f->c2w, sizeof(cam.c2w));
    cam.flength = f->flength; cam.is_rh = f->is_rh; cam.width = f->width; cam.height = f->height;

    for (y = 0; y < f->height; y++) {
        for (x = 0; x < f->width; x++) {
            double sum = 0.0;
            for (s = 0; s < f->spp; s++) {
                const uint64_t sid = ((uint64_t)y * (uint64_t)f->width + (uint64_t)x) * (uint64_t)f->spp + (uint64_t)s;
                uint32_t k = 0;
                double org[3], dir[3], t, u, v, G = 1.0, rad;
                uint32_t prim;
                int depth = 2, alive;
                orc_state_f64 st;
                const double jx = path_uniform(f->seed, sid, k++), jy = path_uniform(f->seed, sid, k++);
                orc_camera_ray(&cam, (double)(x + jx), (double)(y + jy), org, dir);
                nrays++;
                if (!trace_f64(T, &V, org, dir, 0, &t, &u, &v, &prim, NULL)) { sum = sum + f->Le; continue; }
                state_build(T, org, dir, t, prim, &st);
                alive = 1;
                while (alive && depth < f->max_vertices) {                       /* trace_path */
                    double out[3], r0, r1;
                    if (path_uniform(f->seed, sid, k++) > f->kd) break;          /* russian roulette */
                    (void)path_uniform(f->seed, sid, k++);                       /* lobe pick: always 'D' */
                    r0 = path_uniform(f->seed, sid, k++); r1 = path_uniform(f->seed, sid, k++);
                    path_cosweight(out, st.Ng, r0, r1);
                    nrays++;
                    if (!trace_f64(T, &V, st.P, out, 0, &t, &u, &v, &prim, NULL)) { alive = 0; break; }
                    G = G * (f->kd * inv_pi);
                    depth++;
                    {
                        double o2[3] = { st.P[0], st.P[1], st.P[2] };
                        state_build(T, o2, out, t, prim, &st);
                    }
                }
                {                                                                /* connect to the environment */
                    double out[3], r0, r1;
                    (void)path_uniform(f->seed, sid, k++);
                    r0 = path_uniform(f->seed, sid, k++); r1 = path_uniform(f->seed, sid, k++);
                    path_cosweight(out, st.Ng, r0, r1);
                    G = G * (f->kd * inv_pi);
                    nrays++;
                    rad = trace_f64(T, &V, st.P, out, 0, &t, &u, &v, &prim, NULL) ? 0.0 : f->Le;
                }
                sum = sum + rad * G;
            }
            {
                float *dst = rgb + 3 * ((size_t)(f->height - y - 1) * f->width + x);
                dst[0] = dst[1] = dst[2] = (float)(sum / (double)f->spp);
            }
        }
    }
    if (nrays_out) *nrays_out = nrays;
}

/* counter-based generator shared with the product for the synthetic configs */
uint64_t orc_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* ------------------------------------------------------------------ point-based AO batch (ri_b200_occlusion_points_f32)
 * The ray set-up of calculate_occlusion (transport/ambientocclusion.c:56-117) for a batch of shading points (P, Ns): origin
 * P + eps * Ns, basis = ri_ortho_basis(Ns), outer loop j over phi, inner loop i over theta, z0 = (i + u0) / ntheta,
 * z1 = (j + u1) / nphi, cos(theta) = sqrt(z0), local = (cos(phi) cos(theta), sin(phi) cos(theta), sqrt(1 - cos^2(theta))),
 * dir = sum local[k] * basis[k] -- in double, rounded once to the fp32 ray record [ox oy oz 0 dx dy dz 1e38].  Builder-stated
 * substitutions, shared with the device: u0, u1 = the counter-based uniforms of scenes.uniform01 keyed by
 * (seed, point, j, i) in place of the sequential randomMT2 stream, and sin / cos by orc_det_sincos2pi in place of libm. */
/* the same batch with NO rounding to fp32: [n*ntheta*nphi][6] doubles (org, dir) -- what calculate_occlusion traces in the reference's
 * own precision; the double-exact point entry ri_b200_occlusion_points_f64 is checked against it */
void orc_ao_point_rays_f64(const double *points, uint64_t n, uint64_t first_point, int ntheta, int nphi, uint64_t seed, double eps, double *rays_out)
{
    const uint64_t N = (uint64_t)ntheta * (uint64_t)nphi;
    uint64_t p;
    for (p = 0; p < n; p++) {
        const double *pt = points + 6 * p;
        double basis[3][3], nrm[3];
        uint32_t i, j; int k;
        for (k = 0; k < 3; k++) nrm[k] = pt[3 + k];
        ortho_basis_f64(basis, nrm);
        for (j = 0; j < (uint32_t)nphi; j++) {
            for (i = 0; i < (uint32_t)ntheta; i++) {
                const uint64_t kk = (uint64_t)j * (uint64_t)ntheta + i, idx = ((first_point + p) * N + kk) * 2;
                const double u0 = (double)(orc_splitmix64(seed + idx * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
                const double u1 = (double)(orc_splitmix64(seed + (idx + 1) * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
                const double z0 = ((double)i + u0) / (double)ntheta;
                const double z1 = ((double)j + u1) / (double)nphi;
                const double ct = sqrt(z0);
                double sn, cs, lx, ly, lz;
                double *o = rays_out + 6 * (p * N + kk);
                orc_det_sincos2pi(z1, &sn, &cs);
                lx = cs * ct; ly = sn * ct; lz = sqrt(1.0 - ct * ct);
                for (k = 0; k < 3; k++) {
                    o[k] = pt[k] + nrm[k] * eps;
                    o[3 + k] = lx * basis[0][k] + ly * basis[1][k] + lz * basis[2][k];
                }
            }
        }
    }
}

void orc_ao_point_rays_f32(const double *points, uint64_t n, uint64_t first_point, int ntheta, int nphi, uint64_t seed, double eps, float *rays_out)
{
    const uint64_t N = (uint64_t)ntheta * (uint64_t)nphi;
    uint64_t p;
    for (p = 0; p < n; p++) {
        const double *pt = points + 6 * p;
        double basis[3][3], nrm[3];
        uint32_t i, j; int k;
        for (k = 0; k < 3; k++) nrm[k] = pt[3 + k];
        ortho_basis_f64(basis, nrm);
        for (j = 0; j < (uint32_t)nphi; j++) {
            for (i = 0; i < (uint32_t)ntheta; i++) {
                const uint64_t kk = (uint64_t)j * (uint64_t)ntheta + i, idx = ((first_point + p) * N + kk) * 2;      /* keyed by the point's position in the whole batch */
                const double u0 = (double)(orc_splitmix64(seed + idx * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
                const double u1 = (double)(orc_splitmix64(seed + (idx + 1) * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
                const double z0 = ((double)i + u0) / (double)ntheta;
                const double z1 = ((double)j + u1) / (double)nphi;
                const double ct = sqrt(z0);
                double sn, cs, lx, ly, lz;
                float *o = rays_out + 8 * (p * N + kk);
                orc_det_sincos2pi(z1, &sn, &cs);
                lx = cs * ct; ly = sn * ct; lz = sqrt(1.0 - ct * ct);
                for (k = 0; k < 3; k++) {
                    o[k] = (float)(pt[k] + nrm[k] * eps);
                    o[4 + k] = (float)(lx * basis[0][k] + ly * basis[1][k] + lz * basis[2][k]);
                }
                o[3] = 0.0f; o[7] = 1.0e38f;
            }
        }
    }
}
