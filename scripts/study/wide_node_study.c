/* Analysis tool (not product, not oracle): how many traversal steps would occlusion (any-hit) rays take if two levels of the
 * reference's binary BVH were walked per step?  The answer of an occlusion query does not depend on the visiting order and the
 * boxes of the tree nest, so skipping the intermediate box test reaches exactly the same leaves (what is next).
 * Input: the oracle's node array (DFS preorder, orc_node_t), triangles in post-build order, rays (org, dir) in double.
 * Counts, per ray on average: binary inner visits / leaf visits / triangle tests (must equal the reference's counters), and for the
 * two-level walk the visits to even-depth inner nodes (= steps), the child boxes tested, and the same leaf / triangle numbers. */
#include <stdint.h>
#include <math.h>
#include <string.h>

typedef struct { int32_t is_leaf, axis; int64_t child0, child1, tri_start, ntris; double lbox[6], rbox[6]; } node_t;

static int slab(const double *b, const double *org, const double *inv, const int *sg, double tbest)
{
    double tmin, tmax, a, c;
    a = ((sg[0] ? b[3] : b[0]) - org[0]) * inv[0]; c = ((sg[0] ? b[0] : b[3]) - org[0]) * inv[0]; tmin = a; tmax = c;
    a = ((sg[1] ? b[4] : b[1]) - org[1]) * inv[1]; c = ((sg[1] ? b[1] : b[4]) - org[1]) * inv[1]; if (a > tmin) tmin = a; if (c < tmax) tmax = c;
    a = ((sg[2] ? b[5] : b[2]) - org[2]) * inv[2]; c = ((sg[2] ? b[2] : b[5]) - org[2]) * inv[2]; if (a > tmin) tmin = a; if (c < tmax) tmax = c;
    return tmax > 0.0 && tmin <= tmax && tmin < tbest;
}

static int tri_hit(const double *v, const double *org, const double *dir)
{
    double e1[3], e2[3], p[3], s[3], q[3], det, inv, u, w, t;
    int k;
    for (k = 0; k < 3; k++) { e1[k] = v[3 + k] - v[k]; e2[k] = v[6 + k] - v[k]; }
    p[0] = dir[1] * e2[2] - dir[2] * e2[1]; p[1] = dir[2] * e2[0] - dir[0] * e2[2]; p[2] = dir[0] * e2[1] - dir[1] * e2[0];
    det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (fabs(det) <= 1.0e-14) return 0;
    inv = 1.0 / det;
    for (k = 0; k < 3; k++) s[k] = org[k] - v[k];
    u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv;
    if (u < 0.0 || u > 1.0) return 0;
    q[0] = s[1] * e1[2] - s[2] * e1[1]; q[1] = s[2] * e1[0] - s[0] * e1[2]; q[2] = s[0] * e1[1] - s[1] * e1[0];
    w = (dir[0] * q[0] + dir[1] * q[1] + dir[2] * q[2]) * inv;
    if (w < 0.0 || u + w > 1.0) return 0;
    t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
    return !(t < 0.0) && !(t > 1.0e38);
}

/* out[0..2]: binary inner visits, leaf visits, triangle tests; out[3..4]: two-level steps, child boxes tested; out[5]: occluded rays */
void wide_node_study(const node_t *nodes, const double *tris /* [n][9] post-build order */, const double *rays, uint64_t nrays, double *out)
{
    uint64_t r, inner = 0, leaves = 0, tests = 0, steps2 = 0, boxes2 = 0, occluded = 0;
    static int64_t stack[256]; static int depth_of[256];
    for (r = 0; r < nrays; r++) {
        const double *org = rays + 6 * r, *dir = org + 3;
        double inv[3]; int sg[3], k, sp = 0, hit = 0;
        int64_t cur = 0; int depth = 0;
        for (k = 0; k < 3; k++) { sg[k] = dir[k] < 0.0; inv[k] = fabs(dir[k]) > 1.0e-14 ? 1.0 / dir[k] : (dir[k] < 0.0 ? -1.7976931348623157e308 : 1.7976931348623157e308); }
        for (;;) {
            const node_t *n = nodes + cur;
            if (n->is_leaf) {
                int64_t i;
                leaves++;
                for (i = 0; i < n->ntris; i++) { tests++; if (tri_hit(tris + 9 * (n->tri_start + i), org, dir)) hit = 1; }
                if (hit) break;                                     /* the leaf is finished, then the query ends (bvh.c:850) */
            } else {
                const int h0 = slab(n->lbox, org, inv, sg, 1.0e38), h1 = slab(n->rbox, org, inv, sg, 1.0e38);
                const int order = sg[n->axis];
                inner++;
                if ((depth & 1) == 0) {                             /* a two-level step starts here: its own 2 boxes + up to 4 below */
                    steps2++; boxes2 += 2;
                    if (!nodes[n->child0].is_leaf) boxes2 += 2;     /* grandchildren tested in the same step, hit or not */
                    if (!nodes[n->child1].is_leaf) boxes2 += 2;
                }
                if (h0 && h1) {
                    stack[sp] = order ? n->child0 : n->child1; depth_of[sp] = depth + 1; sp++;
                    cur = order ? n->child1 : n->child0; depth++; continue;
                } else if (h0 || h1) { cur = h0 ? n->child0 : n->child1; depth++; continue; }
            }
            if (sp == 0) break;
            sp--; cur = stack[sp]; depth = depth_of[sp];
        }
        occluded += hit;
    }
    out[0] = (double)inner / nrays; out[1] = (double)leaves / nrays; out[2] = (double)tests / nrays;
    out[3] = (double)steps2 / nrays; out[4] = (double)boxes2 / nrays; out[5] = (double)occluded / nrays;
}
