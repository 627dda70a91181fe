/*
 * ri_b200_binding.c -- the lucille-side binding for the B200 accelerator (what a lucille maintainer adds under
 * src/render/).  It implements the three ri_accel_t entry points (src/render/accel.h:24-34) on top of the C ABI of
 * include/lucille_b200.h and registers them for accel method RI_ACCEL_B200 (= 2).
 *
 * This file is compiled against the reference's own headers.  In this repository it is built ONLY as test
 * infrastructure (oracle/build_ref.sh -> oracle/_ref/lsh_b200) to prove the drop-in: the unmodified reference renderer
 * -- its RIB front end substitute, Ri layer, pixel loop, AO transport and MT19937 -- runs with every ri_raytrace()
 * answered by the GPU, and the frame is bit-identical to the pure-CPU reference frame.
 *
 * In a lucille checkout the `__wrap_` indirection below is replaced by two lines in ri_accel_bind();
 * here the reference sources must stay untouched, so the selector is interposed at link time with
 * `-Wl,--wrap=ri_accel_bind`.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "ri.h"
#include "render.h"
#include "scene.h"
#include "geom.h"
#include "accel.h"
#include "bvh.h"
#include "ray.h"
#include "intersection_state.h"
#include "list.h"
#include "log.h"
#include "memory.h"

#include "lucille_b200.h"

#include "ri_b200_binding.h"

/* accel_build_func: `data` is the ri_scene_t (scene.c:96,164).  Flattens the geoms exactly like create_triangle_list()
 * (bvh.c:1736-1826): geoms in list order, triangles in index order, positions already in world space. */
void *ri_b200_accel_build(const void *data)
{
    const ri_scene_t *scene = (const ri_scene_t *)data;
    b200_binding_t *b = (b200_binding_t *)ri_mem_alloc(sizeof(b200_binding_t));
    ri_list_t *itr;
    uint64_t n = 0, idx = 0;
    double *xyz, *nrm;
    int any_normals = 0;
    unsigned int i;
    int k, c;

    memset(b, 0, sizeof(*b));
    for (itr = ri_list_first(scene->geom_list); itr; itr = ri_list_next(itr))
        n += ((ri_geom_t *)itr->data)->nindices / 3;
    b->ntris = n;
    xyz          = (double *)malloc(sizeof(double) * 9 * (n ? n : 1));
    nrm          = (double *)calloc(9 * (n ? n : 1), sizeof(double));
    b->tri_geom  = (ri_geom_t **)malloc(sizeof(ri_geom_t *) * (n ? n : 1));
    b->tri_index = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    b->orig      = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (itr = ri_list_first(scene->geom_list); itr; itr = ri_list_next(itr)) {
        ri_geom_t *geom = (ri_geom_t *)itr->data;
        for (i = 0; i < geom->nindices / 3; i++) {
            for (c = 0; c < 3; c++)
                for (k = 0; k < 3; k++)
                {
                    xyz[9 * idx + 3 * c + k] = geom->positions[geom->indices[3 * i + c]][k];
                    if (geom->normals) { nrm[9 * idx + 3 * c + k] = geom->normals[geom->indices[3 * i + c]][k]; any_normals = 1; }
                }
            b->tri_geom[idx]  = geom;
            b->tri_index[idx] = 3 * i;
            idx++;
        }
    }
    ri_log(LOG_INFO, "(B200  ) Building accelerator on the GPU side: %llu triangles", (unsigned long long)n);
    b->dev = ri_b200_build(xyz, n, RI_B200_PREC_F64 | RI_B200_PREC_F32, 0);
    free(xyz);
    if (b->dev && any_normals) ri_b200_set_normals(b->dev, nrm);     /* used by the batched frame path */
    free(nrm);
    if (!b->dev) {                       /* no error channel in the reference: log and abort (memory.c:87-97 style) */
        ri_log(LOG_FATAL, "(B200  ) %s", ri_b200_last_error());
        abort();
    }
    if (n) ri_b200_triorder(b->dev, b->orig);
    return b;
}

/* accel_free_func (accel.h:27-28) */
void ri_b200_accel_free(void *accel)
{
    b200_binding_t *b = (b200_binding_t *)accel;
    if (!b) return;
    ri_b200_free(b->dev);
    free(b->tri_geom); free(b->tri_index); free(b->orig);
    ri_mem_free(b);
}

/* accel_intersect_func (accel.h:30-34): one ray, called concurrently from the reference's worker threads.
 * Returns 1 with a fully built ri_intersection_state_t on hit (raytrace.c:64-66 copies it verbatim), else 0. */
int ri_b200_accel_intersect(void *accel, ri_ray_t *ray, ri_intersection_state_t *state, void *user)
{
    b200_binding_t *b = (b200_binding_t *)accel;
    ri_b200_hit_f64 hit;
    int rc;
    if (user) {
        /* `user` is an ri_bvh_diag_t* from the testbed (simplerender.cpp:202).  ri_bvh_intersect zeroes it (bvh.c:451-456) and, in a
         * build with RI_BVH_ENABLE_DIAGNOSTICS (bvh.h:38, off by default), counts inner-node visits, leaf visits and leaf calls
         * (bvh.c:826-828, 1126-1147).  The accelerator has those counters in the reference's own traversal order, so it fills them. */
        ri_bvh_diag_t *diag = (ri_bvh_diag_t *)user;
        ri_b200_counters_t cnt;
        double r6[6];
        memset(diag, 0, sizeof(*diag));
        for (rc = 0; rc < 3; rc++) { r6[rc] = ray->org[rc]; r6[3 + rc] = ray->dir[rc]; }
        if (ri_b200_count_batch(b->dev, r6, 1, RI_B200_PREC_F64, 0, &cnt) == 0) {
            diag->ninner_node_traversals = (uint32_t)cnt.ninner;
            diag->nleaf_node_traversals  = (uint32_t)cnt.nleaf;
            diag->ntriangle_isects       = (uint32_t)cnt.nleaf;      /* incremented once per bvh_intersect_leaf_node call, bvh.c:826-828 */
        }
    }
    rc = ri_b200_intersect1(b->dev, ray->org, ray->dir, &hit, NULL);
    if (rc < 0) { ri_log(LOG_FATAL, "(B200  ) %s", ri_b200_last_error()); abort(); }
    if (rc == 0) return 0;
    {
        const uint32_t src = b->orig[hit.prim];
        state->t = hit.t; state->u = hit.u; state->v = hit.v;          /* bvh.c:855-859 */
        state->geom  = b->tri_geom[src];
        state->index = b->tri_index[src];
        ri_intersection_state_build(state, ray->org, ray->dir);        /* bvh.c:537-539: the reference's own post-hit code */
    }
    return 1;
}

/* selector.  In lucille proper: a `case RI_ACCEL_B200:` in ri_accel_bind() (accel.c:80-106). */
extern int __real_ri_accel_bind(ri_accel_t *accel, int method);
int __wrap_ri_accel_bind(ri_accel_t *accel, int method)
{
    if (method == RI_ACCEL_B200) {
        accel->build     = ri_b200_accel_build;
        accel->free      = ri_b200_accel_free;
        accel->intersect = ri_b200_accel_intersect;
        return 0;
    }
    return __real_ri_accel_bind(accel, method);
}
