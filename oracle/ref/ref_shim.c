/*
 * TEST INFRASTRUCTURE (oracle/_ref) -- not part of the product.
 *
 * Thin C wrapper around the UNMODIFIED lucille reference libraries so that Python
 * (ctypes) can (a) run ray batches through the reference's ri_bvh_intersect(),
 * (b) dump the reference-built BVH, (c) read the reference's traversal counters and
 * (d) render a RIB through the reference's own Ri*() -> ri_render_frame() path with
 * float pixels captured by ri_dd_callback().
 *
 * Embedding pattern follows the reference's own testbed
 * (src/testbed/main.cpp:53-65, src/testbed/simplerender.cpp:170-238) and lsh
 * (src/lsh/main.c:105-242).  All arithmetic below this file is reference code.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

#include "ri.h"
#include "render.h"
#include "scene.h"
#include "geom.h"
#include "accel.h"
#include "bvh.h"
#include "camera.h"
#include "option.h"
#include "display.h"
#include "backdoor.h"
#include "parallel.h"
#include "timer.h"
#include "list.h"
#include "log.h"
#include "raytrace.h"
#include "beam.h"
#include "light.h"
#include "sunsky.h"

static void capture_sunsky(void);

/* ---- material texture (ambientocclusion.c:393-401): a float RGBA image handed over by the test driver becomes the
 * material->texture of every geom of the frame, exactly the object Surface "..." "texture" [...] would have loaded
 * (ri/attribute.c:309-327), without going through the image readers. */
#include "texture.h"
#include "material.h"
static ri_texture_t g_frame_tex;
static int          g_have_frame_tex = 0;
void lref_set_frame_texture(const float *rgba, int width, int height)
{
    g_have_frame_tex = 0;
    if (!rgba || width < 1 || height < 1) return;
    memset(&g_frame_tex, 0, sizeof(g_frame_tex));
    g_frame_tex.data = (float *)malloc(sizeof(float) * 4 * (size_t)width * height);
    memcpy(g_frame_tex.data, rgba, sizeof(float) * 4 * (size_t)width * height);
    g_frame_tex.width = width; g_frame_tex.height = height;
    g_have_frame_tex = 1;
}
static void attach_frame_texture(ri_geom_t *geom)
{
    if (!g_have_frame_tex) return;
    if (!geom->material) geom->material = ri_material_new();
    geom->material->texture = &g_frame_tex;
}
/* ri_texture_fetch (render/texture.c:86-236) on a caller-supplied image: uv [n][2] doubles -> out [n][4] doubles */
void lref_texture_fetch(const float *rgba, int width, int height, const double *uv, uint64_t n, double *out)
{
    ri_texture_t t;
    uint64_t i;
    memset(&t, 0, sizeof(t));
    t.data = (float *)rgba; t.width = width; t.height = height;
    for (i = 0; i < n; i++) {
        ri_vector_t c;
        ri_texture_fetch(c, &t, uv[2 * i], uv[2 * i + 1]);
        out[4 * i] = c[0]; out[4 * i + 1] = c[1]; out[4 * i + 2] = c[2]; out[4 * i + 3] = c[3];
    }
}

extern int lref_rib_parse_file(const char *path);

/* ------------------------------------------------------------------ ray level */

typedef struct {
    int32_t  hit;
    uint32_t index;          /* state.index = 3 * (triangle number in its geom), bvh.c:1813 */
    uint32_t geom_id;
    uint32_t pad;
    double   t, u, v;
    double   P[3], Ng[3], Ns[3], tangent[3], binormal[3];
} lref_hit_t;

typedef struct {
    int32_t is_leaf;
    int32_t axis;
    int64_t child0, child1;      /* indices into the DFS-preorder dump (inner nodes) */
    int64_t tri_start, ntris;    /* leaves: slice of the post-build triangle array */
    double  lbox[6];             /* slot 0: min xyz, max xyz */
    double  rbox[6];             /* slot 1 */
} lref_node_t;

typedef struct {
    ri_scene_t    *scene;
    ri_geom_t    **geoms;
    int            ngeoms;
    ri_bvh_t      *bvh;
    ri_triangle_t *tri_base;
    uint64_t       ntris;
    double         build_seconds;
} lref_scene_t;

static int g_inited = 0;

static double now_sec(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int lref_init(void)
{
    if (!g_inited) {
        int argc = 1; char *argv0 = "lref"; char **argv = &argv0;
        ri_parallel_init(&argc, &argv);
        ri_render_init();
        ri_log_set_debug(0);
        g_inited = 1;
    }
    return 0;
}

/* tri_xyz: [ntris][3][3] doubles. geom_sizes: triangles per geom (ngeoms entries) or NULL for one geom.
 * Each geom gets unshared vertices: positions 3 per triangle, indices 0..3n-1, exactly what
 * create_triangle_list (bvh.c:1736-1826) flattens again. */
void *lref_scene_build(const double *tri_xyz, uint64_t ntris, const uint64_t *geom_sizes, int ngeoms)
{
    lref_scene_t *s;
    uint64_t one = ntris;
    uint64_t off = 0;
    int g;
    double t0;

    lref_init();
    s = (lref_scene_t *)calloc(1, sizeof(lref_scene_t));
    s->scene = ri_scene_new();
    if (!geom_sizes) { geom_sizes = &one; ngeoms = 1; }
    s->geoms = (ri_geom_t **)calloc(ngeoms > 0 ? ngeoms : 1, sizeof(ri_geom_t *));
    s->ngeoms = ngeoms;
    s->ntris = ntris;

    for (g = 0; g < ngeoms; g++) {
        uint64_t n = geom_sizes[g], i;
        ri_geom_t *geom = ri_geom_new();
        if (n > 0) {
            ri_vector_t  *pos = (ri_vector_t *)malloc(sizeof(ri_vector_t) * 3 * n);
            unsigned int *idx = (unsigned int *)malloc(sizeof(unsigned int) * 3 * n);
            for (i = 0; i < 3 * n; i++) {
                pos[i][0] = tri_xyz[3 * (3 * off + i) + 0];
                pos[i][1] = tri_xyz[3 * (3 * off + i) + 1];
                pos[i][2] = tri_xyz[3 * (3 * off + i) + 2];
                pos[i][3] = 1.0;
                idx[i] = (unsigned int)i;
            }
            ri_geom_add_positions(geom, (unsigned int)(3 * n), (const ri_vector_t *)pos);
            ri_geom_add_indices(geom, (unsigned int)(3 * n), idx);
            free(pos); free(idx);
        }
        s->geoms[g] = geom;
        ri_scene_add_geom(s->scene, geom);
        off += n;
    }

    {
        ri_accel_t *accel = ri_accel_new();
        ri_accel_bind(accel, RI_ACCEL_BVH);
        ri_scene_set_accel(s->scene, accel);
        t0 = now_sec();
        ri_scene_build_accel(s->scene);
        s->build_seconds = now_sec() - t0;
        s->bvh = (ri_bvh_t *)s->scene->accel->data;
    }
    if (!s->bvh->empty) {
        const ri_qbvh_node_t *n = s->bvh->root;
        while (!n->is_leaf) n = n->child[0];           /* leftmost leaf starts the array */
        s->tri_base = (ri_triangle_t *)n->child[0];
    }
    return s;
}

double lref_build_seconds(void *h) { return ((lref_scene_t *)h)->build_seconds; }
int    lref_is_empty(void *h)      { return ((lref_scene_t *)h)->bvh->empty; }

void lref_scene_bbox(void *h, double *bmin, double *bmax)
{
    lref_scene_t *s = (lref_scene_t *)h;
    int k;
    for (k = 0; k < 3; k++) { bmin[k] = s->bvh->bmin[k]; bmax[k] = s->bvh->bmax[k]; }
}

static void count_nodes(const ri_qbvh_node_t *n, int depth, int64_t *ninner, int64_t *nleaf, int *maxdepth)
{
    if (depth > *maxdepth) *maxdepth = depth;
    if (n->is_leaf) { (*nleaf)++; return; }
    (*ninner)++;
    count_nodes(n->child[0], depth + 1, ninner, nleaf, maxdepth);
    count_nodes(n->child[1], depth + 1, ninner, nleaf, maxdepth);
}

void lref_tree_count(void *h, int64_t *ninner, int64_t *nleaf, int *maxdepth)
{
    lref_scene_t *s = (lref_scene_t *)h;
    *ninner = 0; *nleaf = 0; *maxdepth = 0;
    if (s->bvh->empty) return;
    count_nodes(s->bvh->root, 0, ninner, nleaf, maxdepth);
}

static int64_t dump_node(lref_scene_t *s, const ri_qbvh_node_t *n, lref_node_t *out, int64_t *cursor)
{
    int64_t me = (*cursor)++;
    lref_node_t *o = &out[me];
    memset(o, 0, sizeof(*o));
    o->is_leaf = n->is_leaf;
    if (n->is_leaf) {
        o->ntris = *((uint32_t *)&n->bbox[0]);
        o->tri_start = (ri_triangle_t *)n->child[0] - s->tri_base;
        o->child0 = o->child1 = -1;
        return me;
    }
    o->axis = n->axis0;
    o->lbox[0] = n->bbox[BMIN_X0]; o->lbox[1] = n->bbox[BMIN_Y0]; o->lbox[2] = n->bbox[BMIN_Z0];
    o->lbox[3] = n->bbox[BMAX_X0]; o->lbox[4] = n->bbox[BMAX_Y0]; o->lbox[5] = n->bbox[BMAX_Z0];
    o->rbox[0] = n->bbox[BMIN_X1]; o->rbox[1] = n->bbox[BMIN_Y1]; o->rbox[2] = n->bbox[BMIN_Z1];
    o->rbox[3] = n->bbox[BMAX_X1]; o->rbox[4] = n->bbox[BMAX_Y1]; o->rbox[5] = n->bbox[BMAX_Z1];
    o->child0 = dump_node(s, n->child[0], out, cursor);
    o = &out[me];
    o->child1 = dump_node(s, n->child[1], out, cursor);
    return me;
}

/* DFS preorder dump; `out` must hold ninner+nleaf entries. */
int64_t lref_tree_dump(void *h, lref_node_t *out)
{
    lref_scene_t *s = (lref_scene_t *)h;
    int64_t cursor = 0;
    if (s->bvh->empty) return 0;
    dump_node(s, s->bvh->root, out, &cursor);
    return cursor;
}

/* for each position p of the post-build triangle array: flattened input triangle number */
void lref_tree_triorder(void *h, uint32_t *orig)
{
    lref_scene_t *s = (lref_scene_t *)h;
    uint64_t p;
    uint64_t *geom_off = (uint64_t *)calloc(s->ngeoms + 1, sizeof(uint64_t));
    int g;
    for (g = 0; g < s->ngeoms; g++) geom_off[g + 1] = geom_off[g] + s->geoms[g]->nindices / 3;
    for (p = 0; p < s->ntris; p++) {
        const ri_triangle_t *t = &s->tri_base[p];
        for (g = 0; g < s->ngeoms; g++) if (s->geoms[g] == t->geom) break;
        orig[p] = (uint32_t)(geom_off[g] + t->index / 3);
    }
    free(geom_off);
}

typedef struct {
    lref_scene_t *s;
    const double *rays;      /* [n][6] org.xyz dir.xyz */
    lref_hit_t   *out;
    uint64_t      begin, end;
    int           want_state;
} job_t;

static void *job_run(void *arg)
{
    job_t *j = (job_t *)arg;
    uint64_t i;
    int g, k;
    ri_ray_t ray;
    ri_intersection_state_t state;
    memset(&ray, 0, sizeof(ray));
    for (i = j->begin; i < j->end; i++) {
        const double *r = j->rays + 6 * i;
        lref_hit_t *o = j->out ? &j->out[i] : NULL;
        int hit;
        ray.org[0] = r[0]; ray.org[1] = r[1]; ray.org[2] = r[2]; ray.org[3] = 1.0;
        ray.dir[0] = r[3]; ray.dir[1] = r[4]; ray.dir[2] = r[5]; ray.dir[3] = 0.0;
        hit = ri_bvh_intersect(j->s->bvh, &ray, &state, NULL);
        if (!o) continue;
        o->hit = hit;
        if (hit) {
            o->index = state.index;
            for (g = 0; g < j->s->ngeoms; g++) if (j->s->geoms[g] == state.geom) break;
            o->geom_id = (uint32_t)g;
            o->t = state.t; o->u = state.u; o->v = state.v;
            for (k = 0; k < 3; k++) {
                o->P[k] = state.P[k]; o->Ng[k] = state.Ng[k]; o->Ns[k] = state.Ns[k];
                o->tangent[k] = state.tangent[k]; o->binormal[k] = state.binormal[k];
            }
        } else {
            memset(&o->index, 0, sizeof(*o) - sizeof(int32_t));
        }
    }
    return NULL;
}

/* returns seconds spent inside the query loop (build excluded) */
double lref_intersect(void *h, const double *rays, uint64_t n, int nthreads, lref_hit_t *out)
{
    lref_scene_t *s = (lref_scene_t *)h;
    pthread_t th[256];
    job_t jobs[256];
    int t;
    double t0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    t0 = now_sec();
    for (t = 0; t < nthreads; t++) {
        jobs[t].s = s; jobs[t].rays = rays; jobs[t].out = out;
        jobs[t].begin = n * (uint64_t)t / nthreads;
        jobs[t].end   = n * (uint64_t)(t + 1) / nthreads;
        if (nthreads == 1) job_run(&jobs[t]);
        else pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    if (nthreads > 1) for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return now_sec() - t0;
}

/* ---- vertex colours, texture coordinates, two-sided flag (intersection_state.c:192-246) ------------------------------------
 * colors: [ntris][3][3], st: [ntris][3][2] per corner, in the input triangle order (the shim's geoms have unshared vertices,
 * 3 per triangle, so a corner IS a vertex).  geom_flags[g]: bit0 the geom has Cs, bit1 shared texcoords, bit2 unshared texcoords,
 * bit3 two_side (polygon.c:595-612 doubles the faces; here the second half of the geom's triangles counts as the back side). */
void lref_scene_set_attr(void *h, const double *colors, const double *st, const uint8_t *geom_flags)
{
    lref_scene_t *s = (lref_scene_t *)h;
    uint64_t off = 0;
    int g;
    for (g = 0; g < s->ngeoms; g++) {
        ri_geom_t *geom = s->geoms[g];
        const uint64_t n = geom->nindices / 3;
        uint64_t i;
        if (n && colors && (geom_flags[g] & 1)) {
            ri_vector_t *c = (ri_vector_t *)malloc(sizeof(ri_vector_t) * 3 * n);
            for (i = 0; i < 3 * n; i++) { c[i][0] = colors[3 * (3 * off + i)]; c[i][1] = colors[3 * (3 * off + i) + 1]; c[i][2] = colors[3 * (3 * off + i) + 2]; c[i][3] = 1.0; }
            ri_geom_add_colors(geom, (unsigned int)(3 * n), (const ri_vector_t *)c);
            free(c);
        }
        if (n && st && (geom_flags[g] & 6)) {
            ri_float_t *t = (ri_float_t *)malloc(sizeof(ri_float_t) * 2 * 3 * n);
            for (i = 0; i < 2 * 3 * n; i++) t[i] = st[2 * 3 * off + i];
            if (geom_flags[g] & 2) ri_geom_add_texcoords(geom, (unsigned int)(3 * n), t);
            else ri_geom_add_texcoords_unshared(geom, (unsigned int)(3 * n), t);
            free(t);
        }
        geom->two_side = (geom_flags[g] & 8) ? 1 : 0;
        off += n;
    }
}

typedef struct { double E[3], I[3], color[3], st[2], t; int32_t inside, hit; } lref_ext_t;
void lref_intersect_ext(void *h, const double *rays, uint64_t n, lref_ext_t *out)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_ray_t ray;
    ri_intersection_state_t state;
    uint64_t i;
    int k;
    memset(&ray, 0, sizeof(ray));
    for (i = 0; i < n; i++) {
        const double *r = rays + 6 * i;
        lref_ext_t *o = &out[i];
        ray.org[0] = r[0]; ray.org[1] = r[1]; ray.org[2] = r[2]; ray.org[3] = 1.0;
        ray.dir[0] = r[3]; ray.dir[1] = r[4]; ray.dir[2] = r[5]; ray.dir[3] = 0.0;
        memset(o, 0, sizeof(*o));
        o->hit = ri_bvh_intersect(s->bvh, &ray, &state, NULL);
        if (!o->hit) continue;
        for (k = 0; k < 3; k++) { o->E[k] = state.E[k]; o->I[k] = state.I[k]; o->color[k] = state.color[k]; }
        o->st[0] = state.stqr[0]; o->st[1] = state.stqr[1];
        o->t = state.t; o->inside = state.inside;
    }
}

/* ---- the other transports: compiled into the reference but never called by its pixel loop (render.c:800-804).
 * Called here per eye ray exactly as subsample() would call them: render = ri_render_get() with this scene installed, thread 0,
 * the thread's MT19937 stream re-seeded with 4357 (random.c:98-112) before the batch.  which: 0 ambient occlusion, 1 dirt map, 2 Whitted refraction tracer, 3 ri_transport_sample (transport.c:50-173: white on a hit). */
#include "transport.h"
extern int ri_transport_dirtmap(ri_render_t *render, const ri_ray_t *ray, ri_transport_info_t *result);
extern int ri_transport_ambientocclusion(ri_render_t *render, const ri_ray_t *ray, ri_transport_info_t *result);
extern void seedMT2();
extern int ri_transport_whitted(ri_render_t *render, const ri_ray_t *ray, ri_transport_info_t *result);
/* environment map of the scene (scene->envmap_light->texture, read by whitted.c:72-77,135-140 through ri_texture_ibl_fetch) */
void lref_set_envmap(void *h, const float *rgba, int width, int height)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_light_t *l = ri_light_new();
    ri_texture_t *t = (ri_texture_t *)calloc(1, sizeof(ri_texture_t));
    t->data = (float *)malloc(sizeof(float) * 4 * (size_t)width * height);
    memcpy(t->data, rgba, sizeof(float) * 4 * (size_t)width * height);
    t->width = width; t->height = height;
    l->texture = t;
    s->scene->envmap_light = l;
}
void lref_transport_batch(void *h, int which, const double *rays, uint64_t n, double *radiance3)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_render_t *render = ri_render_get();
    ri_scene_t *saved = render->scene;
    ri_ray_t ray;
    ri_transport_info_t result;
    uint64_t i;
    render->scene = s->scene;
    seedMT2((unsigned long)4357, 0);
    memset(&ray, 0, sizeof(ray));
    for (i = 0; i < n; i++) {
        const double *r = rays + 6 * i;
        ray.org[0] = r[0]; ray.org[1] = r[1]; ray.org[2] = r[2]; ray.org[3] = 1.0;
        ray.dir[0] = r[3]; ray.dir[1] = r[4]; ray.dir[2] = r[5]; ray.dir[3] = 0.0;
        ray.thread_num = 0;
        memset(&result, 0, sizeof(result));
        if (which == 3) ri_transport_sample(render, &ray, &result);
        else if (which == 2) ri_transport_whitted(render, &ray, &result);
        else if (which == 1) ri_transport_dirtmap(render, &ray, &result);
        else ri_transport_ambientocclusion(render, &ray, &result);
        radiance3[3 * i] = result.radiance[0]; radiance3[3 * i + 1] = result.radiance[1]; radiance3[3 * i + 2] = result.radiance[2];
    }
    render->scene = saved;
}

/* The reference's own per-point gathers: occlusion() shadeop (shader.c:680-768), ri_ibl_sample_cosweight (ibl.c:53-228, needs
 * lref_set_envmap first), ri_domelight_sample (ibl.c:231-389).  points = [n][6] (P, N). */
#include "shader.h"
#include "ibl.h"
extern void seedMT(unsigned long seed);
void lref_point_gather_ex(void *h, int kind, int nsamples, const double *points, uint64_t n, const double *col3, double intensity,
                          int use_qmc, const int32_t *instance, int dim, double *out3);
void lref_point_gather(void *h, int kind, int nsamples, const double *points, uint64_t n, const double *col3, double intensity,
                       double *out3)
{ lref_point_gather_ex(h, kind, nsamples, points, n, col3, intensity, 0, NULL, 0, out3); }
/* use_qmc: Option "use_qmc" for the duration of the call; instance[p] -> inray->i, dim -> inray->d */
void lref_point_gather_ex(void *h, int kind, int nsamples, const double *points, uint64_t n, const double *col3, double intensity,
                          int use_qmc, const int32_t *instance, int dim, double *out3)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_render_t *render = ri_render_get();
    ri_scene_t *saved = render->scene;
    static ri_hemisphere_t hemi;
    ri_status_t status;
    ri_ray_t inray;
    ri_light_t *dome = ri_light_new();
    ri_vector_t P, N, eye, power;
    uint64_t i;
    int k;
    ri_option_t *opt = render->context->option;
    const int saved_qmc = opt->use_qmc;
    opt->use_qmc = use_qmc;
    render->scene = s->scene;
    seedMT((unsigned long)4357);
    seedMT2((unsigned long)4357, 0);
    memset(&status, 0, sizeof(status));
    memset(&inray, 0, sizeof(inray));
    memset(eye, 0, sizeof(eye));
    if (col3) { for (k = 0; k < 3; k++) dome->col[k] = col3[k]; dome->col[3] = 1.0; }
    dome->intensity = intensity;
    for (i = 0; i < n; i++) {
        for (k = 0; k < 3; k++) { P[k] = points[6 * i + k]; N[k] = points[6 * i + 3 + k]; }
        P[3] = 1.0; N[3] = 0.0;
        inray.i = instance ? instance[i] : 0; inray.d = dim;
        if (kind == 0) {
            power[0] = power[1] = power[2] = (double)occlusion(&status, P, N, (float)nsamples);
        } else if (kind == 1) {
            ri_ibl_sample_cosweight(power, N, nsamples, &inray, P, eye, s->scene->envmap_light);
        } else if (kind == 3) {                 /* ri_ibl_sample_bruteforce (ibl.c:395-518): needs a square angular map */
            for (k = 0; k < 4; k++) hemi.basis[2][k] = N[k];
            s->scene->envmap_light->type = LIGHTTYPE_IBL;
            ri_ibl_sample_bruteforce(power, &hemi, nsamples, P, eye, s->scene->envmap_light);
        } else {
            for (k = 0; k < 4; k++) hemi.basis[2][k] = N[k];
            ri_domelight_sample(power, &hemi, nsamples, &inray, P, eye, dome);
        }
        for (k = 0; k < 3; k++) out3[3 * i + k] = power[k];
    }
    opt->use_qmc = saved_qmc;
    render->scene = saved;
}

/* ---- the two shading-language callers of ri_raytrace: trace() (shader.c:895-976) and next_lightsource()
 * (shader.c:1116-1186).  trace() ends in the hit geometry's shader procedure; a capturing procedure installed on every geom records
 * the input block the reference hands it.  Both need the scene's first light (get_light, shader.c:1219-1234): an IBL light carrying
 * the environment map set with lref_set_envmap. */
typedef struct { double Cs[3], P[3], N[3], Ng[3], dPdu[3], dPdv[3], I[3], dst[3]; float s, t; int32_t called, ray_depth; } lref_trace_rec_t;
static lref_trace_rec_t *g_trace_rec = NULL;
static void capture_shaderproc(ri_output_t *output, ri_status_t *status, ri_parameter_t *param)
{
    int k;
    (void)param;
    if (g_trace_rec) {
        for (k = 0; k < 3; k++) {
            g_trace_rec->Cs[k] = status->input.Cs[k]; g_trace_rec->P[k] = status->input.P[k]; g_trace_rec->N[k] = status->input.N[k];
            g_trace_rec->Ng[k] = status->input.Ng[k]; g_trace_rec->dPdu[k] = status->input.dPdu[k]; g_trace_rec->dPdv[k] = status->input.dPdv[k];
            g_trace_rec->I[k] = status->input.I[k];
        }
        g_trace_rec->s = status->input.s; g_trace_rec->t = status->input.t;
        g_trace_rec->called = 1; g_trace_rec->ray_depth = status->ray_depth;
    }
    output->Ci[0] = 0.25; output->Ci[1] = 0.5; output->Ci[2] = 0.75; output->Ci[3] = 1.0;
}
static ri_shader_t g_capture_shader = { NULL, capture_shaderproc, NULL };

static void ensure_ibl_light(lref_scene_t *s)
{
    if (ri_list_first(s->scene->light_list)) return;
    if (s->scene->envmap_light) {
        ri_light_t *l = ri_light_new();
        l->type = LIGHTTYPE_IBL;
        l->texture = s->scene->envmap_light->texture;
        ri_list_append(s->scene->light_list, (void *)l);
    }
}

void lref_shade_trace(void *h, const double *pr, uint64_t n, lref_trace_rec_t *out)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_render_t *render = ri_render_get();
    ri_scene_t *saved = render->scene;
    ri_status_t status;
    uint64_t i;
    int g, k;
    render->scene = s->scene;
    ensure_ibl_light(s);
    for (g = 0; g < s->ngeoms; g++) s->geoms[g]->shader = &g_capture_shader;
    memset(&status, 0, sizeof(status));
    for (i = 0; i < n; i++) {
        ri_vector_t P, R, dst;
        for (k = 0; k < 3; k++) { P[k] = pr[6 * i + k]; R[k] = pr[6 * i + 3 + k]; }
        P[3] = 1.0; R[3] = 0.0;
        memset(&out[i], 0, sizeof(out[i]));
        g_trace_rec = &out[i];
        trace(&status, dst, P, R);
        for (k = 0; k < 3; k++) out[i].dst[k] = dst[k];
    }
    g_trace_rec = NULL;
    render->scene = saved;
}

/* the samples an `illuminance` loop receives from next_lightsource() at each point, in order: L and Cl of every returned sample.
 * count_out[p] = number returned at point p (at most maxm are stored). */
void lref_light_samples(void *h, int nsamples, double angle, const double *points, uint64_t n, int maxm,
                        double *L_out, double *Cl_out, int32_t *count_out)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_render_t *render = ri_render_get();
    ri_scene_t *saved = render->scene;
    ri_option_t *opt = render->context->option;
    const unsigned int saved_rays = opt->narealight_rays;
    ri_status_t status;
    uint64_t i;
    int k;
    render->scene = s->scene;
    ensure_ibl_light(s);
    opt->narealight_rays = (unsigned int)nsamples;
    seedMT((unsigned long)4357);
    seedMT2((unsigned long)4357, 0);
    memset(&status, 0, sizeof(status));
    for (i = 0; i < n; i++) {
        ri_vector_t P, N;
        ri_lightsource_t *l;
        int c = 0;
        for (k = 0; k < 3; k++) { P[k] = points[6 * i + k]; N[k] = points[6 * i + 3 + k]; }
        P[3] = 1.0; N[3] = 0.0;
        while ((l = next_lightsource(&status, P, N, (ri_float_t)angle)) != NULL) {
            if (c < maxm)
                for (k = 0; k < 3; k++) { L_out[3 * ((size_t)maxm * i + c) + k] = l->L[k]; Cl_out[3 * ((size_t)maxm * i + c) + k] = l->Cl[k]; }
            c++;
        }
        count_out[i] = c;
    }
    opt->narealight_rays = saved_rays;
    render->scene = saved;
}

/* The byte stream the reference's socket display driver (display/sockdrv.c) sends for a frame: a listener on 127.0.0.1:DEFAULT_PORT in
 * this process stands in for the viewer, sock_dd_open/write/close are called the way bucket_write does (render.c:919-979: pixels in
 * `pixels` order, (x, y) = display coordinates), everything received is returned.  Returns the number of bytes, -1 on failure. */
#include <sys/socket.h>
#include <netinet/in.h>
#include <arpa/inet.h>
#include "sockdrv.h"
#include "sockdrv_defs.h"
typedef struct { int lfd; unsigned char *buf; size_t cap, len; } sock_capture_t;
static void *sock_capture_main(void *arg)
{
    sock_capture_t *c = (sock_capture_t *)arg;
    int fd = accept(c->lfd, NULL, NULL);
    if (fd < 0) return NULL;
    for (;;) {
        ssize_t r;
        if (c->len == c->cap) break;
        r = recv(fd, c->buf + c->len, c->cap - c->len, 0);
        if (r <= 0) break;
        c->len += (size_t)r;
    }
    close(fd);
    return NULL;
}
int64_t lref_sockdrv_stream(const float *rgb_display, int width, int height, const uint32_t *pixels_xy, uint64_t npixels,
                            unsigned char *out, uint64_t cap)
{
    struct sockaddr_in addr;
    sock_capture_t c;
    pthread_t th;
    uint64_t i;
    int one = 1;
    c.lfd = socket(AF_INET, SOCK_STREAM, 0);
    if (c.lfd < 0) return -1;
    setsockopt(c.lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET; addr.sin_port = htons(DEFAULT_PORT); addr.sin_addr.s_addr = inet_addr(LOCALADDR);
    if (bind(c.lfd, (struct sockaddr *)&addr, sizeof(addr)) != 0 || listen(c.lfd, 1) != 0) { close(c.lfd); return -1; }
    c.buf = out; c.cap = cap; c.len = 0;
    pthread_create(&th, NULL, sock_capture_main, &c);
    if (!sock_dd_open("capture", width, height, 32, "rgb", "float")) { close(c.lfd); pthread_cancel(th); pthread_join(th, NULL); return -1; }
    for (i = 0; i < npixels; i++) {                       /* pixels_xy: x | y << 16, y already the display row */
        const int x = (int)(pixels_xy[i] & 0xffffu), y = (int)(pixels_xy[i] >> 16);
        sock_dd_write(x, y, rgb_display + 3 * ((size_t)y * width + x));
    }
    sock_dd_close();
    pthread_join(th, NULL);
    close(c.lfd);
    /* the driver's packet counter is a static that outlives the frame (a lucille process renders one frame): bring it back to zero
     * so that the next capture starts like a fresh process.  The descriptor is closed, the one send() this triggers fails with EBADF. */
    { const float zero[3] = { 0.0f, 0.0f, 0.0f }; for (i = npixels % (MAXPACKETS); i % (MAXPACKETS) != 0; i++) sock_dd_write(0, 0, zero);   /* MAXPACKETS is an unparenthesised 32*32 */ }
    return (int64_t)c.len;
}

/* reference traversal counters (only meaningful in libluciref_stat.so; bvh.c:146,686-688) */
extern ri_bvh_stat_traversal_t g_stattrav;
void lref_stats_reset(void) { memset(&g_stattrav, 0, sizeof(g_stattrav)); }
void lref_stats_get(uint64_t *out6)
{
    out6[0] = g_stattrav.nrays;
    out6[1] = g_stattrav.ninner_node_traversals;
    out6[2] = g_stattrav.nleaf_node_traversals;
    out6[3] = g_stattrav.ntested_triangles;
    out6[4] = g_stattrav.nactually_hit_triangles;
    out6[5] = 0;
}
int lref_has_stats(void)
{
#ifdef RI_BVH_TRACE_STATISTICS
    return 1;
#else
    return 0;
#endif
}

/* beam visibility query (bvh.c:612-667): org[3], dirs[4][3]; returns RI_BEAM_* code */
int lref_beam_visibility(void *h, const double *org, const double *dirs)
{
    lref_scene_t *s = (lref_scene_t *)h;
    ri_beam_t beam;
    ri_vector_t o, d[4];
    int i, k;
    memset(&beam, 0, sizeof(beam));
    for (k = 0; k < 3; k++) o[k] = org[k];
    o[3] = 1.0;
    for (i = 0; i < 4; i++) { for (k = 0; k < 3; k++) d[i][k] = dirs[3 * i + k]; d[i][3] = 0.0; }
    if (ri_beam_set(&beam, o, d) != 0) return -1;            /* directions not in one octant (beam.c:356-377) */
    return ri_bvh_intersect_beam_visibility(s->bvh, &beam, NULL);
}

/* ------------------------------------------------------------------ frame level */

typedef struct {
    int      nthreads, width, height, pixelsamples, gather_nsamples, accel_method;
    float   *rgb;            /* [h][w][3], rows as the display driver receives them (y already flipped) */
    int      w, h;
    double   render_seconds;
    uint64_t nrays;
    /* scene capture (taken in the display-open callback, before the build) */
    double  *tri_xyz; uint64_t ntris; uint32_t *tri_geom; int ngeoms;
    double  *tri_nrm;        /* [ntris][3][3] vertex normals per corner (zeros where the geom has none) */
    double  *tri_st;         /* [ntris][3][2] texture coordinates per corner (zeros where the geom has none) */
    uint8_t *tri_has_st;     /* [ntris] */
    double   c2w[16]; double flength; int is_rh; int ortho; double fov;
    int      xsamples, ysamples, gather, bucket_size, bucket_order;
    int      has_normals;
} frame_t;

static frame_t g_frame;

static void world_begin_cb(void)
{
    ri_option_t  *opt  = ri_render_get()->context->option;
    ri_display_t *disp = ri_option_get_curr_display(opt);
    if (g_frame.pixelsamples > 0) {
        disp->sampling_rates[0] = g_frame.pixelsamples;
        disp->sampling_rates[1] = g_frame.pixelsamples;
    }
    if (g_frame.nthreads > 0) opt->nthreads = g_frame.nthreads;
    if (g_frame.gather_nsamples > 0) opt->gather_nsamples = g_frame.gather_nsamples;
    if (g_frame.accel_method >= 0) opt->accel_method = g_frame.accel_method;      /* what option.c:453-461 sets from the RIB */
    if (g_frame.width > 0 && g_frame.height > 0) {
        opt->camera->horizontal_resolution = g_frame.width;
        opt->camera->vertical_resolution   = g_frame.height;
    }
    disp->display_type   = strdup("callback");
    disp->display_format = strdup("float");
}

static int dd_open(const char *name, int width, int height, int bits, RtToken component, const char *format)
{
    ri_render_t *r = ri_render_get();
    ri_option_t *opt = r->context->option;
    ri_display_t *disp = ri_option_get_curr_display(opt);
    ri_camera_t *cam = opt->camera;
    ri_list_t *itr;
    uint64_t n = 0, idx = 0;
    int g = 0, i, j;

    g_frame.w = width; g_frame.h = height;
    g_frame.rgb = (float *)calloc((size_t)width * height * 3, sizeof(float));

    /* capture the triangle soup exactly as create_triangle_list() will flatten it */
    for (itr = ri_list_first(r->scene->geom_list); itr; itr = ri_list_next(itr)) {
        ri_geom_t *geom = (ri_geom_t *)itr->data;
        n += geom->nindices / 3; g++;
    }
    g_frame.ntris = n; g_frame.ngeoms = g;
    g_frame.tri_xyz  = (double *)malloc(sizeof(double) * 9 * (n ? n : 1));
    g_frame.tri_geom = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    g_frame.tri_nrm  = (double *)calloc(9 * (n ? n : 1), sizeof(double));
    g_frame.tri_st   = (double *)calloc(6 * (n ? n : 1), sizeof(double));
    g_frame.tri_has_st = (uint8_t *)calloc(n ? n : 1, 1);
    g = 0;
    for (itr = ri_list_first(r->scene->geom_list); itr; itr = ri_list_next(itr)) {
        ri_geom_t *geom = (ri_geom_t *)itr->data;
        unsigned int t;
        if (geom->normals) g_frame.has_normals = 1;
        attach_frame_texture(geom);
        for (t = 0; t < geom->nindices / 3; t++) {
            for (i = 0; i < 3; i++)
                for (j = 0; j < 3; j++)
                {
                    g_frame.tri_xyz[9 * idx + 3 * i + j] = geom->positions[geom->indices[3 * t + i]][j];
                    if (geom->normals) g_frame.tri_nrm[9 * idx + 3 * i + j] = geom->normals[geom->indices[3 * t + i]][j];
                    if (j < 2 && geom->texcoords) g_frame.tri_st[6 * idx + 2 * i + j] = geom->texcoords[2 * geom->indices[3 * t + i] + j];
                    else if (j < 2 && geom->texcoords_unshared) g_frame.tri_st[6 * idx + 2 * i + j] = geom->texcoords_unshared[2 * (3 * t + i) + j];
                }
            g_frame.tri_geom[idx] = (uint32_t)g;
            g_frame.tri_has_st[idx] = (geom->texcoords || geom->texcoords_unshared) ? 1 : 0;
            idx++;
        }
        g++;
    }

    /* camera: ri_camera_setup() is what ri_render_frame() calls next (render.c:337); it is idempotent */
    ri_camera_setup(cam);
    for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) g_frame.c2w[4 * i + j] = cam->camera_to_world.f[i][j];
    g_frame.flength = cam->flength;
    g_frame.is_rh   = cam->is_rh;
    g_frame.ortho   = (cam->camera_projection == RI_ORTHOGRAPHIC);
    g_frame.fov     = cam->fov;
    g_frame.xsamples = (int)disp->sampling_rates[0];
    g_frame.ysamples = (int)disp->sampling_rates[1];
    g_frame.gather   = opt->gather_nsamples;
    g_frame.bucket_size  = r->bucket_size;
    g_frame.bucket_order = r->bucket_order;
    capture_sunsky();
    return 1;
}

static int dd_write(int x, int y, const void *pixel)
{
    const float *p = (const float *)pixel;
    float *dst = g_frame.rgb + 3 * ((size_t)y * g_frame.w + x);
    dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
    return 1;
}

static int dd_close(void)
{
    /* ri_raytrace_statistics() has just printed; the counters are still live here */
    g_frame.nrays = ri_render_get()->stat.nrays;
    g_frame.render_seconds = ri_timer_elapsed(ri_render_get()->context->timer, "Render frame");
    return 1;
}

/* One frame per process (the reference keeps one-shot statics, e.g. spiral.c:17 g_n). */
static int g_accel_method = -1;
void lref_set_accel_method(int m) { g_accel_method = m; }

int lref_render_rib(const char *path, int nthreads, int width, int height, int pixelsamples, int gather_nsamples)
{
    char buf[2048];
    const char *slash;
    int argc = 1; char *argv0 = "lref"; char **argv = &argv0;

    memset(&g_frame, 0, sizeof(g_frame));
    g_frame.nthreads = nthreads; g_frame.width = width; g_frame.height = height;
    g_frame.pixelsamples = pixelsamples; g_frame.gather_nsamples = gather_nsamples;
    g_frame.accel_method = g_accel_method;

    ri_parallel_init(&argc, &argv);                     /* lsh/main.c:119 */
    RiBegin(RI_NULL);                                   /* lsh/main.c:153 */
    ri_dd_callback(dd_open, dd_close, dd_write);        /* ri/display.c:118-136 */
    ri_backdoor_world_begin_cb(world_begin_cb);         /* lsh/main.c:162 */
    ri_timer_start(ri_render_get()->context->timer, "RIB parsing");   /* lsh/main.c:165 */

    slash = strrchr(path, '/');                         /* set_ribpath, lsh/main.c:78-101 */
    if (slash) {
        size_t len = (size_t)(slash - path) + 1;
        memcpy(buf, path, len); buf[len] = 0;
        strcpy(ri_render_get()->ribpath, buf);
        ri_option_add_searchpath(ri_render_get()->context->option, buf);
    }
    if (getcwd(buf, sizeof(buf) - 1)) ri_option_add_searchpath(ri_render_get()->context->option, buf);

    if (lref_rib_parse_file(path) != 0) return -1;
    RiEnd();
    return 0;
}

int      lref_frame_width(void)  { return g_frame.w; }
int      lref_frame_height(void) { return g_frame.h; }
float   *lref_frame_rgb(void)    { return g_frame.rgb; }
double   lref_frame_seconds(void){ return g_frame.render_seconds; }
uint64_t lref_frame_nrays(void)  { return g_frame.nrays; }
uint64_t lref_frame_ntris(void)  { return g_frame.ntris; }
double  *lref_frame_tris(void)   { return g_frame.tri_xyz; }
uint32_t*lref_frame_trigeom(void){ return g_frame.tri_geom; }
double  *lref_frame_normals(void){ return g_frame.tri_nrm; }
double  *lref_frame_st(void)     { return g_frame.tri_st; }
uint8_t *lref_frame_has_st(void) { return g_frame.tri_has_st; }
/* ---- sun-sky (row a12) ---------------------------------------------------------------------------------------------
 * block layout = orc_sunsky_t / ri_b200_sunsky_t minus the tables the reference keeps private:
 * out[0..1] sun_theta, sun_phi; [2..6] perez_x; [7..11] perez_y; [12..16] perez_Y; [17..19] zenith x,y,Y;
 * [20] nsun; then nsun x (direction[3], col[3]).  Returns 0 when the scene has no sun-sky light. */
static double g_sunsky[21 + 6 * 4];
static int    g_has_sunsky = 0;
static void capture_sunsky(void)
{
    ri_scene_t *scene = ri_render_get()->scene;
    ri_list_t *itr;
    int i, n = 0;
    g_has_sunsky = 0;
    if (!scene->sunsky_light || !scene->sunsky_light->sunsky) return;
    {
        const ri_sunsky_t *s = scene->sunsky_light->sunsky;
        g_sunsky[0] = s->sun_theta; g_sunsky[1] = s->sun_phi;
        for (i = 0; i < 5; i++) { g_sunsky[2 + i] = s->perez_x[i]; g_sunsky[7 + i] = s->perez_y[i]; g_sunsky[12 + i] = s->perez_Y[i]; }
        g_sunsky[17] = s->zenith_x; g_sunsky[18] = s->zenith_y; g_sunsky[19] = s->zenith_Y;
    }
    for (itr = ri_list_first(scene->light_list); itr; itr = ri_list_next(itr)) {
        const ri_light_t *l = (const ri_light_t *)itr->data;
        if (l->type != LIGHTTYPE_SUNLIGHT || n >= 4) continue;
        for (i = 0; i < 3; i++) { g_sunsky[21 + 6 * n + i] = l->direction[i]; g_sunsky[21 + 6 * n + 3 + i] = l->col[i]; }
        n++;
    }
    g_sunsky[20] = n;
    g_has_sunsky = 1;
}
int lref_frame_sunsky(double *out45)
{
    if (g_has_sunsky) memcpy(out45, g_sunsky, sizeof(g_sunsky));
    return g_has_sunsky;
}
/* ri_sunsky_init() + ri_sunsky_get_sky_rgb() without a scene: the sky colour of n directions, and the block above (no lights) */
void lref_sunsky_eval(float latitude, float longitude, float sm, int jd, float tod, float turb,
                      const float *dirs, uint64_t n, float *rgb, double *out21)
{
    ri_sunsky_t *s = ri_sunsky_new();
    uint64_t k; int i;
    ri_sunsky_init(s, latitude, longitude, sm, jd, tod, turb, 0);
    for (k = 0; k < n; k++) ri_sunsky_get_sky_rgb(rgb + 3 * k, s, dirs + 3 * k);
    out21[0] = s->sun_theta; out21[1] = s->sun_phi;
    for (i = 0; i < 5; i++) { out21[2 + i] = s->perez_x[i]; out21[7 + i] = s->perez_y[i]; out21[12 + i] = s->perez_Y[i]; }
    out21[17] = s->zenith_x; out21[18] = s->zenith_y; out21[19] = s->zenith_Y; out21[20] = 0;
}

/* ---- .hdr display driver (display/hdrdrv.c) driven exactly like bucket_write() drives it: open, one write per pixel, close */
extern int hdr_dd_open(const char *name, int width, int height, int bits, RtToken component, const char *format);
extern int hdr_dd_write(int x, int y, const void *pixel);
extern int hdr_dd_close(void);
int lref_hdr_file(const float *rgb, int width, int height, const char *path)
{
    int x, y;
    if (!hdr_dd_open(path, width, height, 32, RI_RGB, "float")) return -1;
    for (y = 0; y < height; y++)
        for (x = 0; x < width; x++) hdr_dd_write(x, y, rgb + 3 * ((size_t)y * width + x));
    hdr_dd_close();
    return 0;
}

/* out: c2w[16], flength, is_rh, ortho, fov, xsamples, ysamples, gather, bucket_size, bucket_order, has_normals, ngeoms */
void lref_frame_camera(double *out27)
{
    int i;
    for (i = 0; i < 16; i++) out27[i] = g_frame.c2w[i];
    out27[16] = g_frame.flength; out27[17] = g_frame.is_rh; out27[18] = g_frame.ortho; out27[19] = g_frame.fov;
    out27[20] = g_frame.xsamples; out27[21] = g_frame.ysamples; out27[22] = g_frame.gather;
    out27[23] = g_frame.bucket_size; out27[24] = g_frame.bucket_order; out27[25] = g_frame.has_normals;
    out27[26] = g_frame.ngeoms;
}
