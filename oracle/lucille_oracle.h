/*
 * lucille_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of lucille's ray-intersection hot path (BVH build, ray traversal,
 * leaf Moeller-Trumbore, hit-state build, ambient-occlusion sample loop, pixel loop).
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (lucille_b200/) never does.
 *
 * PARITY PINNING: the double-precision instantiation is pinned bit-for-bit against the
 * compiled, unmodified reference (oracle/_ref/libluciref.so, built by oracle/build_ref.sh)
 * -- identical trees, identical (hit,t,u,v,triangle) per ray, identical AO frames -- by
 * tests/test_oracle_vs_reference.py and, where /root/reference is absent, against the
 * golden vectors those runs produced (tests/golden/, generator: tests/golden/make_golden.py).
 * The reference's own tests hold no numbers for this path.
 *
 * The single-precision instantiation is the same code with REAL=float (no FMA contraction:
 * compile with -ffp-contract=off).  It is the bit-exact target for the fp32 CUDA kernels.
 */
#ifndef LUCILLE_ORACLE_H
#define LUCILLE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INFINITY   1.0e38      /* RI_INFINITY, include/ri.h:47 */
#define ORC_EPS        1.0e-14     /* RI_EPS, src/base/common.h:27 */
#define ORC_LEAF_TRIS  16          /* BVH_NTRIS_LEAF, src/render/bvh.c:81 */
#define ORC_BINS       64          /* BVH_BIN_SIZE,  src/render/bvh.c:82 */
#define ORC_MISS_PRIM  0xffffffffu

/* canonical interchange form of a tree: DFS preorder, same layout as ref_shim.c's lref_node_t */
typedef struct {
    int32_t is_leaf;
    int32_t axis;
    int64_t child0, child1;
    int64_t tri_start, ntris;
    double  lbox[6];             /* slot 0: min xyz, max xyz */
    double  rbox[6];             /* slot 1 */
} orc_node_t;

typedef struct { double t, u, v; uint32_t prim; uint32_t hit; } orc_hit_f64;   /* prim = position in post-build order */
typedef struct { float  t, u, v; uint32_t prim; } orc_hit_f32;                 /* miss: t=1e38, prim=ORC_MISS_PRIM */

/* hit-state subset the AO transport consumes (intersection_state.h:34-61) */
typedef struct { double P[3], Ng[3], Ns[3], tangent[3], binormal[3]; } orc_state_f64;

typedef struct {
    uint64_t nrays, ninner, nleaf, ntris, nhit_tris;   /* g_stattrav fields, bvh.h:119-130 */
} orc_counters_t;

typedef struct {
    double c2w[16];              /* camera_to_world, row-vector convention (vector.h:182-) */
    double flength;              /* 1/tan(fov*3.141592/360), camera.c:219 */
    int32_t is_rh;               /* Orientation "rh" -> sign -1, camera.c:269 */
    int32_t width, height;
    int32_t xsamples, ysamples;  /* PixelSamples */
    int32_t ntheta, nphi;        /* (int)sqrt(gather_nsamples) each, ambientocclusion.c:378-387 */
    int32_t bucket_size;         /* 32, render.c:197 */
} orc_frame_t;

typedef struct orc_tree orc_tree;

orc_tree *orc_build(const double *tri_xyz, uint64_t ntris);           /* bvh.c:276-379 */
void      orc_free(orc_tree *t);
int       orc_is_empty(const orc_tree *t);
int64_t   orc_num_nodes(const orc_tree *t);
int64_t   orc_get_nodes(const orc_tree *t, orc_node_t *out);          /* DFS preorder */
void      orc_get_triorder(const orc_tree *t, uint32_t *orig);        /* post-build position -> input triangle */
void      orc_scene_bbox(const orc_tree *t, double *bmin, double *bmax);
int       orc_max_depth(const orc_tree *t);

/* optional per-corner vertex normals, [ntris][3][3] in INPUT triangle order (ri_geom_t.normals through the index list).
 * With normals the shading normal is Ns = (1-u-v) n0 + u n1 + v n2 (ri_lerp_vector, base/geometric.c:40-62), NOT normalised;
 * Ng stays the geometric normal (intersection_state.c:152-180).  A triangle with nine zero components has no normals
 * (its geom's `normals` is NULL) and keeps Ns = Ng.  NULL removes them all. */
void orc_set_normals(orc_tree *t, const double *tri_normals);

/* vertex colours / texture coordinates / back-side flag of the hit state (intersection_state.c:192-246), per INPUT triangle:
 * colors [ntris][3][3] with has_color[ntris] (0: the triangle's geom has no Cs -> colour (1,1,1), :204-207), st [ntris][3][2] with
 * has_st[ntris] (0: no texture coordinates -> st = 0, :228-231), inside[ntris] (1: second half of a two-sided geom, :233-244).
 * Any pointer may be NULL (= absent everywhere). */
void orc_set_attributes(orc_tree *t, const double *colors, const uint8_t *has_color, const double *st, const uint8_t *has_st,
                        const uint8_t *inside);
typedef struct { double E[3], I[3], color[3], st[2], t; int32_t inside, hit; } orc_state_ext_f64;
/* the rest of ri_intersection_state_build (intersection_state.c:123-133, 192-246): E, I = normalize(dir), colour, st, inside */
void orc_state_ext_build_f64(const orc_tree *t, const double *rays, const orc_hit_f64 *hits, uint64_t n, orc_state_ext_f64 *out);

/* closest hit, reference order (bvh.c:430-542, 1092-1188). rays f64: [n][6] org,dir. f32: [n][8] ox,oy,oz,tmin,dx,dy,dz,tmax
 * (tmin/tmax are carried but never read, exactly like ri_ray_t.min_t/max_t). counters may be NULL. */
void orc_intersect_f64(const orc_tree *t, const double *rays, uint64_t n, orc_hit_f64 *out, orc_counters_t *c);
void orc_intersect_f32(const orc_tree *t, const float  *rays, uint64_t n, orc_hit_f32 *out, orc_counters_t *c);
/* occlusion boolean = (closest-hit found a hit); traversal stops at the first leaf that records a hit */
void orc_occluded_f64(const orc_tree *t, const double *rays, uint64_t n, uint8_t *out, orc_counters_t *c);
void orc_occluded_f32(const orc_tree *t, const float  *rays, uint64_t n, uint8_t *out, orc_counters_t *c);

/* post-hit state (intersection_state.c:99-248, geometry without normals/colours/st) */
void orc_state_build_f64(const orc_tree *t, const double *rays, const orc_hit_f64 *hits, uint64_t n, orc_state_f64 *out);

/* MT19937 stream of randomMT2() (random.c:98-112, 211-247): n doubles in [0,1) from seed 4357 */
void orc_mt_stream(uint32_t seed, uint64_t n, double *out);
void orc_mt_stream_u32(uint32_t seed, uint64_t n, uint32_t *out);

/* spiral bucket order (spiral.c:68-130, render.c:582-710): writes nb (x,y,w,h) int32 quadruples, returns nb */
int orc_bucket_list(int width, int height, int bucket_size, int32_t *out_xywh, int max_buckets);

/* Hammersley sub-pixel jitter (render.c:830-917) */
void orc_subpixel_jitter(int xs, int ys, int xsamples, int ysamples, double *jx, double *jy);

/* camera ray (camera.c:248-352 perspective branch + render.c:770-781 normalise) */
void orc_camera_ray(const orc_frame_t *f, double x, double y, double *org3, double *dir3);

/* full ambient-occlusion frame, single MT stream, spiral bucket order = the reference at --nthreads 1
 * (render.c:715-823,1107-1146; ambientocclusion.c:42-151,332-415). rgb: [h][w][3] float, row (H-1-y). */
void orc_render_ao(const orc_tree *t, const orc_frame_t *f, float *rgb, uint64_t *nrays_out);

/* Beam (4-corner frustum) visibility query, SURVEY 8a row a10: ri_beam_set (beam.c:332-466) + ri_bvh_intersect_beam_visibility
 * (bvh.c:612-667) -> test_beam_aabb (n-vertex against the 4 frustum planes, bvh.c:1997-2089), bvh_traverse_beam_visibility
 * (2648-2746, near child = child[dirsign[dominant_axis]]), leaf: test_beam_triangle on the 4 corner rays (2139-2281), first
 * triangle that is not a complete miss decides.  beams: [n][15] doubles = org.xyz, dir0.xyz .. dir3.xyz.
 * out[i]: ORC_BEAM_MISS_COMPLETELY / HIT_COMPLETELY / HIT_PARTIALLY (beam.h:27-29), or ORC_BEAM_INVALID (-1) when the four
 * directions do not share a sign on every axis (ri_beam_set returns -1, beam.c:356-377). */
#define ORC_BEAM_MISS_COMPLETELY 0
#define ORC_BEAM_HIT_COMPLETELY  1
#define ORC_BEAM_HIT_PARTIALLY   2
#define ORC_BEAM_INVALID        (-1)
void orc_beam_visibility(const orc_tree *t, const double *beams, uint64_t n, int32_t *out);

/* Path-trace transport.  The reference's pathtrace.c is not in its build and does not
 * compile against its current headers, so there is no reference binary: this is a restatement of the SKETCH's
 * control flow with builder-stated inputs -- Lambert kd (grey), constant environment Le, counter-based RNG, and a
 * deterministic sin/cos (plain IEEE mul/add polynomial) so that CPU and GPU paths are bit-identical:
 *   trace_pixel (pathtrace.c:189-244): 2 draws jitter the pixel; eye ray; miss -> Le; else trace_path, then one more sampled
 *     direction from the last vertex, G *= brdf, visibility ray, radiance = (hit ? 0 : Le) * G
 *   trace_path (246-314): stop at MAX_PATH_VERTICES (depth starts at 2); russian roulette on kd (1 draw, 386-405); lobe pick
 *     (1 draw, always 'D' here, 407-433); cosine sampling about the UN-flipped Ng with float-rounded local vector (2 draws,
 *     480-508); next ray starts AT P, no offset (289-290); on hit G *= kd/pi (brdf, 510-537), recurse
 * PARITY UNPINNED against the reference (nothing to run); pinned only between this restatement and the CUDA kernel. */
typedef struct {
    double  c2w[16];
    double  flength;
    int32_t is_rh;
    int32_t width, height;
    int32_t spp;                 /* pt_nsamples, option.c:142 (default 4; C4 uses 256) */
    int32_t max_vertices;        /* MAX_PATH_VERTICES 10, pathtrace.c:66 */
    uint32_t seed;
    double  kd, Le;
    int32_t rank, world;         /* product only (tile sharding); the oracle renders everything */
    int32_t bucket_size;
} orc_path_frame_t;
void orc_render_pathtrace(const orc_tree *t, const orc_path_frame_t *f, float *rgb, uint64_t *nrays_out);
/* trace() shadeop up to the shader call (render/shader.c:895-976) and the light samples of next_lightsource() (shader.c:1116-1310) */
void orc_shade_trace(const orc_tree *T, const double *pr, uint64_t n, const float *env, int ew, int eh, int use_env,
                     double *rays_out, orc_hit_f64 *hits, orc_state_f64 *states, orc_state_ext_f64 *exts, double *eye3, double *miss_rgb3);
int orc_light_samples(const orc_tree *T, int nsamples, double angle, uint32_t seed, const double *points, uint64_t n,
                      const float *env, int ew, int eh, double *L_out, double *Cl_out, uint8_t *visible, uint64_t *nrays_out);
void orc_det_sincos2pi(double r, double *s, double *c);
/* ray batch of the point-based AO call (calculate_occlusion's ray set-up, ambientocclusion.c:56-117): [n*ntheta*nphi][8] floats */
void orc_ao_point_rays_f64(const double *points, uint64_t n, uint64_t first_point, int ntheta, int nphi, uint64_t seed, double eps, double *rays_out);
void orc_ao_point_rays_f32(const double *points, uint64_t n, uint64_t first_point, int ntheta, int nphi, uint64_t seed, double eps, float *rays_out);

/* ---- material texture of the AO transport (ambientocclusion.c:393-401): radiance *= ri_texture_fetch(texture, st) per channel.
 * ri_texture_fetch (render/texture.c:86-236, USE_ZORDER 0): wrap by floor, clamp, bilinear over the 2x2 texels at
 * (u (w-1), v (h-1)) with ZERO texels beyond the last row / column.  rgba: [h][w][4] floats (ri_texture_t.data). */
void orc_texture_fetch(const float *rgba, int width, int height, const double *uv, uint64_t n, double *out4);
/* orc_render_ao with every geom carrying this material texture; st per triangle from orc_set_attributes (0 where absent) */
void orc_render_ao_textured(const orc_tree *t, const orc_frame_t *f, const float *rgba, int tex_width, int tex_height,
                            float *rgb, uint64_t *nrays_out);

/* ---- dirt-map transport (transport/dirtmap.c:84-221 calculate_dirt, 223-293 ri_transport_dirtmap): the AO
 * hemisphere loop with a fixed 4 x 4 pattern, origin offset 1e-5 and CLOSEST hits: a gather ray adds black when it hits within 0.1,
 * white beyond 0.5 or on a miss, and (1-p) white - p black with p = clamp(pow(1 - (t-0.1)/0.4, 1)) in between (mix_color, :70-82);
 * Lo = sum / 16 on three channels.  orc_transport_batch: one radiance per eye ray ([n][6] org,dir), the thread's MT19937 stream
 * re-seeded with 4357 and consumed ray after ray (which: 0 ambient occlusion with ntheta x nphi, 1 dirt map).
 * orc_render_dirtmap: the pixel loop of orc_render_ao around the same transport. */
void orc_transport_batch(const orc_tree *t, int which, int ntheta, int nphi, const double *rays, uint64_t n, double *radiance3);
void orc_render_dirtmap(const orc_tree *t, const orc_frame_t *f, float *rgb, uint64_t *nrays_out);

/* ---- Whitted transport (transport/whitted.c:31-83 trace_whitted, 92-151 ri_transport_whitted): from the eye hit
 * a chain of REFRACTED rays (ri_refract with eta 1.33, render/reflection.c:69-127; total internal reflection falls back to ri_reflect,
 * :25-49, whose dot product is a float), origin P + 1e-7 Rd, at most 8 bounces; the radiance is the environment map looked up along
 * the direction that leaves the scene (ri_texture_ibl_fetch, render/texture.c:238-277, angular map) and zero when the chain is cut.
 * env: [h][w][4] floats or NULL (no environment: radiance 0).  One radiance per eye ray; nrays counts ri_raytrace calls. */
void orc_transport_whitted(const orc_tree *t, const float *env, int env_w, int env_h, const double *rays, uint64_t n, double *radiance3,
                           uint64_t *nrays_out);
/* ri_transport_sample / trace_path (transport/transport.c:50-173): white where the eye ray hits (no light geometry), black elsewhere */
void orc_render_hitmask(const orc_tree *t, const orc_frame_t *f, float *rgb, uint64_t *nrays_out);
void orc_render_whitted(const orc_tree *t, const orc_frame_t *f, const float *env, int env_w, int env_h, float *rgb, uint64_t *nrays_out);

/* ---- sun-sky gather (row a12): ambientocclusion.c:153-324 gather_sunsky + contribution_from_sunlight, with the sky lookup
 * ri_sunsky_get_sky_rgb (render/sunsky.c:24-38 angle_between, 136-152 PerezFunction, 297-408; render/specrend.c:127-172
 * xyz_to_rgb, 366-440 spectrum_to_xyz).  The block is what the host side owns after ri_sunsky_init(): Perez coefficients,
 * zenith values and sun angles of ri_sunsky_t (sunsky.h:22-50), the S0/S1/S2 daylight tables of sunsky.dat, the CIE matching
 * table of specrend.c:387-414, the chromaticities of specrend.h CIEsystem, and the LIGHTTYPE_SUNLIGHT lights of the scene
 * (lightsource.c:152-170).  Same layout as ri_b200_sunsky_t (include/lucille_b200.h). */
typedef struct {
    float   sun_theta, sun_phi;
    float   perez_x[5], perez_y[5], perez_Y[5];
    float   zenith_x, zenith_y, zenith_Y;
    float   S0[41], S1[41], S2[41];
    float   cie[81][3];
    float   cs[8];               /* xRed yRed xGreen yGreen xBlue yBlue xWhite yWhite */
    int32_t nsun;                /* <= 4 */
    int32_t pad;
    double  sun_dir[4][3];       /* ri_light_t.direction */
    double  sun_col[4][3];       /* ri_light_t.col */
} orc_sunsky_t;
/* sky colour for n directions (world space, the transport's ray.dir cast to float): rgb [n][3] */
void orc_sunsky_sky_rgb(const orc_sunsky_t *s, const float *dirs, uint64_t n, float *rgb);
/* whole frame with the sun-sky transport: 8x8 gather, eps 1e-5, Lo = (1/pi) * col / 64 (f->ntheta/nphi are ignored) */
void orc_render_sunsky(const orc_tree *t, const orc_frame_t *f, const orc_sunsky_t *s, float *rgb, uint64_t *nrays_out);

/* ---- the output step right after the path: the Radiance .hdr display driver.  hdr_dd_write clamps negative
 * components to 0 and adds onto a zeroed float buffer (display/hdrdrv.c:62-88); hdr_dd_close writes RGBE_WriteHeader +
 * RGBE_WritePixels_RLE (hdrdrv.c:90-102; imageio/rgbe.c:78-96 float2rgbe, 117-139 header, 244-294 run-length coder, 296-340
 * scanline framing, flat pixels when the width is < 8 or > 0x7fff).  rgb: [h][w][3] floats in display order.  Returns the file
 * size (the bytes are written when cap is large enough; call with out == NULL to size). */
uint64_t orc_hdr_encode(const float *rgb, int width, int height, uint8_t *out, uint64_t cap);

/* counter-based uniform for the synthetic configs (shared definition with the product) */
uint64_t orc_splitmix64(uint64_t x);

#ifdef __cplusplus
}
#endif
/* hemisphere gathers at shading points: shader.c occlusion() (kind 0), ri_ibl_sample_cosweight (1), ri_domelight_sample (2) */
void orc_point_gather(const orc_tree *T, int kind, int nsamples, uint32_t seed, const double *points, uint64_t n,
                      const float *env, int ew, int eh, const double *col3, double intensity, double *out3, uint64_t *nrays_out);

/* the quasi-Monte Carlo branches of the IBL (kind 1) and dome-light (kind 2) gathers (Option "use_qmc") */
void orc_point_gather_qmc(const orc_tree *T, int kind, int nsamples, const double *points, uint64_t n, const int32_t *instance, int dim,
                          const float *env, int ew, int eh, const double *col3, double intensity, double *out3, uint64_t *nrays_out);

/* byte stream of the socket display driver (display/sockdrv.c) for a finished frame; returns the size (writes when cap suffices) */
uint64_t orc_sockdrv_encode(const float *rgb, int width, int height, int bucket_size, unsigned char *out, uint64_t cap);

#endif
