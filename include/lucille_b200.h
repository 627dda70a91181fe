/*
 * lucille_b200.h -- C ABI of the B200-native accelerator for lucille's ray-intersection hot path.
 *
 * This is the drop-in boundary: a shared library (lucille_b200/libb200accel.so) with plain-C entry
 * points, no C++/torch types.  Each entry point names the reference interface it replaces
 * (paths relative to the lucille source tree, commit ff81b332).  INTEGRATION.md shows the
 * ~40-line binding a lucille maintainer adds in src/render/accel.c to select it with
 *     Option "raytrace" "accel_method" ["b200"]
 *
 * Threading: build/free are called from one thread (render.c:335,1233).  Batch calls serialise
 * internally on the accelerator's CUDA stream; ri_b200_intersect1() is safe to call from the
 * reference's <=16 worker threads (render.c:1189-1198) -- it takes a mutex.
 *
 * Errors: every int-returning call returns 0 on success, <0 on failure; ri_b200_last_error()
 * returns a static description (the reference has no error channel beyond ri_accel_bind()'s -1;
 * the binding logs it with ri_log(LOG_FATAL, ...) and aborts, mirroring reference style).
 */
#ifndef LUCILLE_B200_H
#define LUCILLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RI_ACCEL_B200          2             /* next to RI_ACCEL_UGRID 0 / RI_ACCEL_BVH 1, accel.h:20-21 */
#define RI_B200_MISS_PRIM      0xffffffffu
#define RI_B200_INFINITY       1.0e38        /* RI_INFINITY, include/ri.h:47 */

#define RI_B200_PREC_F32       1u            /* fp32 node/triangle records: the throughput path */
#define RI_B200_PREC_F64       2u            /* double records: bit-identical to the CPU reference */
/* With BOTH record sets resident, batched double queries -- occlusion (ri_b200_occluded_*_f64, the AO / sun-sky / point gathers of
 * double frames) and closest hit (ri_b200_intersect_*_f64, the Whitted / dirt-map / trace() batches) -- run through the fp32 records
 * with a certified error bound on every box and triangle decision and consult the double records only for the decisions fp32 cannot
 * settle (csrc/hybrid.cuh): the double reference's answer for every ray, bit for bit, at 1.5-2x the rate of the double kernels.
 * Scenes whose vertices are not fp32 numbers, or that lie far from the world origin -- and accelerators built with double records
 * only -- get a set of fp32 records of their own for this at build time (coordinates relative to the scene's centre, derived on the
 * device from the double records; + 48 B per triangle slot, 64 B per node).
 * B200_HYBRID=0 in the environment turns the path off (the double kernels run instead); B200_HYBRID_CLOSEST=0 its closest-hit form alone. */
#define RI_B200_HOST_ONLY      0x100u        /* build + flatten on the host, no device upload: every trace call on
                                                such an accelerator fails loudly (used by the CPU-only tests) */
#define RI_B200_BUILD_DEVICE   0x200u        /* build the tree on the device (level-by-level binned SAH, csrc/bvh_build_gpu.cuh): the same
                                                 tree as the host builder and the reference, bit for bit */
#define RI_B200_BUILD_HOST     0x400u        /* build it with the threaded host builder.  Neither flag: device from 32 Ki triangles up */

typedef struct ri_b200_accel ri_b200_accel_t;   /* opaque; stored in ri_accel_t.data (accel.h:73) */

/* hit records.  prim = position in the post-build triangle order (the reference's ri_triangle_t[]
 * after bvh_construct, bvh.c:1897-1917); ri_b200_triorder() maps it back to the input triangle,
 * from which the binding recovers (geom*, index) exactly as bvh.c:858-859 does. */
typedef struct { float  t, u, v; uint32_t prim; } ri_b200_hit_f32;                 /* miss: t=1e38, prim=MISS */
typedef struct { double t, u, v; uint32_t prim; uint32_t hit; } ri_b200_hit_f64;

/* subset of ri_intersection_state_t (intersection_state.h:34-61) produced on device */
typedef struct { double P[3], Ng[3], Ns[3], tangent[3], binormal[3]; } ri_b200_state_f64;

/* canonical (DFS preorder) view of the tree, for parity checks against the reference-built tree */
typedef struct {
    int32_t is_leaf;
    int32_t axis;                 /* ri_qbvh_node_t.axis0, bvh.h:91 */
    int64_t child0, child1;       /* inner: indices in this dump */
    int64_t tri_start, ntris;     /* leaf: slice of the post-build triangle order */
    double  lbox[6];              /* slot 0 (BMIN_X0.., bvh.h:42-65): min xyz, max xyz */
    double  rbox[6];              /* slot 1 */
} ri_b200_node_t;

typedef struct {
    uint64_t ntris;
    int64_t  ninner, nleaf;
    int32_t  max_depth;
    int32_t  empty;               /* ri_bvh_t.empty, bvh.h:172 */
    uint32_t precisions;          /* RI_B200_PREC_* resident on the device */
    int32_t  device;
    double   bmin[3], bmax[3];    /* ri_bvh_t.bmin/bmax incl. margin, bvh.c:328-338 */
    double   build_seconds;       /* host build, the reference's "BVH Construction" timer */
    double   upload_seconds;
    uint64_t device_bytes;
} ri_b200_info_t;

/* traversal counters in the reference's own units (ri_bvh_stat_traversal_t, bvh.h:119-130) */
typedef struct { uint64_t nrays, ninner, nleaf, ntris, nhit_tris; } ri_b200_counters_t;

/* frame description for the on-device ambient-occlusion transport */
typedef struct {
    double  c2w[16];              /* ri_camera_t.camera_to_world, row-vector convention (vector.h:182-210) */
    double  flength;              /* ri_camera_t.flength, camera.c:219 */
    int32_t is_rh;                /* camera.c:269 */
    int32_t width, height;
    int32_t xsamples, ysamples;   /* PixelSamples (ri_display_t.sampling_rates) */
    int32_t ntheta, nphi;         /* (int)sqrt(gather_nsamples), ambientocclusion.c:378-387 */
    int32_t bucket_size;          /* ri_render_t.bucket_size (32, render.c:197) */
    int32_t rng_mode;             /* 0: MT19937 stream, seed 4357, reference consumption order (random.c:211-247)
                                     1: counter-based SplitMix64 keyed by (seed, sample, j, i, k) */
    uint32_t seed;
    int32_t rank, world;          /* buckets b with b % world == rank are rendered; others left zero */
    int32_t precision;            /* RI_B200_PREC_F64 (reference-exact) or RI_B200_PREC_F32 */
} ri_b200_frame_t;

typedef struct {
    uint64_t nrays_primary, nrays_ao, nhits_primary;
    double   ms_total, ms_primary, ms_rng, ms_ao, ms_resolve;   /* CUDA-event times */
} ri_b200_frame_stats_t;

const char *ri_b200_last_error(void);
int         ri_b200_device_count(void);

/* ---- replaces accel_build_func / ri_bvh_build (accel.h:24-25, bvh.c:276-379) ------------------
 * tri_xyz: [ntris][3][3] doubles in the order create_triangle_list() flattens the scene's geoms
 * (bvh.c:1736-1826).  Builds the reference's binned-SAH tree on the host (identical topology,
 * boxes and triangle order), uploads the requested precisions to `device`.  ntris==0 gives a valid
 * empty accelerator (bvh.c:311-315).  Returns NULL on failure. */
ri_b200_accel_t *ri_b200_build(const double *tri_xyz, uint64_t ntris, uint32_t precisions, int device);

/* optional per-corner vertex normals, [ntris][3][3] doubles in the same (input) triangle order as tri_xyz: what
 * ri_geom_t.normals holds behind the index list (polygon.c:677-733).  With them the shading normal of a hit is
 * Ns = (1-u-v) n0 + u n1 + v n2 (ri_lerp_vector, base/geometric.c:40-62; intersection_state.c:152-180), not normalised, and
 * the AO transport samples about Ns; Ng stays geometric.  A triangle whose nine components are all zero has no normals
 * (its geom's `normals` is NULL).  NULL removes them. */
int ri_b200_set_normals(ri_b200_accel_t *accel, const double *tri_normals);

/* ---- replaces accel_free_func / ri_bvh_free (accel.h:27-28, bvh.c:381-387); frees host and device memory */
void ri_b200_free(ri_b200_accel_t *accel);

int     ri_b200_info(const ri_b200_accel_t *accel, ri_b200_info_t *out);
int64_t ri_b200_export_nodes(const ri_b200_accel_t *accel, ri_b200_node_t *out, int64_t capacity);
int     ri_b200_triorder(const ri_b200_accel_t *accel, uint32_t *orig_out);   /* [ntris] */
/* flat device records of a RI_B200_HOST_ONLY accelerator (layout: lucille_b200/csrc/bvh_build.h); any pointer may be
 * NULL.  header_out[4] = root_word, ninner, top_count, number of triangle slots.  Returns ninner. */
int64_t ri_b200_export_flat(const ri_b200_accel_t *accel, void *nodes32, void *nodes64, void *tris32, void *tris64,
                            uint32_t *header_out);

/* the leaf-transposed copies of the triangle slots (same sizes as tris32 / tris64; layout: lucille_b200/csrc/bvh_build.h) */
int     ri_b200_export_flat_transposed(const ri_b200_accel_t *accel, void *tris32t, void *tris64t);

/* ---- replaces accel_intersect_func / ri_bvh_intersect for ONE ray (accel.h:30-34, bvh.c:430-542),
 * double precision, returns 1 on hit / 0 on miss / <0 on error.  state may be NULL. */
int ri_b200_intersect1(ri_b200_accel_t *accel, const double org[3], const double dir[3],
                       ri_b200_hit_f64 *hit, ri_b200_state_f64 *state);

/* ---- batched entry points (additions; what the batching hook in render.c:803/1133-1146 calls) --
 * HOST buffers; host<->device copies are part of the call and pipelined against the kernels.
 * f32 rays: [n][8] = ox,oy,oz,tmin,dx,dy,dz,tmax (tmin/tmax carried, never read -- ri_ray_t.min_t/max_t
 * are never read by the BVH, bvh.c:780).  f64 rays: [n][6] = org.xyz, dir.xyz. */
int ri_b200_intersect_batch_f32(ri_b200_accel_t *accel, const float  *rays, uint64_t n, ri_b200_hit_f32 *out);
int ri_b200_occluded_batch_f32 (ri_b200_accel_t *accel, const float  *rays, uint64_t n, uint8_t *out);
int ri_b200_intersect_batch_f64(ri_b200_accel_t *accel, const double *rays, uint64_t n, ri_b200_hit_f64 *out);
int ri_b200_occluded_batch_f64 (ri_b200_accel_t *accel, const double *rays, uint64_t n, uint8_t *out);
/* post-hit state, intersection_state.c:99-248 (geometry carrying only "P") */
int ri_b200_state_batch_f64(ri_b200_accel_t *accel, const double *rays, const ri_b200_hit_f64 *hits, uint64_t n,
                            ri_b200_state_f64 *out);

/* the rest of ri_intersection_state_build (intersection_state.c:123-133, 192-246): E, I = normalize(dir), vertex-colour lerp or
 * (1,1,1), st lerp or 0, the two-sided back-side flag.  Attributes are per INPUT triangle, what ri_geom_t holds behind the index
 * list: tri_colors [ntris][3][3] (geom->colors) with has_color[ntris] (0: the triangle's geom has no Cs), tri_st [ntris][3][2]
 * (geom->texcoords or texcoords_unshared) with has_st[ntris], tri_inside[ntris] (1: geom->two_side && index >= nindices/2,
 * polygon.c:595-612).  Any pointer may be NULL (absent everywhere). */
typedef struct { double E[3], I[3], color[3], st[2], t; int32_t inside, hit; } ri_b200_state_ext_f64;
int ri_b200_set_attributes(ri_b200_accel_t *accel, const double *tri_colors, const uint8_t *has_color, const double *tri_st,
                           const uint8_t *has_st, const uint8_t *tri_inside);
int ri_b200_state_ext_batch_f64(ri_b200_accel_t *accel, const double *rays, const ri_b200_hit_f64 *hits, uint64_t n,
                                ri_b200_state_ext_f64 *out);

/* material texture of the AO transport: ri_transport_ambientocclusion multiplies the radiance by ri_texture_fetch(texture, st) when
 * the hit geom's material has one (ambientocclusion.c:393-401; render/texture.c:86-236 bilinear fetch, wrap by floor, zero texels
 * beyond the last row/column).  rgba: [height][width][4] floats = ri_texture_t.data; tri_textured[ntris] marks the INPUT triangles
 * whose geom carries the texture (NULL: all); the st come from ri_b200_set_attributes (either order).  ri_b200_render_ao then writes
 * three channels.  NULL rgba removes the texture. */
int ri_b200_set_texture(ri_b200_accel_t *accel, const float *rgba, int width, int height, const uint8_t *tri_textured);

/* DEVICE buffers, asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the accelerator's own stream).  These entry
 * points take no lock and may be called from several host threads / streams; each launch takes its work counter from a ring of
 * 1024, so at most 1024 launches of one accelerator may be IN FLIGHT at a time (a counter is re-zeroed when its slot comes round).
 *
 * Ray set-up contract (all query entry points): sign[k] = dir[k] < 0 and inv[k] = |dir[k]| > 1e-14 ? 1 / dir[k] : +-REAL_MAX for ALL
 * three axes -- the rule bvh.c:473-497 intends.  The reference itself leaves invdir[1] UNSET when |dir[1]| <= 1e-14 (it assigns
 * invdir[2] a second time, bvh.c:483-487: stale stack contents decide), so for rays whose y component is exactly 0 (or denormal-small)
 * the reference's answer is undefined and cannot be compared; x- and z-parallel rays are well defined there and equal here
 * (tests/test_gpu_parity.py::test_axis_parallel_rays). */
int ri_b200_intersect_dev_f32(ri_b200_accel_t *accel, const float  *d_rays, uint64_t n, ri_b200_hit_f32 *d_out, void *stream);
int ri_b200_occluded_dev_f32 (ri_b200_accel_t *accel, const float  *d_rays, uint64_t n, uint8_t *d_out, void *stream);
int ri_b200_intersect_dev_f64(ri_b200_accel_t *accel, const double *d_rays, uint64_t n, ri_b200_hit_f64 *d_out, void *stream);
int ri_b200_occluded_dev_f64 (ri_b200_accel_t *accel, const double *d_rays, uint64_t n, uint8_t *d_out, void *stream);

/* the reference's RI_BVH_TRACE_STATISTICS counters (bvh.c:460,829-845,1129-1151) for a HOST batch:
 * precision = RI_B200_PREC_*, anyhit = stop at the first committed hit */
int ri_b200_count_batch(ri_b200_accel_t *accel, const void *rays, uint64_t n, uint32_t precision, int anyhit,
                        ri_b200_counters_t *out);

/* kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t ri_b200_launch_count(void);

/* pinned host memory for callers that want the copies to overlap (cudaHostAlloc / cudaFreeHost) */
void *ri_b200_host_alloc(uint64_t bytes);
void  ri_b200_host_free(void *p);

/* ---- replaces the pixel loop + transport for one frame: render_bucket/subsample (render.c:715-823,
 * 1107-1146), ri_transport_ambientocclusion/calculate_occlusion (ambientocclusion.c:42-151,332-415),
 * bucket_write's float path (render.c:919-979).  rgb_out: HOST [height][width][3] floats, row H-1-y. */
int ri_b200_render_ao(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, float *rgb_out, ri_b200_frame_stats_t *stats);
/* same, framebuffer left in DEVICE memory (d_rgb, [height][width][3]) on `stream`; used by the multi-GPU gather */
int ri_b200_render_ao_dev(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, float *d_rgb, void *stream,
                          ri_b200_frame_stats_t *stats);

/* multi-GPU form: only this rank's pixels, PACKED in visiting order ([npixels][3] floats, the NCCL send buffer -- the
 * resolve kernel writes it directly, no staging copy).  ri_b200_frame_pixels() gives that order for any (rank, world):
 * out[i] = x | y << 16 for the i-th pixel of rank `frame->rank` (spiral bucket b belongs to rank b % world, pixels
 * row-major inside a bucket: render.c:582-710, 1131-1146); pure host logic.  Returns the pixel count. */
int     ri_b200_render_ao_tiles_dev(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, float *d_packed, void *stream,
                                    ri_b200_frame_stats_t *stats);
int64_t ri_b200_frame_pixels(const ri_b200_frame_t *frame, uint32_t *out, int64_t capacity);

/* ---- hemisphere gathers at shading points: the per-point loops of three more ri_raytrace callers, batched.
 * points = [n][6] doubles (P, N); out3 = [n][3] doubles; results equal one reference call per point, in order, with the
 * generator seeded `seed` (4357 is the reference's) and `stream_offset` words already drawn before point 0:
 *   RI_B200_GATHER_OCCLUSION  float occlusion(status, P, N, nsamples)                      shader.c:680-768 (all three channels)
 *   RI_B200_GATHER_IBL        ri_ibl_sample_cosweight(power, N, nsamples, ray, P, eye, l)  ibl.c:53-228, l->texture = env_rgba
 *   RI_B200_GATHER_DOME       ri_domelight_sample(power, hemi, nsamples, ray, P, eye, l)   ibl.c:231-389, l->col, l->intensity
 * (Monte Carlo branches by default; use_qmc selects the quasi-Monte Carlo ones.)  nrays_out (may be NULL) = rays traced. */
#define RI_B200_GATHER_OCCLUSION 0
#define RI_B200_GATHER_IBL       1
#define RI_B200_GATHER_DOME      2
typedef struct {
    int32_t  kind;
    int32_t  nsamples;
    uint32_t seed;
    uint32_t pad_;
    uint64_t stream_offset;
    const float *env_rgba;        /* IBL: angular map [env_height][env_width][4] (ri_texture_t.data), host memory */
    int32_t  env_width, env_height;
    double   col[3];              /* DOME: ri_light_t.col */
    double   intensity;           /* DOME: ri_light_t.intensity */
    /* Option "use_qmc" (option.c:139, 520; off by default): the quasi-Monte Carlo branches of the IBL and dome gathers (ibl.c:107-151,
     * 266-320) -- nsamples rays per point from scrambled Halton / Hammersley points over lucille's Faure permutation table
     * (qmc.c:182-260, 329-396), no random numbers: seed and stream_offset are not used.  qmc_instance[p] = inray->i of point p (NULL:
     * all 0), qmc_dim = inray->d (values < 1 count as 1; primes[qmc_dim] must be below 100, i.e. qmc_dim <= 24).  The occlusion()
     * shadeop has no such branch. */
    int32_t  use_qmc;
    int32_t  qmc_dim;
    const int32_t *qmc_instance;
} ri_b200_gather_t;
int ri_b200_gather_points_f64(ri_b200_accel_t *accel, const ri_b200_gather_t *gather, const double *points, uint64_t n, double *out3,
                              uint64_t *nrays_out);

/* ---- the two shading-language callers of ri_raytrace, batched over shading points.
 *
 * trace(status, dst, P, R) (render/shader.c:895-976) for n (P, R) pairs ([n][6] doubles, R NOT normalised): the ray (P + 0.0001 R, R),
 * closest hit, and per pair what the shadeop does with the answer up to the call of the hit geometry's shader procedure (a host
 * function pointer, which stays with the host).  On a hit: the procedure's input block -- Cs (vertex colours of ri_b200_set_attributes,
 * else 1), P, N (the shading normal Ns), Ng, dPdu / dPdv (tangent / binormal), I = normalize(P_hit - P), s = u, t = v
 * (shader.c:951-968) -- and Ci = 0.  On a miss: Ci = ri_texture_ibl_fetch(env, R) when env_rgba is given (the scene's first light is
 * an IBL / sun-sky light, shader.c:927-940; [env_height][env_width][4] floats, host memory), else zero. */
typedef struct {
    double   Cs[3], P[3], N[3], Ng[3], dPdu[3], dPdv[3], I[3], Ci[3];
    double   t;
    float    s, tt;               /* the shader's s and t */
    uint32_t prim;                /* post-build triangle number, RI_B200_MISS_PRIM on a miss */
    int32_t  hit;
} ri_b200_trace_rec_f64;
int ri_b200_shade_trace_f64(ri_b200_accel_t *accel, const double *pr, uint64_t n, const float *env_rgba, int env_width, int env_height,
                            ri_b200_trace_rec_f64 *out);

/* next_lightsource(status, P, N, angle) + init_lightsource (render/shader.c:1116-1186, 1236-1310): the light samples an
 * `illuminance` loop visits at each of n shading points ([n][6] doubles: P, N).  m = ri_b200_light_samples_count(nsamples) =
 * ntheta * 3 ntheta samples per point, ntheta = (int)sqrt((int)(nsamples / 3.0)) (nsamples = Option narealight_rays); two words of
 * the randomMT stream per sample (seed 4357 in the reference; `stream_offset` words drawn before point 0), L = normalize(direction),
 * Cl = ri_texture_ibl_fetch(env, L) / m.  L_out / Cl_out: [n][m][3] doubles, all m samples in drawing order; visible_out [n][m]:
 * 1 for the samples the reference's loop RETURNS, in that order -- inside the cone (dot(L, N) > 0 and acos(dot) < angle), not
 * occluded along normalize(L) from P + 0.0001 N, and not the last sample of the set, which the reference can never return
 * (shader.c:1170-1177).  nrays_out (may be NULL) = shadow rays traced.  Returns m, < 0 on failure. */
typedef struct {
    int32_t  nsamples;
    uint32_t seed;
    uint64_t stream_offset;
    double   angle;
    const float *env_rgba;        /* [env_height][env_width][4] floats (the IBL light's ri_texture_t.data), host memory; NULL: Cl = 0 */
    int32_t  env_width, env_height;
} ri_b200_light_t;
int ri_b200_light_samples_count(int nsamples);
int ri_b200_light_samples_f64(ri_b200_accel_t *accel, const ri_b200_light_t *light, const double *points, uint64_t n, double *L_out,
                              double *Cl_out, uint8_t *visible_out, uint64_t *nrays_out);

/* ---- calculate_occlusion (transport/ambientocclusion.c:42-151) as one batched call: n shading points (P, Ns) in, the number of
 * OCCLUDED gather rays per point out (the reference's `occlusion` counter; Lo = (N - occluded) / N, N = ntheta * nphi,
 * ambientocclusion.c:143-147).  The N rays of a point are set up on the device exactly as the reference does -- origin P + eps * Ns
 * (eps = 1e-6 in ri_transport_ambientocclusion), basis = ri_ortho_basis(Ns), stratified cosine-weighted directions, outer loop over
 * phi, inner loop over theta -- in double, rounded once to fp32 ray records, and traced by the fp32 occlusion traverser; the host
 * moves 48 bytes per point in and 4 bytes per point out.  Two stated substitutions keep a batch order-free and bit-reproducible
 * (SURVEY 8d, C3): the uniforms are counter-based, u(point, j, i, k) = scenes.uniform01(seed)[((point * N) + j * ntheta + i) * 2 + k],
 * not the next words of the one sequential randomMT2 stream (the frame entry points keep that stream), and sin / cos of 2 pi z are
 * evaluated by a fixed sequence of IEEE multiplies and adds (|error| < 1e-15) instead of libm.
 *   points         [n][6] doubles: P.xyz, Ns.xyz (host memory; the _dev form takes device memory and is asynchronous on `stream`)
 *   occluded_out   [n] uint32
 * ri_b200_ao_point_rays_f32 returns the generated batch itself, [n * N][8] fp32 ray records in the order they are traced. */
typedef struct {
    int32_t  ntheta, nphi;        /* (int)sqrt(gather_nsamples) each, ambientocclusion.c:378-387 */
    uint64_t seed;
    double   eps;                 /* origin offset along Ns */
} ri_b200_ao_points_t;
int ri_b200_occlusion_points_f32(ri_b200_accel_t *accel, const ri_b200_ao_points_t *params, const double *points, uint64_t n,
                                 uint32_t *occluded_out);
int ri_b200_occlusion_points_dev_f32(ri_b200_accel_t *accel, const ri_b200_ao_points_t *params, const double *d_points, uint64_t n,
                                     uint32_t *d_occluded, void *stream);
int ri_b200_ao_point_rays_f32(ri_b200_accel_t *accel, const ri_b200_ao_points_t *params, const double *points, uint64_t n, float *rays_out);
/* The same call in the reference's own precision: the N rays of a point are set up as above and traced as DOUBLE rays against the
 * double records -- every count equals what the double reference finds for those rays (oracle: orc_ao_point_rays_f64 + the f64
 * traversal).  This is the entry for real scenes, where fp32 ray records cannot hold an origin 1e-6 above a surface. */
int ri_b200_occlusion_points_f64(ri_b200_accel_t *accel, const ri_b200_ao_points_t *params, const double *points, uint64_t n,
                                 uint32_t *occluded_out);
int ri_b200_occlusion_points_dev_f64(ri_b200_accel_t *accel, const ri_b200_ao_points_t *params, const double *d_points, uint64_t n,
                                     uint32_t *d_occluded, void *stream);

/* rng_mode 0 (the reference's single MT19937 stream, random.c:211-247) on world > 1.  The stream position of a gather ray depends
 * on how many eye samples hit something in every bucket the reference renders EARLIER (render.c:1131-1146 in spiral order), and
 * those buckets belong to other ranks: between the eye pass and the gather pass the frame call hands the host this rank's
 * hit-sample count per bucket (buckets rank, rank + world, ... of the spiral order, in that order) and wants back, for each of
 * them, the number of hit samples in all buckets of the frame that precede it, plus the frame's total.  The host implements it with
 * whatever it has between ranks (an all-gather of a few hundred integers: lucille_b200/distributed.py uses torch.distributed;
 * lucille itself would use ri_parallel_gather + ri_parallel_bcast, parallel.c:102-196).  Return 0, or non-zero to fail the frame.
 * The call is a collective of the host's: every rank of the frame makes it exactly once per frame.  A rank whose frame call fails
 * BEFORE it has counts still makes it, with bucket_hits == NULL and both output pointers NULL ("this rank has failed"): the host's
 * exchange must then fail on every rank (distributed.hit_exchange carries a failure flag through its all-reduce), so that nobody
 * waits for a rank that has gone.  The callback runs on the calling thread while the accelerator's lock is held: it must not call
 * back into this accelerator.
 * With it the multi-GPU frame equals the reference's single-thread frame bit for bit, like the one-GPU frame. */
typedef int (*ri_b200_hit_exchange_fn)(void *user, const uint32_t *bucket_hits, uint32_t nbuckets, uint64_t *bucket_base_out,
                                       uint64_t *frame_hits_out);
int ri_b200_set_hit_exchange(ri_b200_accel_t *accel, ri_b200_hit_exchange_fn fn, void *user);

/* fused multi-GPU resolve (one process per GPU of ONE node): rank 0 allocates the framebuffer with ri_b200_peer_alloc ([h][w][3]
 * floats, zeroed) and hands the 64-byte CUDA IPC handle to the other ranks (any transport: torch.distributed, MPI, a pipe); they
 * map it with ri_b200_peer_open.  ri_b200_render_ao_peer_dev renders the buckets of frame->rank and its resolve kernel stores them
 * at their framebuffer positions in that shared buffer -- over NVLink / NVSwitch peer memory for ranks > 0 -- without clearing
 * it: the gather that would follow the frame is done by the stores.  After every rank's stream has drained (and a barrier), rank 0
 * reads the frame with ri_b200_peer_read.  ri_b200_peer_close (ranks > 0) / ri_b200_peer_free (rank 0) release it. */
void *ri_b200_peer_alloc(uint64_t bytes, int device, uint8_t handle_out[64]);
void *ri_b200_peer_open(const uint8_t handle[64], int device);
int   ri_b200_peer_close(void *p, int device);
int   ri_b200_peer_free(void *p, int device);
int   ri_b200_peer_read(const void *p, void *host, uint64_t bytes, int device);
int   ri_b200_render_ao_peer_dev(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, float *d_rgb_shared, void *stream,
                                 ri_b200_frame_stats_t *stats);

/* ---- sun-sky variant of the same transport: gather_sunsky + contribution_from_sunlight (ambientocclusion.c:153-324), taken by
 * ri_transport_ambientocclusion whenever the scene has an AreaLightSource "sunsky" (ambientocclusion.c:369-376): fixed 8 x 8
 * gather (frame->ntheta/nphi are ignored), origin offset 1e-5, the sky colour ri_sunsky_get_sky_rgb(dir) (render/sunsky.c:322-408)
 * summed over the MISSED rays, one shadow ray per LIGHTTYPE_SUNLIGHT light, Lo = (1/pi) * col / 64, three channels.
 * The block is what the host owns after ri_sunsky_init() (sunsky.c:176-295) -- set-up stays on the reference side:
 *   sun_theta .. zenith_Y   ri_sunsky_t fields of scene->sunsky_light->sunsky (sunsky.h:22-50)
 *   S0 S1 S2                S0Amplitudes / S1Amplitudes / S2Amplitudes (render/sunsky.dat)
 *   cie                     cie_colour_match (render/specrend.c:387-414)
 *   cs                      CIEsystem: xRed yRed xGreen yGreen xBlue yBlue xWhite yWhite (render/specrend.h)
 *   sun_dir / sun_col       ri_light_t.direction / .col of the scene's LIGHTTYPE_SUNLIGHT lights, light_list order (lightsource.c:152-170)
 * The lookup is float arithmetic around double libm calls; the device's sin/cos/acos/atan2/exp differ from the host's in the
 * last place, so parity with the reference is to float tolerance, not bit-exact (tests/test_gpu_parity.py states it). */
typedef struct {
    float   sun_theta, sun_phi;
    float   perez_x[5], perez_y[5], perez_Y[5];
    float   zenith_x, zenith_y, zenith_Y;
    float   S0[41], S1[41], S2[41];
    float   cie[81][3];
    float   cs[8];
    int32_t nsun;                 /* 0..4 */
    int32_t pad;
    double  sun_dir[4][3];
    double  sun_col[4][3];
} ri_b200_sunsky_t;
int ri_b200_render_sunsky(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, const ri_b200_sunsky_t *sky, float *rgb_out,
                          ri_b200_frame_stats_t *stats);
int ri_b200_render_sunsky_tiles_dev(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, const ri_b200_sunsky_t *sky, float *d_packed,
                                    void *stream, ri_b200_frame_stats_t *stats);
/* ---- dirt-map transport: ri_transport_dirtmap + calculate_dirt (transport/dirtmap.c:84-293), compiled into
 * lucille but not wired into its pixel loop (render.c:800-804).  Same frame description; fixed 4 x 4 gather (frame->ntheta/nphi are
 * ignored), origin offset 1e-5, CLOSEST hits: black within 0.1, white beyond 0.5 or on a miss, linear mix in between; times the
 * material texture when one is set. */
int ri_b200_render_dirtmap(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, float *rgb_out, ri_b200_frame_stats_t *stats);

/* ---- Whitted transport: ri_transport_whitted + trace_whitted (transport/whitted.c:31-151): chains of refracted
 * rays (eta 1.33, at most 8 bounces) from every eye hit, radiance = the angular-map environment along the direction that leaves the
 * scene (ri_texture_ibl_fetch, render/texture.c:238-277), zero when the chain is cut.  env_rgba: HOST [h][w][4] floats =
 * scene->envmap_light->texture->data, or NULL (no environment).  fp64 records; whole frames (world == 1). */
int ri_b200_render_whitted(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, const float *env_rgba, int env_width, int env_height,
                           float *rgb_out, ri_b200_frame_stats_t *stats);

/* ri_transport_sample / trace_path (transport/transport.c:50-173): white where the eye ray hits, black elsewhere.  fp64 records. */
int ri_b200_render_sample(ri_b200_accel_t *accel, const ri_b200_frame_t *frame, float *rgb_out, ri_b200_frame_stats_t *stats);

/* ri_sunsky_get_sky_rgb for a HOST batch of directions ([n][3] floats in, [n][3] floats out), computed on `device` */
int ri_b200_sunsky_rgb(const ri_b200_sunsky_t *sky, const float *dirs, uint64_t n, float *rgb_out, int device);

/* ---- the output step right after the path: lucille's Radiance .hdr display driver (display/hdrdrv.c:62-102 hdr_dd_write clamp +
 * accumulate, hdr_dd_close; imageio/rgbe.c:78-96 float2rgbe, 117-139 RGBE_WriteHeader, 244-340 RGBE_WritePixels_RLE) computed on the
 * device.  rgb: [height][width][3] floats in display order -- a DEVICE pointer when rgb_on_device != 0 (e.g. the buffer
 * ri_b200_render_ao_dev filled), else a HOST pointer.  Returns the size of the file image in bytes (or -1) and writes it to `out`
 * (HOST) when cap is large enough; byte-identical to the file the reference writes. */
int64_t ri_b200_hdr_encode(const float *rgb, int width, int height, uint8_t *out, uint64_t cap, int device, int rgb_on_device);

/* The other float display driver: the byte stream lucille's socket driver sends to its viewer for a finished frame
 * (display/sockdrv.c:118-262 sock_dd_open / sock_dd_write / sock_dd_close, sockdrv_defs.h), packed on the device -- header, one
 * message per 1024 pixels in bucket_write's order (render.c:919-979; frame->bucket_size), finish command; the remainder below 1024
 * pixels is never sent by the reference and is not here either.  The stream only: connect() and send() stay with the host.  rgb as
 * for ri_b200_hdr_encode; frame gives width, height and bucket_size.  Returns the stream's size (writes it when cap suffices), -1 on
 * failure. */
int64_t ri_b200_sockdrv_encode(const float *rgb, const ri_b200_frame_t *frame, uint8_t *out, uint64_t cap, int device, int rgb_on_device);

/* ---- replaces ri_beam_set + ri_bvh_intersect_beam_visibility (beam.c:332-466, bvh.c:612-667) for a batch of beams.
 * beams: HOST [n][15] doubles = org.xyz, dir0.xyz .. dir3.xyz (consecutive corners of the frustum).  out[i] = RI_BEAM_MISS_COMPLETELY 0 /
 * RI_BEAM_HIT_COMPLETELY 1 / RI_BEAM_HIT_PARTIALLY 2 (beam.h:27-29), or -1 where ri_beam_set would fail (corner directions
 * not in one octant, beam.c:356-377).  fp64 records. */
int ri_b200_beam_visibility_batch(ri_b200_accel_t *accel, const double *beams, uint64_t n, int32_t *out);

/* ---- path-trace transport (src/transport/pathtrace.c:131-537; NOT in the reference build): the sketch's control
 * flow with builder-stated inputs -- Lambert kd, constant environment Le, counter-based RNG keyed by (pixel, sample, draw),
 * deterministic sin/cos.  fp64 records.  rgb_out: HOST [height][width][3], row H-1-y (pathtrace.c:183). */
typedef struct {
    double  c2w[16];
    double  flength;
    int32_t is_rh;
    int32_t width, height;
    int32_t spp;                  /* ri_option_t.pt_nsamples (option.c:142, 551-556) */
    int32_t max_vertices;         /* MAX_PATH_VERTICES 10 (pathtrace.c:66) */
    uint32_t seed;
    double  kd, Le;
    int32_t rank, world;          /* tile sharding as in ri_b200_frame_t */
    int32_t bucket_size;
} ri_b200_path_frame_t;
int ri_b200_render_pathtrace(ri_b200_accel_t *accel, const ri_b200_path_frame_t *frame, float *rgb_out, ri_b200_frame_stats_t *stats);
int ri_b200_render_pathtrace_tiles_dev(ri_b200_accel_t *accel, const ri_b200_path_frame_t *frame, float *d_packed, void *stream,
                                       ri_b200_frame_stats_t *stats);

/* MT19937 stream of randomMT2() generated on the device (random.c:98-112,211-247); HOST output, for tests.
 * accel == NULL: one CTA walks the stream on `device`; accel != NULL: the frame path -- window states at segment starts
 * by GF(2) jump-ahead (doubling tree, cached per seed in the accelerator), then one CTA per 638 976-word segment. */
int ri_b200_mt_stream(ri_b200_accel_t *accel, uint32_t seed, uint64_t n, uint32_t *out_u32, int device);
/* Optional warm-up: every entry point that draws from the reference's MT19937 stream (frames with rng_mode 0, ri_b200_gather_points_f64,
 * ri_b200_light_samples_f64) generates the stream in parallel from a table of generator states at segment starts, built by GF(2)
 * jump-ahead on first use and cached per (accelerator, seed).  Building it costs about 5 ms per doubling of the stream length beyond
 * 638 976 words; this call builds it ahead of time for streams of up to max_words words, asynchronously on the accelerator's stream
 * (lucille seeds its generators once, 4357: call it once after ri_b200_build). */
int ri_b200_mt_prepare(ri_b200_accel_t *accel, uint32_t seed, uint64_t max_words);

#ifdef __cplusplus
}
#endif
#endif /* LUCILLE_B200_H */
