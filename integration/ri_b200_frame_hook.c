/*
 * ri_b200_frame_hook.c -- the BATCHED frame hook: the reference renderer's whole pixel loop as one call into the GPU library.
 *
 * What it replaces.  ri_render_frame() (src/render/render.c:318-360) sets the frame up (ri_render_setup, ri_scene_setup -- which
 * builds the accelerator through ri_accel_t.build --, ri_camera_setup, create_bucket_list), then runs render_frame_controller()
 * (render.c:1168-1207): `nthreads` workers pop 32x32 buckets and call render_bucket -> subsample -> ri_transport_ambientocclusion ->
 * ri_raytrace for every primary and gather ray (render.c:715-823, 1043-1146; ambientocclusion.c:42-151, 332-415), and bucket_write
 * hands every pixel to the display driver (render.c:919-979).  With the per-ray vtable slot alone (ri_b200_binding.c) each of those
 * ri_raytrace calls is one kernel launch; here the frame is ONE ri_b200_render_ao() call and the display driver receives the same
 * floats in the same (bucket, row, column) order.
 *
 * How it is spliced in WITHOUT touching the reference sources (this repository compiles /root/reference where it lies).  In a lucille
 * checkout the body of b200_render_frame_batched() goes into ri_render_frame() in place of the render_frame_controller() call, as
 * INTEGRATION.md section 2 shows.  render_frame_controller and its worker function are `static`, so here the splice is made at the
 * two PUBLIC functions on that path, interposed at link time (-Wl,--wrap=...):
 *     ri_render_frame   -> __wrap_ri_render_frame   marks "a frame is being rendered" around the real function;
 *     ri_thread_create  -> __wrap_ri_thread_create  (base/thread.c:226) the first worker the controller starts while a frame is being
 *                          rendered with the B200 accelerator renders the WHOLE frame in one batched call and drains the bucket
 *                          queue; that worker and the others then have nothing to do.  Everything else ri_render_frame does -- timers,
 *                          statistics, display close, scene free -- runs unchanged.
 * render->stat.nrays receives the rays the GPU traced, so ri_raytrace_statistics() (raytrace.c:71-112) prints the frame's real
 * "M Rays/sec" line.  Scenes the batched call does not cover (sun-sky light, orthographic camera, non-float displays, RI_B200_FRAME=0)
 * fall through to the reference's own loop on the per-ray slot.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "ri.h"
#include "render.h"
#include "scene.h"
#include "accel.h"
#include "camera.h"
#include "display.h"
#include "option.h"
#include "queue.h"
#include "thread.h"
#include "list.h"
#include "log.h"
#include "memory.h"

#include "ri_b200_binding.h"

extern void __real_ri_render_frame(void);
extern int  __real_ri_thread_create(ri_thread_t *thread, void *(*func)(void *), void *arg);

static int g_in_frame = 0;        /* between entry and exit of ri_render_frame (one frame at a time: render.c:70 `grender`) */
static int g_frame_batched = 0;   /* this frame went through ri_b200_render_ao */

typedef struct { int x, y, w, h; } b200_bucket_t;     /* leading fields of bucket_t (render.c:88-96): region of the bucket */

/* 1 when the batched call covers this frame's configuration */
static int b200_frame_applies(const ri_render_t *render)
{
    const ri_option_t  *opt  = render->context->option;
    const ri_display_t *disp = ri_option_get_curr_display((ri_option_t *)opt);
    const char *env = getenv("RI_B200_FRAME");
    if (env && atoi(env) == 0) return 0;
    if (opt->accel_method != RI_ACCEL_B200) return 0;
    if (!render->scene || !render->scene->accel || !render->scene->accel->data) return 0;
    if (render->bucket_order != BUCKET_ORDER_SPIRAL) return 0;                   /* the only order the renderer ever selects, render.c:198 */
    if (render->scene->sunsky_light) return 0;                                   /* ri_b200_render_sunsky: INTEGRATION.md */
    if (opt->camera->camera_projection == RI_ORTHOGRAPHIC) return 0;
    if (!(strcmp(disp->display_format, "float") == 0 || strcmp(disp->display_type, "hdr") == 0 ||
          strcmp(disp->display_type, "openexr") == 0 || strcmp(disp->display_type, "socket") == 0 ||
          strcmp(disp->display_type, RI_FILE) == 0)) return 0;                   /* bucket_write's float path, render.c:948-963 */
    return 1;
}

static void b200_render_frame_batched(ri_render_t *render)
{
    ri_option_t     *opt  = render->context->option;
    ri_camera_t     *cam  = opt->camera;
    ri_display_t    *disp = ri_option_get_curr_display(opt);
    b200_binding_t  *b    = (b200_binding_t *)render->scene->accel->data;
    const int        w = cam->horizontal_resolution, h = cam->vertical_resolution;
    ri_b200_frame_t  f;
    ri_b200_frame_stats_t st;
    float           *rgb = (float *)ri_mem_alloc(sizeof(float) * 3 * (size_t)w * (size_t)h);
    void            *item;
    uint32_t         item_size;
    int              i, j, sx, sy;

    memset(&f, 0, sizeof(f));
    for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) f.c2w[4 * i + j] = cam->camera_to_world.f[i][j];      /* camera.c:214-245 */
    f.flength = cam->flength;
    f.is_rh   = cam->is_rh;
    f.width = w; f.height = h;
    f.xsamples = (int)disp->sampling_rates[0];
    f.ysamples = (int)disp->sampling_rates[1];
    f.ntheta = f.nphi = (int)sqrt((double)opt->gather_nsamples);                   /* ambientocclusion.c:378-387 */
    f.bucket_size = render->bucket_size;
    f.rng_mode = 0; f.seed = 4357;        /* the reference's one randomMT2 stream in its consumption order: the --nthreads 1 image */
    f.rank = 0; f.world = 1;
    f.precision = RI_B200_PREC_F64;       /* reference-exact records */
    memset(&st, 0, sizeof(st));
    if (ri_b200_render_ao(b->dev, &f, rgb, &st) < 0) {
        ri_log(LOG_FATAL, "(B200  ) %s", ri_b200_last_error());
        abort();
    }
    render->stat.nrays += st.nrays_primary + st.nrays_ao;                          /* raytrace.c:43: one count per ri_raytrace() */

    /* the display driver sees what bucket_write() would have sent: buckets in queue order, rows, then columns (render.c:944-965);
     * rgb is already row-flipped like `screenheight - (sy + y) - 1` */
    while (ri_mt_queue_pop(render->bucket_queue, &item, &item_size) == 0) {
        const b200_bucket_t *bk = (const b200_bucket_t *)item;
        for (sy = 0; sy < bk->h; sy++)
            for (sx = 0; sx < bk->w; sx++) {
                const int X = sx + bk->x, Y = h - (sy + bk->y) - 1;
                render->display_drv->write(X, Y, &rgb[3 * ((size_t)Y * w + X)]);
            }
    }
    ri_mem_free(rgb);
    ri_log(LOG_INFO, "(B200  ) frame rendered in one batched call: %llu rays, %.2f ms on the device",
           (unsigned long long)(st.nrays_primary + st.nrays_ao), st.ms_total);
}

static void *b200_idle_worker(void *arg) { (void)arg; return NULL; }

int __wrap_ri_thread_create(ri_thread_t *thread, void *(*func)(void *), void *arg)
{
    if (g_in_frame) {
        ri_render_t *render = ri_render_get();
        if (!g_frame_batched && b200_frame_applies(render)) {
            b200_render_frame_batched(render);         /* in the caller's thread: the controller is about to wait for its workers anyway */
            g_frame_batched = 1;
        }
        if (g_frame_batched) return __real_ri_thread_create(thread, b200_idle_worker, arg);      /* the queue is empty: nothing left to render */
    }
    return __real_ri_thread_create(thread, func, arg);
}

void __wrap_ri_render_frame(void)
{
    g_in_frame = 1; g_frame_batched = 0;
    __real_ri_render_frame();
    g_in_frame = 0;
}
