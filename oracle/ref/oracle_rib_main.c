/*
 * TEST INFRASTRUCTURE (oracle/_ref) -- not part of the product.
 *
 * `oracle_rib`: renders one RIB with the UNMODIFIED reference renderer (via ref_shim.c /
 * rib_reader.c) and writes the captured float framebuffer and the flattened scene.
 * Stands in for `lsh` (src/lsh/main.c), whose flex/bison front end cannot be generated here.
 *
 *   oracle_rib scene.rib [--nthreads N] [--width W --height H] [--pixelsamples P]
 *                        [--gather G] [--out frame.bin] [--scene scene.bin] [--sunsky sunsky.bin] [--accel bvh|b200]
 *
 * frame.bin : "LFRM" u32 w, u32 h, u32 0, f64 seconds("Render frame"), u64 nrays(stat.nrays), f32 rgb[h][w][3]
 * sunsky.bin: f64[45] = lref_frame_sunsky() block (only written when the scene has an AreaLightSource "sunsky")
 * --texture tex.bin (in): "LTEX" u32 w, u32 h, f32 rgba[h][w][4] -> material texture of every geom of the frame
 * --attr attr.bin (out): f64 st[ntris][3][2], u8 has_st[ntris]
 * scene.bin : "LSCN" u32 0, u64 ntris, f64 cam[27], f64 tri[ntris][9], u32 geom[ntris], f64 normals[ntris][9] (zeros if none)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

extern void     lref_set_accel_method(int m);
extern int      lref_render_rib(const char *path, int nthreads, int width, int height, int pixelsamples, int gather);
extern int      lref_frame_width(void);
extern int      lref_frame_height(void);
extern float   *lref_frame_rgb(void);
extern double   lref_frame_seconds(void);
extern uint64_t lref_frame_nrays(void);
extern uint64_t lref_frame_ntris(void);
extern double  *lref_frame_tris(void);
extern uint32_t*lref_frame_trigeom(void);
extern double  *lref_frame_normals(void);
extern double  *lref_frame_st(void);
extern uint8_t *lref_frame_has_st(void);
extern void     lref_set_frame_texture(const float *rgba, int width, int height);
extern void     lref_frame_camera(double *out27);
extern int      lref_frame_sunsky(double *out45);

int main(int argc, char **argv)
{
    const char *rib = NULL, *out = NULL, *scene = NULL, *sunsky = NULL, *texture = NULL, *attr = NULL;
    int nthreads = 1, width = 0, height = 0, ps = 0, gather = 0, i;
    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--nthreads") && i + 1 < argc) nthreads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--width") && i + 1 < argc) width = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--height") && i + 1 < argc) height = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--pixelsamples") && i + 1 < argc) ps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gather") && i + 1 < argc) gather = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--accel") && i + 1 < argc) {     /* bvh (default) | b200: Option "raytrace" "accel_method" */
            const char *m = argv[++i];
            lref_set_accel_method(!strcmp(m, "b200") ? 2 : 1);
        }
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--scene") && i + 1 < argc) scene = argv[++i];
        else if (!strcmp(argv[i], "--sunsky") && i + 1 < argc) sunsky = argv[++i];
        else if (!strcmp(argv[i], "--texture") && i + 1 < argc) texture = argv[++i];
        else if (!strcmp(argv[i], "--attr") && i + 1 < argc) attr = argv[++i];
        else rib = argv[i];
    }
    if (!rib) { fprintf(stderr, "usage: oracle_rib scene.rib [options]\n"); return 2; }
    if (texture) {
        FILE *fp = fopen(texture, "rb");
        char magic[4];
        uint32_t tw = 0, th = 0;
        float *rgba;
        if (!fp || fread(magic, 1, 4, fp) != 4 || memcmp(magic, "LTEX", 4) || fread(&tw, 4, 1, fp) != 1 || fread(&th, 4, 1, fp) != 1) return 2;
        rgba = (float *)malloc(sizeof(float) * 4 * (size_t)tw * th);
        if (fread(rgba, sizeof(float) * 4, (size_t)tw * th, fp) != (size_t)tw * th) return 2;
        fclose(fp);
        lref_set_frame_texture(rgba, (int)tw, (int)th);
        free(rgba);
    }
    if (lref_render_rib(rib, nthreads, width, height, ps, gather) != 0) return 1;

    {
        uint32_t w = (uint32_t)lref_frame_width(), h = (uint32_t)lref_frame_height(), zero = 0;
        double sec = lref_frame_seconds();
        uint64_t nrays = lref_frame_nrays();
        fprintf(stderr, "\n[oracle_rib] %ux%u  rays=%llu  render=%.3fs  (%.3f Mrays/s, %d thread%s)\n",
                w, h, (unsigned long long)nrays, sec, nrays / 1e6 / (sec > 0 ? sec : 1), nthreads, nthreads > 1 ? "s" : "");
        if (out) {
            FILE *fp = fopen(out, "wb");
            if (!fp) return 1;
            fwrite("LFRM", 1, 4, fp); fwrite(&w, 4, 1, fp); fwrite(&h, 4, 1, fp); fwrite(&zero, 4, 1, fp);
            fwrite(&sec, 8, 1, fp); fwrite(&nrays, 8, 1, fp);
            fwrite(lref_frame_rgb(), sizeof(float), (size_t)w * h * 3, fp);
            fclose(fp);
        }
        if (sunsky) {
            double blk[45];
            memset(blk, 0, sizeof(blk));
            if (lref_frame_sunsky(blk)) {
                FILE *fp = fopen(sunsky, "wb");
                if (!fp) return 1;
                fwrite(blk, 8, 45, fp);
                fclose(fp);
            }
        }
        if (attr) {
            uint64_t n = lref_frame_ntris();
            FILE *fp = fopen(attr, "wb");
            if (!fp) return 1;
            fwrite(lref_frame_st(), 8, (size_t)n * 6, fp);
            fwrite(lref_frame_has_st(), 1, (size_t)n, fp);
            fclose(fp);
        }
        if (scene) {
            double cam[27];
            uint64_t n = lref_frame_ntris();
            FILE *fp = fopen(scene, "wb");
            if (!fp) return 1;
            lref_frame_camera(cam);
            fwrite("LSCN", 1, 4, fp); fwrite(&zero, 4, 1, fp); fwrite(&n, 8, 1, fp);
            fwrite(cam, 8, 27, fp);
            fwrite(lref_frame_tris(), 8, (size_t)n * 9, fp);
            fwrite(lref_frame_trigeom(), 4, (size_t)n, fp);
            fwrite(lref_frame_normals(), 8, (size_t)n * 9, fp);
            fclose(fp);
        }
    }
    return 0;
}
