/*
 * TEST INFRASTRUCTURE (oracle/_ref front end) -- not part of the product.
 *
 * A small RIB-subset reader that drives the UNMODIFIED reference renderer through its
 * own RenderMan entry points (Ri*V in /root/reference/src/ri/ri.c).  It stands in for
 * the flex/bison front end of `lsh` (src/lsh/lexrib.l, src/lsh/parserib.y), which cannot
 * be generated in this image (no flex/bison).  Only the commands the two example scenes
 * use are understood; everything downstream of the Ri call is reference code.
 *
 * Conventions mirrored from the reference front end:
 *   - every NUM becomes an RtFloat (float) via atof()           lexrib.l:212-214, parserib.y:117-124
 *   - parameter lists are (token, float[] | string[]) pairs     parserib.y:291-351
 *   - ReadArchive resolves through ri_option_find_file()         lexrib.l:54-101
 *   - PointsPolygons is dropped when sum(nverts) != len(verts)   parserib.y:573-634
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#include "ri.h"
#include "render.h"
#include "option.h"
#include "log.h"

typedef enum { T_EOF, T_ID, T_STR, T_NUM, T_LB, T_RB } tok_kind;

typedef struct {
    FILE *fp;
    int   have_peek;
    tok_kind peek_kind;
    char  peek_text[4096];
    tok_kind kind;
    char  text[4096];
} lexer_t;

static tok_kind lex_raw(lexer_t *lx, char *out)
{
    int c;
    for (;;) {
        c = fgetc(lx->fp);
        if (c == EOF) return T_EOF;
        if (c == '#') { while ((c = fgetc(lx->fp)) != EOF && c != '\n') {} continue; }
        if (isspace(c)) continue;
        break;
    }
    if (c == '[') return T_LB;
    if (c == ']') return T_RB;
    if (c == '"') {
        int n = 0;
        while ((c = fgetc(lx->fp)) != EOF && c != '"') { if (n < 4095) out[n++] = (char)c; }
        out[n] = 0;
        return T_STR;
    }
    {
        int n = 0;
        out[n++] = (char)c;
        while ((c = fgetc(lx->fp)) != EOF && !isspace(c) && c != '[' && c != ']' && c != '"' && c != '#') {
            if (n < 4095) out[n++] = (char)c;
        }
        if (c != EOF) ungetc(c, lx->fp);
        out[n] = 0;
        if (isdigit((unsigned char)out[0]) || out[0] == '-' || out[0] == '+' || out[0] == '.') return T_NUM;
        return T_ID;
    }
}

static tok_kind lex_next(lexer_t *lx)
{
    if (lx->have_peek) {
        lx->have_peek = 0;
        lx->kind = lx->peek_kind;
        strcpy(lx->text, lx->peek_text);
        return lx->kind;
    }
    lx->kind = lex_raw(lx, lx->text);
    return lx->kind;
}

static tok_kind lex_peek(lexer_t *lx)
{
    if (!lx->have_peek) {
        lx->peek_kind = lex_raw(lx, lx->peek_text);
        lx->have_peek = 1;
    }
    return lx->peek_kind;
}

/* growable float array */
typedef struct { RtFloat *v; int n, cap; } farr_t;
static void farr_push(farr_t *a, RtFloat x)
{
    if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 64; a->v = (RtFloat *)realloc(a->v, sizeof(RtFloat) * a->cap); }
    a->v[a->n++] = x;
}

static RtFloat num_of(const char *s) { return (RtFloat)atof(s); }   /* lexrib.l:213 */

/* NUM | [ NUM* ] */
static int read_num_array(lexer_t *lx, farr_t *a)
{
    a->v = NULL; a->n = 0; a->cap = 0;
    if (lex_next(lx) == T_NUM) { farr_push(a, num_of(lx->text)); return 1; }
    if (lx->kind != T_LB) return 0;
    while (lex_next(lx) == T_NUM) farr_push(a, num_of(lx->text));
    return lx->kind == T_RB;
}

#define MAX_PARAMS 64
typedef struct {
    int       n;
    RtToken   tokens[MAX_PARAMS];
    RtPointer args[MAX_PARAMS];
} plist_t;

/* ( STRING ( NUM | STRING | [ NUM* ] | [ STRING* ] ) )*   parserib.y:282-351 */
static void read_param_list(lexer_t *lx, plist_t *pl)
{
    pl->n = 0;
    while (lex_peek(lx) == T_STR) {
        char name[4096];
        lex_next(lx);
        strcpy(name, lx->text);
        tok_kind k = lex_peek(lx);
        RtPointer arg = NULL;
        if (k == T_NUM) {
            lex_next(lx);
            RtFloat *f = (RtFloat *)malloc(sizeof(RtFloat));
            *f = num_of(lx->text);
            arg = f;
        } else if (k == T_STR) {
            lex_next(lx);
            RtToken *s = (RtToken *)malloc(sizeof(RtToken));
            *s = strdup(lx->text);
            arg = s;
        } else if (k == T_LB) {
            lex_next(lx);
            if (lex_peek(lx) == T_STR) {
                int n = 0, cap = 8;
                RtToken *s = (RtToken *)malloc(sizeof(RtToken) * cap);
                while (lex_next(lx) == T_STR) {
                    if (n == cap) { cap *= 2; s = (RtToken *)realloc(s, sizeof(RtToken) * cap); }
                    s[n++] = strdup(lx->text);
                }
                arg = s;
            } else {
                farr_t a = {0, 0, 0};
                while (lex_next(lx) == T_NUM) farr_push(&a, num_of(lx->text));
                arg = a.v;
            }
        } else {
            break;
        }
        if (pl->n < MAX_PARAMS) {
            pl->tokens[pl->n] = strdup(name);
            pl->args[pl->n]   = arg;
            pl->n++;
        }
    }
}

static int read_nums(lexer_t *lx, RtFloat *out, int n)
{
    int i;
    for (i = 0; i < n; i++) {
        if (lex_next(lx) != T_NUM) return 0;
        out[i] = num_of(lx->text);
    }
    return 1;
}

static void to_matrix(RtMatrix m, const RtFloat *v)
{
    int i, j;
    for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) m[i][j] = v[4 * i + j];
}

int lref_rib_parse_file(const char *path);

static void do_read_archive(const char *name)
{
    char full[2048];
    ri_option_t *opt = ri_render_get()->context->option;
    if (!name || !name[0]) return;
    if (!ri_option_find_file(full, opt, name)) {
        fprintf(stderr, "[rib_reader] ReadArchive: can't find \"%s\"\n", name);
        return;
    }
    lref_rib_parse_file(full);
}

int lref_rib_parse_file(const char *path)
{
    lexer_t *lx = (lexer_t *)calloc(1, sizeof(lexer_t));
    plist_t  pl;
    RtFloat  f[16];
    char     s1[4096], s2[4096], s3[4096];

    lx->fp = fopen(path, "r");
    if (!lx->fp) { fprintf(stderr, "[rib_reader] can't open %s\n", path); free(lx); return -1; }

    while (lex_next(lx) != T_EOF) {
        if (lx->kind != T_ID) continue;               /* stray token: skip (lexer SKIP mode) */
        const char *c = lx->text;

        if (!strcmp(c, "version")) { read_nums(lx, f, 1); }
        else if (!strcmp(c, "Display")) {
            lex_next(lx); strcpy(s1, lx->text);
            lex_next(lx); strcpy(s2, lx->text);
            lex_next(lx); strcpy(s3, lx->text);
            read_param_list(lx, &pl);
            RiDisplayV(strdup(s1), strdup(s2), strdup(s3), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "Format"))       { if (read_nums(lx, f, 3)) RiFormat((RtInt)f[0], (RtInt)f[1], f[2]); }
        else if (!strcmp(c, "PixelSamples")) { if (read_nums(lx, f, 2)) RiPixelSamples(f[0], f[1]); }
        else if (!strcmp(c, "Shutter"))      { if (read_nums(lx, f, 2)) RiShutter(f[0], f[1]); }
        else if (!strcmp(c, "ShadingRate"))  { if (read_nums(lx, f, 1)) RiShadingRate(f[0]); }
        else if (!strcmp(c, "Sides"))        { if (read_nums(lx, f, 1)) RiSides((RtInt)f[0]); }
        else if (!strcmp(c, "ShadingInterpolation")) { lex_next(lx); RiShadingInterpolation(strdup(lx->text)); }
        else if (!strcmp(c, "Orientation"))  { lex_next(lx); RiOrientation(strdup(lx->text)); }
        else if (!strcmp(c, "Projection")) {
            lex_next(lx); strcpy(s1, lx->text);
            read_param_list(lx, &pl);
            RiProjectionV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "Option")) {
            lex_next(lx); strcpy(s1, lx->text);
            read_param_list(lx, &pl);
            RiOptionV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "Attribute")) {
            lex_next(lx); strcpy(s1, lx->text);
            read_param_list(lx, &pl);
            RiAttributeV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "AreaLightSource")) {          /* parserib.y:371-375: name, sequence number, parameters */
            lex_next(lx); strcpy(s1, lx->text);
            read_nums(lx, f, 1);
            read_param_list(lx, &pl);
            RiAreaLightSourceV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "Surface")) {
            lex_next(lx); strcpy(s1, lx->text);
            read_param_list(lx, &pl);
            RiSurfaceV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "Atmosphere")) {
            lex_next(lx); strcpy(s1, lx->text);
            read_param_list(lx, &pl);
            RiAtmosphereV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "Imager")) {
            lex_next(lx); strcpy(s1, lx->text);
            read_param_list(lx, &pl);
            RiImagerV(strdup(s1), pl.n, pl.tokens, pl.args);
        }
        else if (!strcmp(c, "ConcatTransform") || !strcmp(c, "Transform")) {
            int concat = (c[0] == 'C');
            farr_t a;
            if (read_num_array(lx, &a) && a.n == 16) {
                RtMatrix m; to_matrix(m, a.v);
                if (concat) RiConcatTransform(m); else RiTransform(m);
            } else {
                fprintf(stderr, "[rib_reader] %s: expected 16 numbers\n", concat ? "ConcatTransform" : "Transform");
            }
            free(a.v);
        }
        else if (!strcmp(c, "Identity"))       RiIdentity();
        else if (!strcmp(c, "Translate"))      { if (read_nums(lx, f, 3)) RiTranslate(f[0], f[1], f[2]); }
        else if (!strcmp(c, "Rotate"))         { if (read_nums(lx, f, 4)) RiRotate(f[0], f[1], f[2], f[3]); }
        else if (!strcmp(c, "Scale"))          { if (read_nums(lx, f, 3)) RiScale(f[0], f[1], f[2]); }
        else if (!strcmp(c, "WorldBegin"))     RiWorldBegin();
        else if (!strcmp(c, "WorldEnd"))       RiWorldEnd();
        else if (!strcmp(c, "AttributeBegin")) RiAttributeBegin();
        else if (!strcmp(c, "AttributeEnd"))   RiAttributeEnd();
        else if (!strcmp(c, "TransformBegin")) RiTransformBegin();
        else if (!strcmp(c, "TransformEnd"))   RiTransformEnd();
        else if (!strcmp(c, "FrameBegin"))     { if (read_nums(lx, f, 1)) RiFrameBegin((RtInt)f[0]); }
        else if (!strcmp(c, "FrameEnd"))       RiFrameEnd();
        else if (!strcmp(c, "ReadArchive"))    { lex_next(lx); strcpy(s1, lx->text); do_read_archive(s1); }
        else if (!strcmp(c, "PointsPolygons")) {
            farr_t a, b;
            int i, total = 0;
            read_num_array(lx, &a);
            read_num_array(lx, &b);
            read_param_list(lx, &pl);
            RtInt *nverts = (RtInt *)malloc(sizeof(RtInt) * (a.n ? a.n : 1));
            RtInt *verts  = (RtInt *)malloc(sizeof(RtInt) * (b.n ? b.n : 1));
            for (i = 0; i < a.n; i++) { nverts[i] = (RtInt)a.v[i]; total += nverts[i]; }
            for (i = 0; i < b.n; i++) verts[i] = (RtInt)b.v[i];
            if (total != b.n) {
                fprintf(stderr, "[rib_reader] PointsPolygons: expected %d indices, got %d\n", total, b.n);
            } else {
                RiPointsPolygonsV(a.n, nverts, verts, pl.n, pl.tokens, pl.args);
            }
            free(a.v); free(b.v);
        }
        else {
            fprintf(stderr, "[rib_reader] unknown RIB command: %s\n", c);
        }
    }
    fclose(lx->fp);
    free(lx);
    return 0;
}
