/*
 * ri_b200_binding.h -- what the lucille-side binding keeps behind ri_accel_t.data (shared by ri_b200_binding.c, the per-ray vtable
 * slots, and ri_b200_frame_hook.c, the batched frame).
 */
#ifndef RI_B200_BINDING_H
#define RI_B200_BINDING_H

#include <stdint.h>

#include "geom.h"
#include "lucille_b200.h"

#ifndef RI_ACCEL_B200
#define RI_ACCEL_B200 2              /* next to RI_ACCEL_UGRID 0 and RI_ACCEL_BVH 1 (src/render/accel.h:20-21) */
#endif

typedef struct {
    ri_b200_accel_t *dev;
    uint64_t         ntris;
    uint32_t        *orig;          /* post-build position -> flattened input triangle (ri_b200_triorder) */
    ri_geom_t      **tri_geom;      /* flattened input triangle -> owning geom   (ri_triangle_t.geom,  bvh.c:1812) */
    uint32_t        *tri_index;     /* flattened input triangle -> 3 * i         (ri_triangle_t.index, bvh.c:1813) */
} b200_binding_t;

#endif
