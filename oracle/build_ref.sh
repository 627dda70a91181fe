#!/bin/bash
# TEST INFRASTRUCTURE — builds the *unmodified* lucille reference (read-only at
# /root/reference) into oracle/_ref/ so the restatement in oracle/ can be pinned
# against it and so bench.py has a CPU baseline of kind "reference".
#
# Nothing from /root/reference is copied into the repository: sources are compiled
# where they lie; only objects/archives/binaries (and the two example RIB inputs,
# which are measurement inputs, not sources) land in oracle/_ref/ (git-ignored,
# but shipped to the GPU box by gpurun like any other built artefact).
#
# The recipe mirrors the reference's own SConscripts (src/{base,ri,imageio,display,
# transport,render}/SConscript) with plain gcc; flags are the reference's own
# enable_64bit / enable_sse switches (SConstruct:98-112).  scons/flex/bison are
# absent in this image, so `lsh` itself is not built: oracle/ref/rib_reader.c feeds
# the same Ri*() calls the bison actions (src/lsh/parserib.y) would make.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LUCILLE_REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/render" ]; then
    echo "[build_ref] $REF not present: keeping prebuilt oracle/_ref as is"
    exit 0
fi
mkdir -p "$OUT/obj" "$OUT/obj_stat" "$OUT/scenes"

CFLAGS="-O2 -g -msse2 -m64 -std=gnu89 -fcommon -w -fPIC -D__64bit__ -D__x86__ -DLINUX -DWITH_PTHREAD -DWITH_SSE"
INC="-I$REF/src/base -I$REF/src/ri -I$REF/src/render -I$REF/src/transport -I$REF/src/imageio -I$REF/src/display -I$REF/include"

BASE="array dlload hash list log matrix memory parallel quaternion queue random stack thread timer util vector geometric system"
RI="apitable attribute backdoor camera context declare display lightsource option quadric ri subdivision transform"
IMAGEIO="rgbe image_loader image_saver"
DISPLAY="framebufferdrv hdrdrv openexrdrv sockdrv"
TRANSPORT="transport ambientocclusion dirtmap whitted"
RENDER="accel beam brdf bvh film filter geom hilbert hilbert2d ibl intersection_state light material mc noise polygon qmc raster ray raytrace reflection render scene shader shading specrend spectrum spiral sss subdivision sunsky texture texture_loader tonemap triangle ugrid zorder2d"

list=""
for d in base:"$BASE" ri:"$RI" imageio:"$IMAGEIO" display:"$DISPLAY" transport:"$TRANSPORT" render:"$RENDER"; do
    dir="${d%%:*}"; files="${d#*:}"
    for f in $files; do list="$list $dir/$f"; done
done

compile_one() {
    local rel="$1"; local dir="${rel%%/*}"; local f="${rel#*/}"
    local src="$REF/src/$dir/$f.c"; local obj="$OUT/obj/${dir}_$f.o"
    if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then
        gcc $CFLAGS $INC -c "$src" -o "$obj"
    fi
}
export -f compile_one; export REF OUT CFLAGS INC
echo $list | tr ' ' '\n' | grep -v '^$' | xargs -P 8 -I{} bash -c 'compile_one {}'

# second copy of bvh.c with the reference's own traversal counters switched on
# (bvh.h:39, counters at bvh.c:460,829-845,1129-1151) -> I and T of the roofline formula
gcc $CFLAGS $INC -DRI_BVH_TRACE_STATISTICS -c "$REF/src/render/bvh.c" -o "$OUT/obj_stat/render_bvh.o"

rm -f "$OUT"/libluciref_core.a "$OUT"/libluciref_core_stat.a
ar rcs "$OUT/libluciref_core.a" "$OUT"/obj/*.o
ar rcs "$OUT/libluciref_core_stat.a" $(ls "$OUT"/obj/*.o | grep -v render_bvh.o) "$OUT/obj_stat/render_bvh.o"

# my own drivers around the reference API (sources live in oracle/ref/, they are not reference code)
DRV="$HERE/ref"
gcc -O2 -g -std=gnu99 -fPIC -w -D__64bit__ -D__x86__ -DLINUX -DWITH_PTHREAD -DWITH_SSE $INC -shared \
    "$DRV/ref_shim.c" "$DRV/rib_reader.c" -Wl,--whole-archive "$OUT/libluciref_core.a" -Wl,--no-whole-archive \
    -lm -ldl -lpthread -o "$OUT/libluciref.so"
gcc -O2 -g -std=gnu99 -fPIC -w -D__64bit__ -D__x86__ -DLINUX -DWITH_PTHREAD -DWITH_SSE -DRI_BVH_TRACE_STATISTICS $INC -shared \
    "$DRV/ref_shim.c" "$DRV/rib_reader.c" -Wl,--whole-archive "$OUT/libluciref_core_stat.a" -Wl,--no-whole-archive \
    -lm -ldl -lpthread -o "$OUT/libluciref_stat.so"
gcc -O2 -g -std=gnu99 -w "$DRV/oracle_rib_main.c" -L"$OUT" -Wl,-rpath,'$ORIGIN' -lluciref -lm -ldl -lpthread -o "$OUT/oracle_rib"

# drop-in proof: the same front end + the UNMODIFIED reference renderer, with ri_accel_bind() interposed so that accel
# method 2 selects the GPU accelerator (integration/ri_b200_binding.c -> lucille_b200/libb200accel.so) and with the batched frame
# hook (integration/ri_b200_frame_hook.c) spliced in at ri_render_frame / ri_thread_create: `--accel b200` renders the frame in ONE
# ri_b200_render_ao call, `RI_B200_FRAME=0 ... --accel b200` through the per-ray vtable slot.  Needs the product library to exist.
if [ -f "$HERE/../lucille_b200/libb200accel.so" ]; then
    gcc -O2 -g -std=gnu99 -w -D__64bit__ -D__x86__ -DLINUX -DWITH_PTHREAD -DWITH_SSE -DLREF_WITH_B200 $INC -I"$HERE/../include" \
        "$DRV/oracle_rib_main.c" "$DRV/ref_shim.c" "$DRV/rib_reader.c" "$HERE/../integration/ri_b200_binding.c" \
        "$HERE/../integration/ri_b200_frame_hook.c" -I"$HERE/../integration" \
        -Wl,--wrap=ri_accel_bind -Wl,--wrap=ri_render_frame -Wl,--wrap=ri_thread_create -Wl,--whole-archive "$OUT/libluciref_core.a" -Wl,--no-whole-archive \
        -L"$HERE/../lucille_b200" -lb200accel -Wl,-rpath,'$ORIGIN/../../lucille_b200' -lm -ldl -lpthread -o "$OUT/lsh_b200"
fi

# measurement inputs (BASELINE.json configs[0], configs[3]); not sources
cp -f "$REF/examples/ambient_occlusion/ambient_occlusion.rib" "$OUT/scenes/"
rm -rf "$OUT/scenes/plane_sphere"; cp -r "$REF/examples/plane_sphere" "$OUT/scenes/plane_sphere"
echo "[build_ref] ok -> $OUT"
