#!/bin/bash
# TEST INFRASTRUCTURE: compiles psdr-cuda's own renderer sources, in place from /root/reference, against oracle/ref_dyn (CPU stand-in for
# Enoki) + oracle/ref_render_shim.cpp (stand-in for the OptiX glue, C entry points) into oracle/_ref/libref_render.so. Called by build_ref.sh.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PSDR_REFERENCE:-/root/reference}"
OBJ="$HERE/_ref/obj"
mkdir -p "$OBJ"
FLAGS="-O1 -std=c++17 -ffp-contract=off -fPIC -w -fopenmp -I$HERE/ref_dyn -I$REF/include"
pids=()
for f in bsdf/diffuse bsdf/ggx bsdf/roughconductor core/bitmap core/bitmap_loader core/cube_distrb core/pmf core/sampler emitter/area emitter/envmap \
         integrator/direct integrator/field integrator/integrator scene/scene scene/scene_loader sensor/perspective sensor/sensor shape/mesh; do
    g++ $FLAGS -c "$REF/src/$f.cpp" -o "$OBJ/$(basename $f).o" & pids+=($!)
done
g++ $FLAGS -c "$HERE/ref_render_shim.cpp" -o "$OBJ/shim.o" & pids+=($!)
g++ -O1 -fPIC -w -I"$REF/include" -c "$REF/src/core/miniz.cpp" -o "$OBJ/miniz.o" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
g++ -shared -fopenmp -o "$HERE/_ref/libref_render.so" "$OBJ"/*.o
rm -rf "$OBJ"
echo "built $HERE/_ref/libref_render.so"
