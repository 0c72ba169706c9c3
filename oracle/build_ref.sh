#!/bin/bash
# TEST INFRASTRUCTURE: compiles, from /root/reference and in place, (1) the reference's own vendored ingest code (tinyobj, tinyexr +
# miniz) into oracle/_ref/libref_ingest.so, (2) its header-only math + src/bsdf/ggx.cpp against a scalar Enoki stand-in into
# oracle/_ref/libref_math.so and (3) its whole renderer (every src/**/*.cpp on the rendering path) against a CPU stand-in for Enoki and for
# the OptiX glue into oracle/_ref/libref_render.so (git-ignored; they travel to the GPU box with the snapshot). No reference source is
# copied. The renderer as shipped (Enoki + OptiX + CUDA) cannot be built here; see DESIGN.md §2.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PSDR_REFERENCE:-/root/reference}"
[ -d "$REF/include/tiny_obj_loader" ] || { echo "reference not present at $REF: keeping any prebuilt oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
g++ -O2 -std=c++17 -shared -fPIC -w -I"$REF/include" "$HERE/ref_ingest_shim.cpp" "$REF/src/core/miniz.cpp" -o "$HERE/_ref/libref_ingest.so"
echo "built $HERE/_ref/libref_ingest.so"
# the reference's own header-only math + src/bsdf/ggx.cpp, compiled unmodified against the scalar Enoki stand-in (oracle/ref_stub)
g++ -O1 -std=c++17 -ffp-contract=off -shared -fPIC -w -I"$HERE/ref_stub" -I"$REF/include" -I"$REF" "$HERE/ref_math_shim.cpp" "$REF/src/core/miniz.cpp" -o "$HERE/_ref/libref_math.so"
echo "built $HERE/_ref/libref_math.so"
# the reference's renderer, compiled unmodified against oracle/ref_dyn (CPU stand-in for Enoki) + oracle/ref_render_shim.cpp
bash "$HERE/build_ref_render.sh"
