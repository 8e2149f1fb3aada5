#!/bin/bash
# Builds oracle/_ref/ — the reference's OWN sources (/root/reference/src/nanogi.cpp, src/tinyexr.cc, include/nanogi/*.hpp),
# compiled where they lie against the stand-in third-party headers of oracle/refshim/ (see oracle/refshim/README.md).
# TEST INFRASTRUCTURE: the result pins the CPU oracle; nothing in the product path loads it.
#
# One transformation is applied on the way to the compiler, into a temporary directory that is deleted afterwards: lines
# consisting of `#pragma region ...` / `#pragma endregion` are blanked (line numbers are preserved). GCC >= 13 parses these
# pragmas as statements, and the reference puts them between `}` and `else` (e.g. include/nanogi/rt.hpp:1836-1842), which
# GCC 13 rejects with "'else' without a previous 'if'"; the reference's own toolchain (GCC 4.8 / MSVC) ignores them.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NGI_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/nanogi.cpp" ]; then echo "build_ref.sh: $REF not present (GPU box?): keeping the prebuilt oracle/_ref"; exit 0; fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/nanogi_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/include/nanogi" "$TMP/src"
strip_pragmas() { sed -E 's/^[[:space:]]*#pragma (end)?region.*$//' "$1" > "$2"; }
for f in "$REF"/include/nanogi/*.hpp "$REF"/include/nanogi/*.h; do strip_pragmas "$f" "$TMP/include/nanogi/$(basename "$f")"; done
strip_pragmas "$REF/src/nanogi.cpp" "$TMP/src/nanogi.cpp"
cp "$REF/src/tinyexr.cc" "$TMP/src/tinyexr.cc"
CXXFLAGS="-std=c++14 -O2 -DNDEBUG -fPIC -ffp-contract=off -mfma -w -I$HERE/refshim -I$TMP/include"
g++ $CXXFLAGS -c "$TMP/src/tinyexr.cc" -o "$TMP/tinyexr.o"
# (1) the reference application itself: src/nanogi.cpp's own main()
g++ $CXXFLAGS "$TMP/src/nanogi.cpp" "$TMP/tinyexr.o" -o "$OUT/nanogi_ref" -lz -pthread
# (2) the same translation unit behind a C interface for the tests (oracle/ref_harness.cpp includes src/nanogi.cpp)
g++ $CXXFLAGS -shared -Wl,-Bsymbolic-functions -DNGI_REF_SRC="\"$TMP/src/nanogi.cpp\"" "$HERE/ref_harness.cpp" "$TMP/tinyexr.o" -o "$OUT/libnanogi_ref.so" -lz -pthread
# (3) the reference application with the INTEGRATION.md B bridge compiled in (oracle/ref_gpu_bridge.hpp): a second temporary copy
#     of src/nanogi.cpp gets the bridge's #include after `using namespace nanogi;` and ONE statement in front of the
#     `switch (Type)` of Renderer::Render (src/nanogi.cpp:203). Linked against the product library; skipped when that is not built.
GPU_LIB_DIR="$HERE/../nanogi_b200"
if [ -f "$GPU_LIB_DIR/libnanogi_gpu.so" ]; then
  sed -E -e '0,/^using namespace nanogi;/s//using namespace nanogi;\n#include "ref_gpu_bridge.hpp"/' \
         -e '0,/switch \(Type\)/s//if (nanogi_gpu_bridge::Selected()) { nanogi_gpu_bridge::RenderOnGpu(scene, (int)Type, Params.NumSamples, Params.MaxNumVertices, Params.Width, Params.Height, film); NGI_LOG_INFO("Elapesed time: " + std::to_string((double)(std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - start).count()) \/ 1000.0)); return; }\n\t\t\tswitch (Type)/' \
         "$TMP/src/nanogi.cpp" > "$TMP/src/nanogi_gpu.cpp"
  grep -q "nanogi_gpu_bridge::RenderOnGpu" "$TMP/src/nanogi_gpu.cpp" || { echo "build_ref.sh: could not insert the bridge call"; exit 1; }
  g++ $CXXFLAGS -I"$HERE" -I"$HERE/../include" "$TMP/src/nanogi_gpu.cpp" "$TMP/tinyexr.o" -o "$OUT/nanogi_ref_gpu" -L"$GPU_LIB_DIR" -lnanogi_gpu -lz -pthread \
      -Wl,-rpath,'$ORIGIN/../../nanogi_b200'
  echo "built $OUT/nanogi_ref_gpu (reference Run / Scene::Load / SaveImage + RenderOnGpu -> libnanogi_gpu.so)"
fi
echo "built $OUT/nanogi_ref and $OUT/libnanogi_ref.so from $REF"
