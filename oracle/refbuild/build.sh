#!/bin/bash
# Recipe for oracle/_ref: the reference's own hot path compiled from the sources where they lie.
# TEST INFRASTRUCTURE ONLY.  Outputs go to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
#
#   libivslam_ref.so        reference translation units built as the reference builds them
#                           (introspective_ORB_SLAM/CMakeLists.txt:16-17: -O3 -march=native, Release => -DNDEBUG,
#                           GCC's default -ffp-contract=fast, so a*b+c is FMA-contracted: SURVEY Q9)
#   libivslam_ref_nofma.so  the same sources with -ffp-contract=off (the canonical float semantics of the oracle)
#
# -march=native is replaced by -march=x86-64-v3 (AVX2 + FMA): the library is built in this container and executed on the
# GPU box, whose host CPU may lack this machine's AVX-512; float results depend on FMA being available, not on the vector
# width (no -ffast-math, so vectorisation never re-associates).
# The OpenCV-compat layer and the cv2-pinned pixel primitives (oracle/ivslam_oracle.cpp) are always built with
# -ffp-contract=off: in the real system they live inside libopencv, which is not built with the reference's flags.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${IVSLAM_REFERENCE:-/root/reference/introspective_ORB_SLAM}"
OUT="$HERE/../_ref"
CXX="${CXX:-g++}"
if [ ! -f "$REF/src/ORBextractor.cc" ]; then
  echo "reference sources not found at $REF: keeping whatever is already in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/gen" "$OUT/obj"
python3 "$HERE/extract_reference.py" "$REF" "$OUT/gen"

COMMON="-std=c++17 -fPIC -pthread -I$HERE/cvcompat -I$REF/include -I$OUT -Wno-unused-variable -Wno-sign-compare"
ARCH="-O3 -march=x86-64-v3 -DNDEBUG"
# cv2-pinned primitives + compat layer: canonical float semantics
$CXX $COMMON $ARCH -ffp-contract=off -c "$HERE/../ivslam_oracle.cpp" -o "$OUT/obj/oracle.o"
$CXX $COMMON $ARCH -ffp-contract=off -c "$HERE/cvcompat_impl.cpp" -o "$OUT/obj/cvcompat.o"
for variant in asbuilt nofma; do
  if [ $variant = asbuilt ]; then FP="-ffp-contract=fast -DREF_FP_CONTRACT=1"; SO=libivslam_ref.so; else FP="-ffp-contract=off -DREF_FP_CONTRACT=0"; SO=libivslam_ref_nofma.so; fi
  $CXX $COMMON $ARCH $FP -c "$REF/src/ORBextractor.cc" -o "$OUT/obj/ORBextractor_$variant.o"
  $CXX $COMMON $ARCH $FP -c "$HERE/ref_capi.cpp" -o "$OUT/obj/ref_capi_$variant.o"
  $CXX -shared -pthread -Wl,-Bsymbolic -o "$OUT/$SO.tmp" "$OUT/obj/ORBextractor_$variant.o" "$OUT/obj/ref_capi_$variant.o" "$OUT/obj/cvcompat.o" "$OUT/obj/oracle.o"
  mv "$OUT/$SO.tmp" "$OUT/$SO"
done
sha256sum "$REF/src/ORBextractor.cc" "$REF/src/Frame.cc" "$REF/src/ORBmatcher.cc" "$REF/include/ORBextractor.h" > "$OUT/SOURCES.sha256"
echo "built $OUT/libivslam_ref.so and $OUT/libivslam_ref_nofma.so"
