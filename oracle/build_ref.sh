#!/usr/bin/env bash
# Test infrastructure only: compiles the UNMODIFIED reference sources where they lie
# (/root/reference/src) into oracle/_ref/libproxtv_ref.so.  Nothing is copied into the repo;
# the output directory is git-ignored (but travels to the GPU box with gpurun).
# LAPACK symbols used by the out-of-scope PN/L2/Lp solvers are redirected to scipy's bundled OpenBLAS
# (SURVEY.md §8c).  Skips silently when /root/reference is absent (GPU box): the prebuilt .so is used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${PROXTV_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$SRC" ]; then
  echo "build_ref: $SRC not present; keeping prebuilt $OUT/libproxtv_ref.so (if any)"; exit 0
fi
OB=$(python3 -c "import scipy,os;print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)),'scipy.libs'))")
OBLIB=$(basename "$OB"/libscipy_openblas-*.so)
CXX=${CXX_REF:-/usr/bin/g++}
$CXX -O3 -fopenmp -fPIC -shared -DNOMATLAB=1 -Ddpttrf_=scipy_dpttrf_ -Ddpttrs_=scipy_dpttrs_ \
  -I"$SRC" "$SRC"/*.cpp -o "$OUT/libproxtv_ref.so" -L"$OB" -l:"$OBLIB" -Wl,-rpath,"$OB" 2> "$OUT/build.log" \
  || { cat "$OUT/build.log"; exit 1; }
echo "build_ref: built $OUT/libproxtv_ref.so"
