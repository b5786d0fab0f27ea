#!/bin/bash
# Builds oracle/_ref/parsy_ref from the reference sources where they lie under /root/reference.
# TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing from the reference is copied into the repository:
# the headers are staged in a throw-away temp dir only to add the eight missing `return` statements
# (GCC 13 turns the fall-through into UB) and the temp dir is deleted again.
# Output: oracle/_ref/parsy_ref (git-ignored, travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PARSY_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/cholesky" ] || { echo "reference not present at $REF; keeping prebuilt $OUT" >&2; exit 0; }
METIS="${PARSY_METIS:-/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a}"
BLAS="${PARSY_OPENBLAS:-$(python3 - <<'PY'
import glob, sysconfig
print(sorted(glob.glob(sysconfig.get_paths()["purelib"] + "/opencv_python_headless.libs/libopenblasp-*.so"))[0])
PY
)}"
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/cholesky" "$REF/common" "$REF/triangularSolve" "$TMP/"
patch_ret() { sed -i "$2s/^}/ return 1; }/" "$TMP/$1"; }
patch_ret common/def.h 236; patch_ret common/def.h 257; patch_ret common/Util.h 223
patch_ret common/TreeUtils.h 471; patch_ret common/BFS.h 52; patch_ret common/BFS.h 91
patch_ret cholesky/Inspection_Prune.h 22; patch_ret cholesky/Inspection_Block.h 131
g++ -O2 -DNDEBUG -std=c++11 -fpermissive -w -fopenmp -DMKL -DMETIS \
    -I"$HERE/shim" -I"$TMP/cholesky" -I"$TMP/common" -I"$TMP/triangularSolve" \
    "$HERE/ref_driver.cpp" "$METIS" "$BLAS" -Wl,--disable-new-dtags,-rpath,"$(dirname "$BLAS")" -o "$OUT/parsy_ref"
echo "built $OUT/parsy_ref"
# The same driver with the call sites forwarded to the CUDA executor through include/parsy_cuda_dropin.h (tests only:
# the reference's own inspector and harness drive libparsy_cuda); needs the library built first.
LIBDIR="$HERE/../parsy_bench_b200"
if [ -f "$LIBDIR/libparsy_cuda.so" ]; then
  g++ -O2 -DNDEBUG -std=c++11 -fpermissive -w -fopenmp -DMKL -DMETIS -DPARSY_GPU_FORWARD \
      -I"$HERE/shim" -I"$TMP/cholesky" -I"$TMP/common" -I"$TMP/triangularSolve" -I"$HERE/../include" \
      "$HERE/ref_driver.cpp" "$METIS" "$BLAS" -L"$LIBDIR" -l:libparsy_cuda.so -L/usr/local/cuda/lib64 -lcudart \
      -Wl,--disable-new-dtags,-rpath,"$(dirname "$BLAS")",-rpath,'$ORIGIN/../../parsy_bench_b200',-rpath,/usr/local/cuda/lib64 \
      -o "$OUT/parsy_ref_gpu"
  echo "built $OUT/parsy_ref_gpu"
fi
# examples/MakingLowerHalf.cpp is a self-contained program (std headers only): compiled where it lies, used by
# tests/test_mmio.py to pin parsy_make_lower_half byte for byte.
g++ -O2 -w "$REF/examples/MakingLowerHalf.cpp" -o "$OUT/making_lower_half"
echo "built $OUT/making_lower_half"
