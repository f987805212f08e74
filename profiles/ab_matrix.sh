#!/bin/bash
# A/B harness for kernel variants (round-2 aid; not a bench).  Builds one library per variant HERE (nvcc cross-compiles
# without a GPU), then ONE gpurun call times them all on the C2 batch and prints, per variant, the kernel times and a
# checksum of the result — a variant only counts if its checksum equals the baseline's.
#
#   bash profiles/ab_matrix.sh build   "base:" "mb6:-DEG3D_K3B_MIN_BLOCKS=6" "gn2:-DEG3D_GN_UNROLL=2" "dltcv:-DEG3D_DLT_OPENCV=1"   # here
#   gpurun --timeout 600 -- 'bash profiles/ab_matrix.sh run > gpurun_out/ab_matrix.log 2>&1'             # on the B200
#
# Variant libraries live in edgegraph3d_b200/variants/ (git-ignored like every .so; they travel with the gpurun snapshot).
set -e
cd "$(dirname "$0")/.."
mkdir -p edgegraph3d_b200/variants
case "$1" in
  build)
    shift
    for spec in "$@"; do
      tag=${spec%%:*}; flags=${spec#*:}
      echo "building $tag ($flags)"
      OUT=$PWD/edgegraph3d_b200/variants/libeg3d_$tag.so bash edgegraph3d_b200/csrc/build.sh $flags
    done ;;
  run)
    for lib in edgegraph3d_b200/variants/libeg3d_*.so; do
      EG3D_LIB=$PWD/$lib python profiles/ab_compare.py "${REPS:-3}"
    done ;;
  *) echo "usage: $0 build tag:flags... | run"; exit 2 ;;
esac
