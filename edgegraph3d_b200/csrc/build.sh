#!/bin/bash
# Builds edgegraph3d_b200/libeg3d.so for sm_100a (B200).  -fmad=false: float/double expressions must round exactly as
# written (threshold parity with the reference's non-FMA x86-64 build); IEEE division / sqrt are nvcc defaults.
set -e
# OUT=<path> and extra flags (e.g. -DEG3D_K3_PROFILE for the per-phase cycle counters) give A/B builds; lib.py loads $EG3D_LIB.
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -Xcompiler -O3 -Xcompiler -ffp-contract=off -shared -o "${OUT:-../libeg3d.so}" eg3d_capi.cu eg3d_plg_build.cpp eg3d_fundamental.cpp eg3d_sfm_io.cpp -ldl -lpthread "$@"
