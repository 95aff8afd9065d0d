#!/bin/sh
# Builds libgsplat_b200.so (sm_100a only) in-tree.  Usage: sh build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
OBJS=""
for f in gs_capi gs_preprocess gs_binning gs_composite_fwd gs_composite_bwd gs_preprocess_bwd gs_metrics gs_adapter gs_cameras; do
  stale=0
  for dep in "$f.cu" gs_common.cuh gs_tile_sort.cuh ../../include/gsplat_b200.h build.sh; do
    if [ ! -f "$f.o" ] || [ "$dep" -nt "$f.o" ]; then stale=1; fi
  done
  if [ "$stale" = 1 ]; then
    $NVCC $FLAGS "$@" -c "$f.cu" -o "$f.o" &
  fi
  OBJS="$OBJS $f.o"
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libgsplat_b200.so $OBJS -lcudart
echo "built $(pwd)/libgsplat_b200.so"
