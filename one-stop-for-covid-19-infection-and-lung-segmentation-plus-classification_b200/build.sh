#!/bin/bash
# Builds libb200unet.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
set -o pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
mkdir -p build
pids=()
for f in csrc/*.cu; do
  o=build/$(basename "${f%.cu}").o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ csrc/common.cuh -nt "$o" ] || [ csrc/launch.cuh -nt "$o" ] || [ csrc/internal.h -nt "$o" ] || [ ../include/b200unet.h -nt "$o" ] || [ csrc/tc_common.cuh -nt "$o" ]; then
    # preprocess.cu restates OpenCV's CPU arithmetic bit for bit: no fused multiply-add contraction there (the CV_64F
    # resize compares equal to cv2 only with separately rounded products and sums)
    EXTRA=""
    [ "$(basename "$f")" = "preprocess.cu" ] && EXTRA="--fmad=false"
    $NVCC $FLAGS $EXTRA ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o libb200unet.so build/*.o -ldl
echo "built $(pwd)/libb200unet.so"
