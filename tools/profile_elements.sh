#!/bin/bash
# One `ncu --set full` capture per element kernel (run under gpurun, 1 GPU). Output: gpurun_out/<tag>_<what>.ncu-rep
# Usage: tools/profile_elements.sh <tag> [what ...]      what = run_one.py names (default: the set below)
TAG=${1:-r01}; shift
WHATS=${@:-dilate exclusion chromahold lut4 remap gaussblur}
NCU=/usr/local/cuda/bin/ncu
mkdir -p gpurun_out
for W in $WHATS; do
  case $W in
    dilate) K=dilate_tma_kernel; SZ=4k;;
    exclusion|chromahold|lut4) K=stream_kernel; SZ=4k;;
    remap) K=remap4_kernel; SZ=8k;;
    remap_packed) K=remap4_packed_kernel; SZ=8k;;
    gaussblur) K=gaussblur_kernel; SZ=4k;;
    direct) K=bayer2rgb_direct; SZ=4k;;
    rgb2bayer) K=rgb2bayer_kernel; SZ=4k;;
    sad) K=sad_kernel; SZ=4k;;
    videodiff) K=videodiff_kernel; SZ=4k;;
    zebrastripe) K=zebra_planar_kernel; SZ=4k;;
    smooth) K=smooth_kernel; SZ=4k;;
    *) K=$W; SZ=4k;;
  esac
  timeout 300 $NCU --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_${W} \
      python tools/run_one.py $W $SZ > gpurun_out/${TAG}_${W}.stdout 2>&1
  tail -1 gpurun_out/${TAG}_${W}.stdout
done
ls -la gpurun_out | grep ${TAG}_
