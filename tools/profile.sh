#!/bin/bash
# Run under gpurun (1 GPU). Produces in gpurun_out/:
#   launches_<tag>.csv   every kernel launch of a short bench run with its device time (shares, not absolutes)
#   <tag>_<kernel>.ncu-rep  one `--set full` capture per listed kernel regex
# Usage: tools/profile.sh <tag> <kernel-regex> [<kernel-regex> ...]
set -u
TAG=${1:-r01}; shift
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --frames 24 --no-elements > gpurun_out/launches_${TAG}.stdout 2>&1
for K in "$@"; do
  $NCU --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_${K} \
      python bench.py --steps 2 --warmup 3 --frames 24 --profile > gpurun_out/${TAG}_${K}.stdout 2>&1
done
ls -la gpurun_out
