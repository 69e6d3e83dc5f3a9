#!/bin/bash
# Run on the GPU box (via gpurun): plain bench lines first, then the ncu launch list and one
# full capture of the trace megakernel per phase.  Outputs land in gpurun_out/.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SMALL="--steps 2 --warmup 3 --rays 134217728 --no-cpu"
for PH in ring point; do
  python bench.py --phase $PH $SMALL > $OUT/plain_${PH}_${TAG}.log 2>&1 || { echo "plain $PH failed"; tail -5 $OUT/plain_${PH}_${TAG}.log; exit 1; }
done
for PH in ring point; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/launches_${PH}_${TAG}.csv python bench.py --phase $PH $SMALL > $OUT/ncu_launch_${PH}_${TAG}.log 2>&1
  ncu --set full --clock-control none --import-source on -k 'regex:ort_(trace_|ring_cull)' -s 3 -c 1 \
      -f -o $OUT/prof_${PH}_${TAG} python bench.py --phase $PH $SMALL > $OUT/ncu_full_${PH}_${TAG}.log 2>&1
done
ls -la $OUT
