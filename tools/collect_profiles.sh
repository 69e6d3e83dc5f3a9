#!/bin/bash
# Condense what a tools/gpu_profile.sh run left in gpurun_out/ into the tracked summaries under profiles/.
# usage: tools/collect_profiles.sh <capture tag in gpurun_out, e.g. r02h> <name under profiles/, e.g. r02>
set -eu
T=$1; N=$2
here="$(cd "$(dirname "$0")/.." && pwd)"; cd "$here"
mkdir -p /tmp/ort_cub && (cd /tmp/ort_cub && rm -f *.cubin && cuobjdump -xelf all "$here/opticalraytrace_b200/libort.so" > /dev/null)
for PH in ring point; do
  python tools/ncu_summary.py gpurun_out/prof_${PH}_${T}.ncu-rep > profiles/${N}_${PH}_full.txt
  cp gpurun_out/launches_${PH}_${T}.csv profiles/${N}_${PH}_launches.csv
done
# SASS opcode-class histograms: static, executed (dynamic) and per source line
K_POINT=ort_trace_kernelILi2ELi1ELi0EdE; K_RING=ort_ring_cull_kernelILb0
{ python tools/sass_hist.py static opticalraytrace_b200/libort.so $K_POINT; echo; python tools/sass_hist.py dynamic gpurun_out/prof_point_${T}.ncu-rep; echo;
  python tools/sass_hist.py lines gpurun_out/prof_point_${T}.ncu-rep /tmp/ort_cub/ort_cuda.sm_100a.cubin $K_POINT; } > profiles/${N}_point_sass_hist.txt
{ python tools/sass_hist.py static opticalraytrace_b200/libort.so $K_RING; echo; python tools/sass_hist.py dynamic gpurun_out/prof_ring_${T}.ncu-rep; echo;
  python tools/sass_hist.py lines gpurun_out/prof_ring_${T}.ncu-rep /tmp/ort_cub/ort_cuda.sm_100a.cubin $K_RING; } > profiles/${N}_ring_sass_hist.txt
ls -la profiles | tail -12
