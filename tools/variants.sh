#!/bin/bash
# tuning helper (GPU box): short bench for each experimental build variants_tmp_*.so
# usage: tools/variants.sh [phases]   (default "ring point")
PHASES=${1:-ring point}
for so in variants_tmp_*.so; do
  for ph in $PHASES; do
    ORT_LIB=$PWD/$so python bench.py --phase $ph --rays 4294967296 --steps 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$so', '$ph', '%.4e rays/s' % d['value'], 'frac %.3f' % d['roofline']['frac'])"
  done
done
