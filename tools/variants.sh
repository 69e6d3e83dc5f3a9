#!/bin/bash
# tuning helper (GPU box): short bench for each experimental build variants_tmp_*.so
for so in variants_tmp_*.so; do
  for ph in ring point; do
    ORT_LIB=$PWD/$so python bench.py --phase $ph --rays 2147483648 --steps 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$so', '$ph', '%.3e rays/s' % d['value'], 'frac %.3f' % d['roofline']['frac'])"
  done
done
