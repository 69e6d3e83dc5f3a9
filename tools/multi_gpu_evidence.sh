#!/bin/bash
# Run on an N-GPU box (gpurun --gpus N): the multi-process bench line (torchrun, one rank per GPU), the
# single-process one (ort_init(N)), the multi-device parity test and the reference's launcher interface.
# usage: tools/multi_gpu_evidence.sh N <tag>      -> gpurun_out/*_<tag>.{json,log}
set -u
N=$1; T=$2; OUT=gpurun_out; mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N > $OUT/bench_${N}gpu_$T.json 2> $OUT/bench_${N}gpu_$T.err
python bench.py --single-process --gpus $N > $OUT/bench_single_process_${N}gpu_$T.json 2> $OUT/bench_single_process_${N}gpu_$T.err
{
  echo "== pytest tests/test_gpu_multi.py ($N devices visible)"
  python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
  echo "== ./install.sh -n $N -f settings-config2.params"
  ./install.sh -n $N -f settings-config2.params 2>&1 | grep -v "^\s*$" | tail -8
  if [ "$N" = 2 ]; then
    echo "== ./install.sh -n 3 -f settings-config2.params  (more devices asked for than visible)"
    ./install.sh -n 3 -f settings-config2.params 2>&1 | grep -v "^\s*$" | tail -5
  fi
} > $OUT/multi_${N}gpu_$T.log 2>&1
python - <<PY
import json
for f in ("$OUT/bench_${N}gpu_$T.json", "$OUT/bench_single_process_${N}gpu_$T.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "%.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], {k: ("%.4g" % v["value"] if isinstance(v, dict) and "value" in v else v) for k, v in d.get("extra", {}).items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -20 $OUT/multi_${N}gpu_$T.log
