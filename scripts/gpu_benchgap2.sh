#!/bin/bash
for cfg in "none 0.05" "nvml 0.05" "nvml 0.25" "smi 0.2"; do
  set -- $cfg
  echo "== sampler=$1 interval=$2"
  ND_BENCH_SAMPLER=$1 ND_BENCH_SAMPLE_S=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'], d['e2e']['value'])"
done
