#!/bin/bash
mkdir -p gpurun_out/workloads
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "beam" 2>&1 | tail -2
timeout 300 python scripts/profile_step.py l2t 5 2>&1 | sed -n 3,3p
for w in l2t_beam5_b1024 t2t_greedy_b1024 t2t512_greedy_b1024 nano2rnn_greedy_b1024 brnn2rnn_greedy_b1024 cnn2cnn_greedy_b1024; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/workloads/$w.json 2> gpurun_out/workloads/$w.err
  echo "$w exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/workloads/$w.json"))
print("  %.1f ms/step, %.0f chunks/s, %.0f bases/s, e2e %.0f, roofline %s %.0f GB/s (%.3f)" % (d["ms_per_step"], d["chunks_per_s"], d["value"], d["e2e"]["value"], d["roofline"]["kernel"].split(" ")[0], d["roofline"]["achieved"], d["roofline"]["frac"]))
PY
done
