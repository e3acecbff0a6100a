#!/bin/bash
# round 2, fourth call: slice variants of the fixed-point cross attention, GPU tests, bench, compute-sanitizer
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/kv_modes.py 0,3,4 1,3 > $O/r02d_kv_modes.txt 2>&1; echo "kv_modes exit $?"; cat $O/r02d_kv_modes.txt
timeout -k 10 1500 python -m pytest tests -q -m gpu > $O/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02d_pytest_gpu.log | tail -30
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/r02d_bench.json 2> $O/r02d_bench.err; echo "bench exit $?"; tail -3 $O/r02d_bench.err; cut -c1-1500 $O/r02d_bench.json
bash scripts/gpu_sanitizer.sh r02d
