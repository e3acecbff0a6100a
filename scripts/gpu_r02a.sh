#!/bin/bash
# round 2, first call: GPU tests, bench (C5 default + extras), greedy identity rates (fp32 / q24 / q16 memory K/V)
O=gpurun_out; mkdir -p $O
timeout -k 10 1200 python -m pytest tests -q -m gpu -s > $O/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|kv_mode" $O/r02a_pytest_gpu.log | tail -30
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench exit $?"; tail -3 $O/r02a_bench.err; cat $O/r02a_bench.json
timeout 1500 python scripts/identity_rates.py --only greedy --out $O/r02a_identity_greedy.json > $O/r02a_identity_greedy.log 2>&1; echo "identity exit $?"
cut -c1-400 $O/r02a_identity_greedy.log | tail -20
