#!/bin/bash
# round 2 final evidence set: full GPU suite, smoke, both bench arms with the driver's arguments
O=gpurun_out; mkdir -p $O
timeout -k 10 1800 python -m pytest tests -q -m gpu > $O/r02y_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02y_pytest_gpu.log | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02y_bench.json 2> $O/r02y_bench.err; echo "bench exit $? in ${SECONDS}s"; tail -3 $O/r02y_bench.err; cut -c1-400 $O/r02y_bench.json
SECONDS=0; timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02y_bench_ref.json 2> $O/r02y_bench_ref.err; echo "reference exit $? in ${SECONDS}s"; cut -c1-300 $O/r02y_bench_ref.json
