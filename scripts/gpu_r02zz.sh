#!/bin/bash
# round 2, last evidence set on the final code: full GPU suite, smoke, both bench arms with the driver's arguments, then the
# ncu launch list of the bench command (the same command having exited 0 without ncu first)
O=gpurun_out; mkdir -p $O
timeout -k 10 1800 python -m pytest tests -q -m gpu > $O/r02zz_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02zz_pytest_gpu.log | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02zz_bench.json 2> $O/r02zz_bench.err; echo "bench exit $? in ${SECONDS}s"; tail -3 $O/r02zz_bench.err; cut -c1-300 $O/r02zz_bench.json
SECONDS=0; timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02zz_bench_ref.json 2> $O/r02zz_bench_ref.err; echo "reference exit $? in ${SECONDS}s"; cut -c1-200 $O/r02zz_bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r02zz_bench_short.json 2>/dev/null; echo "short bench exit $?"
SECONDS=0; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file $O/r02zz_launches_bench_c5.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_zz.log 2>&1; echo "ncu launch list exit $? in ${SECONDS}s"
gzip -f $O/r02zz_launches_bench_c5.csv; ls -la $O | grep r02zz
