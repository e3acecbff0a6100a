#!/bin/bash
# round 2 evidence set (after the fixed-point ring beam kernel and the ResNet-stem encoders): full GPU suite, smoke,
# both bench arms with the driver's arguments, one ncu capture of the beam ring kernel over q23 planes
O=gpurun_out; mkdir -p $O
timeout -k 10 1800 python -m pytest tests -q -m gpu > $O/r02t_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02t_pytest_gpu.log | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02t_bench.json 2> $O/r02t_bench.err; echo "bench exit $? in ${SECONDS}s"; tail -3 $O/r02t_bench.err; cut -c1-400 $O/r02t_bench.json
SECONDS=0; timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02t_bench_ref.json 2> $O/r02t_bench_ref.err; echo "reference exit $? in ${SECONDS}s"; cut -c1-300 $O/r02t_bench_ref.json
ND_MINLEN=99 timeout 300 python scripts/profile_step.py l2t 5 > $O/r02t_profile_l2t_5.txt 2>&1; tail -9 $O/r02t_profile_l2t_5.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn_ring -s 400 -c 1 -o $O/r02t_ring_q23 -f env ND_MINLEN=99 python scripts/profile_step.py l2t 5 > $O/ncu_t_ring.log 2>&1; echo "ncu exit $?"
ncu -i $O/r02t_ring_q23.ncu-rep --page raw --csv > $O/r02t_ring_q23.raw.csv 2>/dev/null; gzip -f $O/r02t_ring_q23.raw.csv; rm -f $O/r02t_ring_q23.ncu-rep
