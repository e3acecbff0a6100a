#!/bin/bash
# round 2, sixth call: vectorised K/V packer, tail CTAs with 16 rows in flight, self attention with 8 loads in flight,
# fp64 front end; full GPU suite, step profiles, bench
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/kv_modes.py 0,3,4 1,3 > $O/r02f_kv_modes.txt 2>&1; echo "kv_modes exit $?"; cat $O/r02f_kv_modes.txt
timeout -k 10 1500 python -m pytest tests -q -m gpu > $O/r02f_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02f_pytest_gpu.log | tail -30
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 > $O/r02f_profile_t2t512_1.txt 2>&1; cat $O/r02f_profile_t2t512_1.txt
timeout 600 python scripts/profile_step.py l2t 1 > $O/r02f_profile_l2t_1.txt 2>&1; cat $O/r02f_profile_l2t_1.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench exit $?"; tail -3 $O/r02f_bench.err; cut -c1-700 $O/r02f_bench.json
