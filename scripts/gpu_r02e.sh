#!/bin/bash
# round 2, fifth call: split-tail cross attention, new goldens (rnn2rnn H=256), C5 / C2 step profiles, ncu captures
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/kv_modes.py 0,3,4 1,3 > $O/r02e_kv_modes.txt 2>&1; echo "kv_modes exit $?"; cat $O/r02e_kv_modes.txt
timeout -k 10 1500 python -m pytest tests -q -m gpu -k "full_batch or fixed_point or rnn2rnn or pooling" > $O/r02e_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02e_pytest_gpu.log | tail -30
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 > $O/r02e_profile_t2t512_1.txt 2>&1; cat $O/r02e_profile_t2t512_1.txt
timeout 600 python scripts/profile_step.py l2t 1 > $O/r02e_profile_l2t_1.txt 2>&1; cat $O/r02e_profile_l2t_1.txt
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 900 ncu --set full --clock-control none --import-source on -k regex:cross_attn_packed -s 1210 -c 2 -o $O/r02e_prof_cross_q23_d512 -f python scripts/profile_step.py t2t 1 > $O/ncu_e1.log 2>&1; echo "ncu d512 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cross_attn_packed -s 310 -c 2 -o $O/r02e_prof_cross_q23_d256 -f python scripts/profile_step.py l2t 1 > $O/ncu_e2.log 2>&1; echo "ncu d256 exit $?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file $O/r02e_launches_bench_c5.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_e3.log 2>&1; echo "ncu launch list exit $?"; tail -2 $O/ncu_e3.log | cut -c1-300
gzip -f $O/r02e_launches_bench_c5.csv; ls -la $O | grep r02e
