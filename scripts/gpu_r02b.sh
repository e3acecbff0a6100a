#!/bin/bash
# round 2, second call: K/V storage formats (time + error), ncu of three of them, greedy identity rates
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/kv_modes.py > $O/r02b_kv_modes.txt 2>&1; echo "kv_modes exit $?"; cat $O/r02b_kv_modes.txt
for m in 1 3 5; do
ND_OPTS=kv_mode=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn_packed -s 310 -c 2 -o $O/r02b_prof_cross_kv$m -f python scripts/profile_step.py l2t 1 > $O/ncu_kv$m.log 2>&1; echo "ncu kv$m exit $?"
done
timeout 1500 python scripts/identity_rates.py --only greedy --greedy-opts "kv_mode=0;kv_mode=3;kv_mode=5;kv_mode=4" --out $O/r02b_identity_greedy.json > $O/r02b_identity_greedy.log 2>&1; echo "identity exit $?"
cut -c1-330 $O/r02b_identity_greedy.log | tail -30
