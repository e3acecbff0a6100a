#!/bin/bash
export ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)"
python scripts/profile_step.py t2t 1 512 > gpurun_out/plain_c5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn -s 620 -c 2 -o gpurun_out/r01c_prof_cross_d512 -f python scripts/profile_step.py t2t 1 512 > gpurun_out/ncu_c5.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_c5.log
