#!/bin/bash
# A/B: 128-wide tiles for the N = 1536 decode projection at d = 512
O=gpurun_out; mkdir -p $O
for v in 0 1; do
ND_OPTS=gemm_wide_wave=$v ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 > $O/r02h_profile_t2t512_wide$v.txt 2>&1; echo "== gemm_wide_wave=$v"; cat $O/r02h_profile_t2t512_wide$v.txt
done
timeout 600 python -m pytest tests -q -m gpu -k "d512 or gemm" 2>&1 | tail -3
