#!/bin/bash
echo "=== C5 t2t d=512 6+6 B=1024 greedy"
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 1024 2>&1
echo "=== C3 l2t beam5"
timeout 300 python scripts/profile_step.py l2t 5 2>&1 | head -4
echo "=== t2t d=256"
timeout 300 python scripts/profile_step.py t2t 1 2>&1 | head -12
