#!/bin/bash
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "greedy_matches or beam_matches" 2>&1 | tail -2
timeout 300 python scripts/profile_step.py l2t 1 2>&1 | sed -n 3,6p
timeout 300 python scripts/profile_step.py nano2rnn 1 2>&1 | sed -n 3,8p
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 1024 2>&1 | sed -n 3,6p
