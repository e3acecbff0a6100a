#!/bin/bash
# self attention templated on the head size: parity subset + step profiles
O=gpurun_out; mkdir -p $O
timeout -k 10 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu > $O/r02m_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02m_pytest_gpu.log | tail -10
ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 600 python scripts/profile_step.py t2t 1 > $O/r02m_profile_t2t512_1.txt 2>&1; cat $O/r02m_profile_t2t512_1.txt
timeout 600 python scripts/profile_step.py l2t 1 > $O/r02m_profile_l2t_1.txt 2>&1; cat $O/r02m_profile_l2t_1.txt
ND_MINLEN=99 timeout 600 python scripts/profile_step.py l2t 5 > $O/r02m_profile_l2t_5.txt 2>&1; cat $O/r02m_profile_l2t_5.txt
