#!/bin/bash
# round 2, third call: fast packed cross-attention variants, full GPU test suite, beam identity rates
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/kv_modes.py 0,3,4,5 1,2,0 > $O/r02c_kv_modes.txt 2>&1; echo "kv_modes exit $?"; cat $O/r02c_kv_modes.txt
timeout -k 10 1200 python -m pytest tests -q -m gpu > $O/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02c_pytest_gpu.log | tail -30
timeout 1800 python scripts/identity_rates.py --only beam --out $O/r02c_identity_beam.json > $O/r02c_identity_beam.log 2>&1; echo "identity exit $?"
cut -c1-600 $O/r02c_identity_beam.log | tail -12
