#!/bin/bash
# ring beam kernel, leaner consumer code: parity subset + timing
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -q -m gpu -k "beam or object" > $O/r02r_pytest_beam.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02r_pytest_beam.log | tail -10
for o in kv_beam_packed=1 kv_beam_packed=0 "kv_beam_packed=1,kv_mode=4"; do
  echo "== l2t beam 5, $o"; ND_MINLEN=99 ND_OPTS=$o timeout 300 python scripts/profile_step.py l2t 5 2>&1 | tail -9 | head -5
done
for k in 4 8; do
  echo "== l2t beam $k"; ND_MINLEN=99 timeout 300 python scripts/profile_step.py l2t $k 512 2>&1 | tail -9 | head -4
done
