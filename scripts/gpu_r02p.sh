#!/bin/bash
# ring beam kernel over fixed-point planes: parity tests, then A/B timing against fp32 rows on C3
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -q -m gpu -k "beam or object or fixed_point" > $O/r02p_pytest_beam.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02p_pytest_beam.log | tail -10
for o in kv_beam_packed=1 kv_beam_packed=0 "kv_beam_packed=1,kv_mode=4"; do
  echo "== l2t beam 5, $o"; ND_MINLEN=99 ND_OPTS=$o timeout 300 python scripts/profile_step.py l2t 5 2>&1 | tail -9 | head -5
done
for k in 2 4 8; do
for o in kv_beam_packed=1 kv_beam_packed=0; do
  echo "== l2t beam $k, $o"; ND_MINLEN=99 ND_OPTS=$o timeout 300 python scripts/profile_step.py l2t $k 512 2>&1 | tail -9 | head -4
done; done
