#!/bin/bash
# multi-query fixed-point beam cross attention: parity tests, then A/B timing against the fp32 ring kernel on C3
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -q -m gpu -k "beam or object or fixed_point" > $O/r02o_pytest_beam.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02o_pytest_beam.log | tail -10
for o in kv_beam_packed=1 kv_beam_packed=0 "kv_beam_packed=1,kv_mode=4"; do
  echo "== l2t beam 5, $o"; ND_MINLEN=99 ND_OPTS=$o timeout 300 python scripts/profile_step.py l2t 5 2>&1 | tail -9
done
echo "== t2t512 beam 5 packed"; ND_MINLEN=99 ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 300 python scripts/profile_step.py t2t 5 256 2>&1 | tail -8
echo "== t2t512 beam 5 fp32"; ND_OPTS=kv_beam_packed=0 ND_MINLEN=99 ND_KW="dict(d_model=512,enc_layers=6,dec_layers=6)" timeout 300 python scripts/profile_step.py t2t 5 256 2>&1 | tail -8
