#!/bin/bash
for pdl in 1 0; do
  for f in "l2t 1" "nano2rnn 1"; do
    echo "=== pdl=$pdl profile_step $f"
    ND_PDL=$pdl timeout 300 python scripts/profile_step.py $f 2>&1 | head -3
  done
done
