#!/bin/bash
for pdl in 2 0; do
  echo "=== pdl=$pdl"
  ND_PDL=$pdl timeout 300 python scripts/profile_step.py l2t 1 2>&1 | head -3
done
