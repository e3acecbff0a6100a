#!/bin/bash
# packed mlp / conv attention, 4 CTAs per SM: A/B timing on the C4 workloads
for f in nano2rnn cnn2cnn; do for m in 3 0 4; do
  echo "== $f kv_mode=$m"; ND_OPTS=kv_mode=$m timeout 300 python scripts/profile_step.py $f 1 2>&1 | tail -8 | grep -E "graph replay|mlp_attn"
done; done
echo "== nano2rnn beam 5 kv 3 / 0"; for m in 3 0; do ND_MINLEN=99 ND_OPTS=kv_mode=$m timeout 300 python scripts/profile_step.py nano2rnn 5 2>&1 | tail -8 | grep -E "graph replay|mlp_attn"; done
