#!/bin/bash
# fixed-point planes for the RNN / CNN decoders' attention memory: parity, then A/B timing on the C4 workloads
O=gpurun_out; mkdir -p $O
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "rnn or cnn or nano2 or attention_memory" > $O/r02u_pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02u_pytest.log | tail -10
grep "kv_mode" $O/r02u_pytest.log | grep "rel err" | head -24
for f in nano2rnn brnn2rnn cnn2cnn; do for m in 3 0 4; do
  echo "== $f kv_mode=$m"; ND_OPTS=kv_mode=$m timeout 300 python scripts/profile_step.py $f 1 2>&1 | tail -8 | head -4
done; done
echo "== nano2rnn beam 5 kv 3 / 0"; for m in 3 0; do ND_MINLEN=99 ND_OPTS=kv_mode=$m timeout 300 python scripts/profile_step.py nano2rnn 5 2>&1 | tail -8 | head -3; done
