#!/bin/bash
# fixed-point attention memory of the RNN / CNN decoders: parity tests, then 1024-chunk identity rates vs the oracle
O=gpurun_out; mkdir -p $O
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rnn or cnn or nano2 or attention_memory" > $O/r02x_pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02x_pytest.log | tail -10
timeout 1500 python scripts/identity_rates.py --families nano2rnn,brnn2rnn,cnn2cnn --out $O/r02x_identity.json > $O/r02x_identity.log 2>&1; echo "identity exit $?"
cut -c1-330 $O/r02x_identity.log | tail -14
