#!/bin/bash
# ResNet-stem encoders (resnet / crnn / ctransformer): golden parity on the GPU
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "resnet or crnn or ctrans" > $O/r02s_pytest_resnet.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR|Error|error" $O/r02s_pytest_resnet.log | tail -30
for f in resnet2t resnet2rnn; do echo "== $f greedy B=1024"; timeout 300 python scripts/profile_step.py $f 1 2>&1 | tail -12; done
