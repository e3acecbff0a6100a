#!/bin/bash
# full GPU suite after GRU / front end / GEMM tile rule; smoke; CLI scale
O=gpurun_out; mkdir -p $O
timeout -k 10 1800 python -m pytest tests -q -m gpu > $O/r02k_pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02k_pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
