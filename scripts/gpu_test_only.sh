#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_step.py l2t 1 2>&1
