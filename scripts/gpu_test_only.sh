#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/pytest_gpu.log
