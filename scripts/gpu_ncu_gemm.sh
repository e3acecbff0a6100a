#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_step.py l2t 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 0 -c 12 -o /tmp/prof_gemm -f python scripts/profile_step.py l2t 1 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/gemm_raw.csv 2>/dev/null
ncu -i /tmp/prof_gemm.ncu-rep --page source --csv --print-source sass > gpurun_out/gemm_source_sass.csv 2>/dev/null
ncu -i /tmp/prof_gemm.ncu-rep --page source --csv --print-source cuda > gpurun_out/gemm_source_cuda.csv 2>/dev/null
ls -la /tmp/prof_gemm.ncu-rep gpurun_out/
gzip -f gpurun_out/gemm_source_sass.csv
