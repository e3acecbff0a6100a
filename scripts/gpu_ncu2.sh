#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_step.py l2t 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1212 -c 6 -o /tmp/prof_gemm -f python scripts/profile_step.py l2t 1 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/gemm2_raw.csv 2>/dev/null
ncu -i /tmp/prof_gemm.ncu-rep --page source --csv --print-source sass > gpurun_out/gemm2_source_sass.csv 2>/dev/null
gzip -f gpurun_out/gemm2_source_sass.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_attn -s 520 -c 2 -o /tmp/prof_self -f python scripts/profile_step.py l2t 1 > gpurun_out/ncu_self.log 2>&1
ncu -i /tmp/prof_self.ncu-rep --page raw --csv > gpurun_out/self_raw.csv 2>/dev/null
ncu -i /tmp/prof_self.ncu-rep --page source --csv --print-source sass > gpurun_out/self_source_sass.csv 2>/dev/null
gzip -f gpurun_out/self_source_sass.csv
ls -la gpurun_out /tmp/*.ncu-rep
