#!/bin/bash
# front end: histogram statistics + vectorised chunk gather vs the general kernels
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -k "frontend or cli or smoke" 2>&1 | tail -5
timeout 600 python scripts/bench_frontend.py 1024 > $O/r02i_frontend_bench.txt 2>&1; cat $O/r02i_frontend_bench.txt
