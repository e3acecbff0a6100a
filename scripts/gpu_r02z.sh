#!/bin/bash
# round 2, final 2-GPU sanity: bench.py under torchrun exactly as the driver launches it, and the reference arm
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02z_bench_n2.json 2> $O/r02z_bench_n2.err; echo "bench n2 exit $?"; tail -3 $O/r02z_bench_n2.err; cut -c1-900 $O/r02z_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r02z_bench_ref_n2.json 2> $O/r02z_bench_ref_n2.err; echo "bench ref n2 exit $?"; cut -c1-400 $O/r02z_bench_ref_n2.json
