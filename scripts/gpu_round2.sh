#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
echo "=== pytest -m gpu" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary.txt
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a gpurun_out/summary.txt
echo "=== profile_step l2t" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/profile_step.py l2t 1 > gpurun_out/profile_l2t.log 2>&1; cat gpurun_out/profile_l2t.log | tee -a gpurun_out/summary.txt
echo "=== profile_step l2t beam5" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/profile_step.py l2t 5 > gpurun_out/profile_l2t_beam.log 2>&1; cat gpurun_out/profile_l2t_beam.log | tee -a gpurun_out/summary.txt
echo "=== profile_step t2t" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/profile_step.py t2t 1 > gpurun_out/profile_t2t.log 2>&1; cat gpurun_out/profile_t2t.log | tee -a gpurun_out/summary.txt
echo "=== profile_step nano2rnn" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/profile_step.py nano2rnn 1 > gpurun_out/profile_nano2rnn.log 2>&1; cat gpurun_out/profile_nano2rnn.log | tee -a gpurun_out/summary.txt
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/bench.err
echo "=== bench reference" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | tee -a gpurun_out/summary.txt
echo "=== ncu launch list" | tee -a gpurun_out/summary.txt
python scripts/profile_step.py l2t 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches_l2t.csv python scripts/profile_step.py l2t 1 > gpurun_out/ncu_launch.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
echo "=== ncu full cross_attn" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn -s 310 -c 3 -o gpurun_out/prof_cross_attn -f python scripts/profile_step.py l2t 1 > gpurun_out/ncu_full.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out | tail -20
