#!/bin/bash
# Round evidence: GPU tests, smoke, event profiles, bench (both arms), ncu launch list + full captures.
# usage: bash scripts/gpu_evidence.sh <tag>      (outputs under gpurun_out/<tag>_*)
TAG=${1:-ev}
O=gpurun_out
mkdir -p $O
S=$O/${TAG}_summary.txt
rm -f $S
echo "=== pytest -m gpu" | tee -a $S
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" | tee -a $S
tail -4 $O/${TAG}_pytest_gpu.log | tee -a $S
echo "=== smoke" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $S
for f in "l2t 1" "l2t 5" "t2t 1" "nano2rnn 1" "brnn2rnn 1" "cnn2cnn 1"; do
  n=$(echo $f | tr ' ' '_')
  echo "=== profile_step $f" | tee -a $S
  timeout 300 python scripts/profile_step.py $f > $O/${TAG}_profile_$n.txt 2>&1; cat $O/${TAG}_profile_$n.txt | tee -a $S
done
echo "=== bench" | tee -a $S
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "exit $?" | tee -a $S
cat $O/${TAG}_bench.json | tee -a $S; tail -3 $O/${TAG}_bench.err
echo "=== bench reference" | tee -a $S
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cat $O/${TAG}_bench_ref.json | tee -a $S
if [ "$2" != "noncu" ]; then
echo "=== ncu launch list" | tee -a $S
python scripts/profile_step.py l2t 1 > $O/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/${TAG}_launches_l2t.csv python scripts/profile_step.py l2t 1 > $O/ncu_launch.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full: cross_attn, lstm, decode gemm" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn -s 310 -c 3 -o $O/${TAG}_prof_cross_attn -f python scripts/profile_step.py l2t 1 > $O/ncu_full1.log 2>&1; echo "exit $?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 3 -c 3 -o $O/${TAG}_prof_lstm -f python scripts/profile_step.py l2t 1 > $O/ncu_full2.log 2>&1; echo "exit $?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1212 -c 6 -o $O/${TAG}_prof_gemm -f python scripts/profile_step.py l2t 1 > $O/ncu_full3.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full: beam ring cross attention (beam 5, min_length 99)" | tee -a $S
ND_MINLEN=99 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cross_attn_ring -s 20 -c 2 -o $O/${TAG}_prof_cross_ring -f python scripts/profile_step.py l2t 5 > $O/ncu_full4.log 2>&1; echo "exit $?" | tee -a $S
fi
ls -la $O | tail -25
