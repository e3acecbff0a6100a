#!/bin/bash
# ring beam kernel over fixed-point planes, any stage count: parity subset, timing, one ncu capture with the source page
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -q -m gpu -k "beam and (kernels_agree or nondegenerate or object)" > $O/r02q_pytest_beam.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" $O/r02q_pytest_beam.log | tail -10
for o in kv_beam_packed=1 kv_beam_packed=0 "kv_beam_packed=1,kv_mode=4"; do
  echo "== l2t beam 5, $o"; ND_MINLEN=99 ND_OPTS=$o timeout 300 python scripts/profile_step.py l2t 5 2>&1 | tail -9 | head -5
done
cap() {  # name, regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/r02q_$name -f "$@" > $O/ncu_q_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i $O/r02q_$name.ncu-rep --page raw --csv > $O/r02q_$name.raw.csv 2>/dev/null; gzip -f $O/r02q_$name.raw.csv
  ncu -i $O/r02q_$name.ncu-rep --page source --csv > $O/r02q_$name.source.csv 2>/dev/null; gzip -f $O/r02q_$name.source.csv
  rm -f $O/r02q_$name.ncu-rep
}
ND_MINLEN=99 cap ring_q23 cross_attn_ring 400 1 python scripts/profile_step.py l2t 5
ND_MINLEN=99 ND_OPTS=kv_mode=4 cap ring_q15 cross_attn_ring 400 1 python scripts/profile_step.py l2t 5
ls -la $O | grep r02q
