#!/bin/bash
# ncu capture of the fixed-point mlp attention kernel (nano2rnn, q23) and of the fp32 one beside it
O=gpurun_out; mkdir -p $O
cap() {  # name, regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/r02v_$name -f "$@" > $O/ncu_v_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i $O/r02v_$name.ncu-rep --page raw --csv > $O/r02v_$name.raw.csv 2>/dev/null; gzip -f $O/r02v_$name.raw.csv
  ncu -i $O/r02v_$name.ncu-rep --page source --csv > $O/r02v_$name.source.csv 2>/dev/null; gzip -f $O/r02v_$name.source.csv
  rm -f $O/r02v_$name.ncu-rep
}
ND_OPTS=kv_mode=3 cap mlp_q23 mlp_attn_packed 120 1 python scripts/profile_step.py nano2rnn 1
ND_OPTS=kv_mode=0 cap mlp_f32 mlp_attn_kernel 120 1 python scripts/profile_step.py nano2rnn 1
ls -la $O | grep r02v
