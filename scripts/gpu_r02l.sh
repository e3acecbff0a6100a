#!/bin/bash
# ncu of decode self attention at step 50 (d = 512) and of the generator kernel
O=gpurun_out; mkdir -p $O
C5='dict(d_model=512,enc_layers=6,dec_layers=6)'
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/r02l_$name -f "$@" > $O/ncu_l_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i $O/r02l_$name.ncu-rep --page raw --csv > $O/r02l_$name.raw.csv 2>/dev/null; gzip -f $O/r02l_$name.raw.csv
  ncu -i $O/r02l_$name.ncu-rep --page source --csv > $O/r02l_$name.source.csv 2>/dev/null; gzip -f $O/r02l_$name.source.csv
  rm -f $O/r02l_$name.ncu-rep
}
ND_KW="$C5" cap self_attn_t50 self_attn_kernel 300 1 python scripts/profile_step.py t2t 1
ND_KW="$C5" cap generator generator_kernel 50 1 python scripts/profile_step.py t2t 1
ls -la $O | grep r02l
