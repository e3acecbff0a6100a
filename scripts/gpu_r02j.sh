#!/bin/bash
# ncu --set full captures of the round-2 kernels and of the other large items of the C5 step (one launch or two each)
O=gpurun_out; mkdir -p $O
C5='dict(d_model=512,enc_layers=6,dec_layers=6)'
cap() {  # name, regex, skip, count, env-prefixed command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/r02j_$name -f "$@" > $O/ncu_j_$name.log 2>&1
  echo "ncu $name exit $?"
  # only the metric tables travel back (gpurun_out is capped at 64 MiB): raw page as CSV, then drop the report
  ncu -i $O/r02j_$name.ncu-rep --page raw --csv > $O/r02j_$name.raw.csv 2>/dev/null; gzip -f $O/r02j_$name.raw.csv
  rm -f $O/r02j_$name.ncu-rep
}
ND_KW="$C5" python scripts/profile_step.py t2t 1 > $O/r02j_plain.log 2>&1; echo "plain exit $?"
ND_KW="$C5" cap self_attn_d512 self_attn_kernel 1210 2 python scripts/profile_step.py t2t 1
ND_KW="$C5" cap gemm_decode_d512 "gemm_tc_kernel" 1236 6 python scripts/profile_step.py t2t 1
ND_KW="$C5" cap gemm_persist_d512 gemm_tc_persist 30 4 python scripts/profile_step.py t2t 1
ND_KW="$C5" cap enc_attn64 enc_attn_tc64 6 1 python scripts/profile_step.py t2t 1
ND_KW="$C5" cap kv_pack kv_pack_kernel 6 1 python scripts/profile_step.py t2t 1
cap frontend "stats_hist_kernel|chunks_vec_kernel" 4 2 python scripts/bench_frontend.py 256
ND_KW="dict()" cap lstm_h256 "lstm_kernel" 3 1 python scripts/profile_step.py rnn2rnn 1
timeout 300 python scripts/profile_step.py rnn2rnn 1 > $O/r02j_profile_rnn2rnn_1.txt 2>&1; cat $O/r02j_profile_rnn2rnn_1.txt
ls -la $O | grep r02j
