#!/bin/bash
# compute-sanitizer over the reduced case set (scripts/sanitizer_case.py); logs -> gpurun_out/<tag>_sanitizer_<tool>.log
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
python scripts/sanitizer_case.py > $O/${TAG}_sanitizer_plain.log 2>&1; echo "plain exit $?"; tail -2 $O/${TAG}_sanitizer_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 python scripts/sanitizer_case.py > $O/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|all cases" $O/${TAG}_sanitizer_$tool.log | tail -14
done
