#!/bin/bash
# compute-sanitizer over tools/sanitizer_workload.py, one log per tool under gpurun_out/
T=${1:-r2s}
mkdir -p gpurun_out
python tools/sanitizer_workload.py > gpurun_out/${T}_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/${T}_plain.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/${T}_sanitizer_${tool}.log python tools/sanitizer_workload.py > gpurun_out/${T}_${tool}_stdout.log 2>&1
  echo "$tool rc=$? $(grep -c 'done' gpurun_out/${T}_${tool}_stdout.log) workloads; $(tail -1 gpurun_out/${T}_sanitizer_${tool}.log)"
done
