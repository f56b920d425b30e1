#!/bin/bash
# one GPU-box visit for profiles/: launch list (time per kernel) and one `--set full` capture of steady-state iterations
# usage: bash tools/profile_round.sh TAG [bench args...]
T=${1:-r2}; shift
mkdir -p gpurun_out
ARGS="--only-headline --no-single --no-cpu-baseline --steps 1 --warmup 1 --iters 6 $@"
K='regex:k_spec|k_update|k_loss_stop|k_tick|k_rot|k_resample|k_point'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 27 -c 45 --csv --log-file gpurun_out/${T}_launches.csv python bench.py $ARGS > gpurun_out/${T}_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 27 -c 9 -f -o gpurun_out/${T}_full python bench.py $ARGS > gpurun_out/${T}_full.log 2>&1
ls -la gpurun_out/${T}_*
