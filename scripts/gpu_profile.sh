#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list + one full capture of the dominant kernel.
mkdir -p gpurun_out
B=${PROF_BATCH:-71040}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch-per-gpu $B"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${PROF_KERNEL:-mpc_loss_kernel} -s 1 -c 1 -f -o gpurun_out/prof_mpc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain.log
