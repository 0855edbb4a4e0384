#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list + one full capture of the dominant kernel of bench.py.
mkdir -p gpurun_out
B=${PROF_BATCH:-37888}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-closed-loop --no-parity --batch-per-gpu $B"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${PROF_KERNEL:-mpc_loss_pair_kernel} -s 3 -c 1 -f -o gpurun_out/prof_mpc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain.log | cut -c1-400
