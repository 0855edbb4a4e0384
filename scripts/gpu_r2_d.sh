#!/bin/bash
# round 2, call D: cycle split + ncu capture of the quint kernel (one pass, B = 37888)
mkdir -p gpurun_out
FC_LIB_PATH=build/libforging_b200_timing.so FC_TC_TIMING=1 AB_MODE=4 AB_B=37888 AB_K=1 timeout 300 python scripts/ab_sustained.py > gpurun_out/r02_quint_timing_B37888.txt 2>&1
tail -4 gpurun_out/r02_quint_timing_B37888.txt
AB_MODE=4 AB_B=37888 AB_K=1 timeout 300 python scripts/ab_sustained.py > gpurun_out/plain_q.log 2>&1 &&
AB_MODE=4 AB_B=37888 AB_K=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_loss_quint -s 3 -c 1 -f -o gpurun_out/prof_quint python scripts/ab_sustained.py > gpurun_out/ncu_quint.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_quint.log
