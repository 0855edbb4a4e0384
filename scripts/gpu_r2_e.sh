#!/bin/bash
# round 2, call E: quint kernel v2 (unit-pair chunks): parity, A/B against the pair kernel, cycle split, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mpc_loss.py -m gpu -q -x --timeout 300 -k "quint" > gpurun_out/r02_pytest_quint.log 2>&1; echo "pytest quint rc=$?"
tail -3 gpurun_out/r02_pytest_quint.log
for mode in 3 4 3 4; do AB_MODE=$mode timeout 300 python scripts/ab_sustained.py; done > gpurun_out/r02_ab_quint2.jsonl 2>&1
cat gpurun_out/r02_ab_quint2.jsonl
for mode in 3 4; do AB_MODE=$mode AB_B=37888 AB_K=20 timeout 300 python scripts/ab_sustained.py; done > gpurun_out/r02_ab_quint2_onepass.jsonl 2>&1
cat gpurun_out/r02_ab_quint2_onepass.jsonl
FC_LIB_PATH=build/libforging_b200_timing.so FC_TC_TIMING=1 AB_MODE=4 AB_B=37888 AB_K=1 timeout 300 python scripts/ab_sustained.py > gpurun_out/r02_quint2_timing_B37888.txt 2>&1
tail -4 gpurun_out/r02_quint2_timing_B37888.txt
AB_MODE=4 AB_B=37888 AB_K=1 timeout 300 python scripts/ab_sustained.py > gpurun_out/plain_q.log 2>&1 &&
AB_MODE=4 AB_B=37888 AB_K=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_loss_quint -s 3 -c 1 -f -o gpurun_out/prof_quint2 python scripts/ab_sustained.py > gpurun_out/ncu_quint2.log 2>&1
echo "ncu rc=$?"
