#!/bin/bash
# round 2, call C: first GPU run of the quint kernel: parity (all MPC-loss tests), A/B against the pair kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mpc_loss.py -m gpu -q -x --timeout 300 -k "quint" > gpurun_out/r02_pytest_quint.log 2>&1; echo "pytest quint rc=$?"
tail -15 gpurun_out/r02_pytest_quint.log
for mode in 3 4 3 4; do AB_MODE=$mode timeout 300 python scripts/ab_sustained.py; done > gpurun_out/r02_ab_quint.jsonl 2>&1
cat gpurun_out/r02_ab_quint.jsonl
for mode in 3 4; do AB_MODE=$mode AB_B=37888 AB_K=20 timeout 300 python scripts/ab_sustained.py; done > gpurun_out/r02_ab_quint_onepass.jsonl 2>&1
cat gpurun_out/r02_ab_quint_onepass.jsonl
