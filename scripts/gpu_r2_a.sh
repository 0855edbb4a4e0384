#!/bin/bash
# round 2, call A: in-kernel cycle breakdown of the pair kernel (timing build) + per-config timings of the round-1 tree
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
FC_LIB_PATH=build/libforging_b200_timing.so FC_TC_TIMING=1 AB_B=37888 AB_K=2 timeout 300 python scripts/ab_sustained.py > gpurun_out/r02_timing_B37888.txt 2>&1
FC_LIB_PATH=build/libforging_b200_timing.so FC_TC_TIMING=1 AB_B=524288 AB_K=2 timeout 300 python scripts/ab_sustained.py > gpurun_out/r02_timing_B524288.txt 2>&1
timeout 300 python scripts/ab_sustained.py > gpurun_out/r02_ab_base.txt 2>&1
timeout 600 python scripts/bench_configs.py > gpurun_out/r02_bench_configs_base.jsonl 2>&1
tail -4 gpurun_out/r02_timing_B524288.txt; cat gpurun_out/r02_ab_base.txt
