#!/bin/bash
# round 2, call G: one-tile kernel with 32/64/96-row tiles + quadrant skipping: parity and per-config timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpc_loss.py -m gpu -q -x --timeout 300 -k "tcgen05" > gpurun_out/r02_pytest_tc.log 2>&1; echo "pytest tc rc=$?"
tail -3 gpurun_out/r02_pytest_tc.log
timeout 600 python scripts/bench_configs.py > gpurun_out/r02_bench_configs_tile_rows.jsonl 2>&1
grep -v ffma gpurun_out/r02_bench_configs_tile_rows.jsonl
