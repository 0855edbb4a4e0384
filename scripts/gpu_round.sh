#!/bin/bash
# Run on the GPU box (via gpurun): smoke, GPU parity tests, short bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E 'PASSED|FAILED|ERROR|passed|failed|closed loop' gpurun_out/pytest_gpu.log | tail -80
timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_EXTRA} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
