#!/bin/bash
# Run on the GPU box (via gpurun): smoke, GPU parity tests, bench (both arms).  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup ${BENCH_WARMUP:-5} ${BENCH_EXTRA} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps ${BENCH_STEPS:-20} --warmup ${BENCH_WARMUP:-5} > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
tail -c 600 gpurun_out/bench_ref.log
