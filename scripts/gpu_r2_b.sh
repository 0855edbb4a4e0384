#!/bin/bash
# round 2, call B: GPU parity tests + the new bench line (closed loop, parity checks, reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r02_bench.log; tail -5 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.log 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
tail -c 3000 gpurun_out/r02_bench_ref.log
