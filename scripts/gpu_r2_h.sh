#!/bin/bash
# round 2, call H: closed-loop kernel with MUFU transcendentals: parity tests, throughput, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_closed_loop.py -m gpu -q -x --timeout 600 > gpurun_out/r02_pytest_cl.log 2>&1; echo "pytest cl rc=$?"
tail -15 gpurun_out/r02_pytest_cl.log
timeout 600 python scripts/bench_closed_loop.py > gpurun_out/r02_closed_loop_bench.json 2> gpurun_out/r02_closed_loop_bench.err; echo "bench rc=$?"
cat gpurun_out/r02_closed_loop_bench.json; tail -3 gpurun_out/r02_closed_loop_bench.err
timeout 300 python scripts/bench_closed_loop.py --batch 262144 --steps 300 --log-batch 4096 > gpurun_out/cl_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:closed_loop_kernel -s 1 -c 1 -f -o gpurun_out/prof_cl2 python scripts/bench_closed_loop.py --batch 262144 --steps 300 --log-batch 4096 > gpurun_out/ncu_cl2.log 2>&1
echo "ncu rc=$?"
