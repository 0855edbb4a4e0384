# ncu launch lists and full captures kept under profiles/ (run through gpurun; each command exits 0 without ncu first)
set -x
cd $GRAFT_REPO_ROOT
# (1) the commands exit 0 without ncu first
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-closed-loop --no-parity > gpurun_out/r02b_plain.log 2>&1 || exit 1
# (2) launch list of the same command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-closed-loop --no-parity > gpurun_out/r02b_ncu_launches.log 2>&1
# (3) full capture of the pair kernel, one pass (B = 37888)
ncu --set full --clock-control none --import-source on -k regex:mpc_loss_pair_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02b_prof_pair \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-closed-loop --no-parity --batch-per-gpu 37888 > gpurun_out/r02b_ncu_pair.log 2>&1
# (4) replica kernel (B = 4096, N = 5 and B = 15 N = 10 come from bench_midbatch; capture the B = 4096 N = 10 launch)
AB_B=4096 AB_K=2 python scripts/ab_sustained.py > gpurun_out/r02b_plain_replica.log 2>&1 || exit 1
AB_B=4096 AB_K=2 ncu --set full --clock-control none --import-source on -k regex:mpc_loss_replica_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02b_prof_replica \
  python scripts/ab_sustained.py > gpurun_out/r02b_ncu_replica.log 2>&1
# (5) surrogate training, tensor-core path: launch list + full captures of the training kernel (forward + reverse) and dw_kernel
python scripts/prof_surrogate_tc.py 37888 > gpurun_out/r02b_plain_surrogate.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02b_launches_surrogate.csv \
  python scripts/prof_surrogate_tc.py 37888 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_train_pair_kernel --launch-skip 5 -c 1 -f -o gpurun_out/r02b_prof_train \
  python scripts/prof_surrogate_tc.py 37888 > gpurun_out/r02b_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dw_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02b_prof_dw \
  python scripts/prof_surrogate_tc.py 37888 > gpurun_out/r02b_ncu_dw.log 2>&1
ls -la gpurun_out/*.ncu-rep
