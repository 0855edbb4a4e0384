#!/bin/bash
# round 2, call F: timing ablations of the pair kernel (which resource is the wall?): one pass and sustained
mkdir -p gpurun_out
out=gpurun_out/r02_pair_ablations.jsonl; : > $out
for v in base ONE_TERM NO_REC_TRAFFIC NO_SWAP NO_MUFU; do
  if [ $v = base ]; then lib=forging_control_b200/libforging_b200.so; else lib=build/libfc_abl_$v.so; fi
  FC_LIB_PATH=$lib AB_MODE=3 AB_B=37888 AB_K=20 timeout 300 python scripts/ab_sustained.py >> $out 2>&1
  FC_LIB_PATH=$lib AB_MODE=3 timeout 300 python scripts/ab_sustained.py >> $out 2>&1
done
cat $out
