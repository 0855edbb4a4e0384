export FC_MPC_KERNEL=pair
for v in default NOPOLY NOPOLY_FOLD; do
  if [ $v = default ]; then unset FC_LIB_PATH; else export FC_LIB_PATH=build/libfc_$v.so; fi
  echo "=== $v"
  python scripts/diag_precision.py 2>&1 | tail -20
  python scripts/diag_trace.py 2>&1 | tail -1
done
