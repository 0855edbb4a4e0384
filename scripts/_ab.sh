for i in 1 2; do
  for v in prev cur; do
    if [ $v = cur ]; then unset FC_LIB_PATH; else export FC_LIB_PATH=build/libfc_$v.so; fi
    python scripts/ab_sustained.py | cut -c1-120
    AB_B=37888 AB_K=20 python scripts/ab_sustained.py | cut -c1-120
  done
done
