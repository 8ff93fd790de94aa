for cfg in "TVL1_T2=0" "TVL1_T2_FIRST=0" "TVL1_NO_TB=1" "TVL1_T2_STAGE=0" "TVL1_TAIL_PAIRS=0" "A=1"; do
  out=$(env TVL1_RES_MAX_CLUSTER=8 $cfg timeout 120 python profiles/run_solve.py 256 2>&1 | tail -2 | cut -c1-200)
  echo "== $cfg :: $out"
done
