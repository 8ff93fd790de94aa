for cfg in "TVL1_T2_SHARED=0" "TVL1_T2_STAGE_SHARED=0" "TVL1_T2_SHARED=0" "TVL1_T2_STAGE_SHARED=0" "TVL1_T2_SHARED=0" "TVL1_T2_STAGE_SHARED=0" "A=1"; do
  out=$(env TVL1_RES_MAX_CLUSTER=8 REPS=4 $cfg timeout 150 python profiles/run_e2e_pipe.py "[(16,3,{})]" 2>&1 | grep "max_batch" | cut -c1-30,100-130,150-250)
  echo "== $cfg :: $out"
done
