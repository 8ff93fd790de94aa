# ncu --set full capture of the two-iterations-per-launch kernel (one GPU), after the same command exited 0 without ncu
set -x
NCU="ncu --set full --clock-control none --import-source on"
TVL1_BENCH_TB=2 python profiles/run_iterate.py 32 1920 1080 10 > gpurun_out/r3m_run_iterate_t2.txt 2>&1 || exit 1
TVL1_BENCH_TB=2 $NCU -k regex:k_iterate_t2 -s 3 -c 1 -o gpurun_out/r3m_iterate_t2 python profiles/run_iterate.py 32 1920 1080 10 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
# the cluster-resident kernel on level 2 (480x270, clusters of 16) of a 144-pair batch: first warp step of the level
python profiles/run_solve.py 144 > gpurun_out/r3w_run_solve144.txt 2>&1 || exit 1
TVL1_NO_GRAPH=1 $NCU -k regex:k_iterate_resident -s 10 -c 1 -o gpurun_out/r3w_resident_l2 python profiles/run_solve.py 144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
