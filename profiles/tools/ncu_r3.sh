# ncu --set full capture of the two-iterations-per-launch kernel (one GPU), after the same command exited 0 without ncu
set -x
NCU="ncu --set full --clock-control none --import-source on"
TVL1_BENCH_TB=2 python profiles/run_iterate.py 32 1920 1080 10 > gpurun_out/r3m_run_iterate_t2.txt 2>&1 || exit 1
TVL1_BENCH_TB=2 $NCU -k regex:k_iterate_t2 -s 3 -c 1 -o gpurun_out/r3m_iterate_t2 python profiles/run_iterate.py 32 1920 1080 10 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
