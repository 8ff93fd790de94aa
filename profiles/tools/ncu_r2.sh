# ncu --set full captures of round 2 (one GPU).  Each capture only after the same command exited 0 without ncu.
set -x
NCU="ncu --set full --clock-control none --import-source on"
python profiles/run_solve.py 32 > gpurun_out/r2i_run_solve32.txt 2>&1 || exit 1
TVL1_NO_GRAPH=1 $NCU -k regex:^k_warp -s 20 -c 1 -o gpurun_out/r2i_warp python profiles/run_solve.py 32 > /dev/null 2>&1
TVL1_NO_GRAPH=1 $NCU -k regex:^k_gauss_march -s 4 -c 2 -o gpurun_out/r2i_gauss python profiles/run_solve.py 32 > /dev/null 2>&1
TVL1_NO_GRAPH=1 $NCU -k regex:^k_zoom_in_flow -s 7 -c 1 -o gpurun_out/r2i_zoom_in python profiles/run_solve.py 32 > /dev/null 2>&1
python profiles/run_iterate.py 32 1920 1080 10 > gpurun_out/r2i_run_iterate.txt 2>&1
$NCU -k regex:k_iterate_t1 -s 5 -c 1 -o gpurun_out/r2i_iterate_t1 python profiles/run_iterate.py 32 1920 1080 10 > /dev/null 2>&1
python profiles/run_occ.py 32 640 480 2 > gpurun_out/r2i_run_occ.txt 2>&1 || exit 1
$NCU -k regex:k_occ_rof_gs -s 12 -c 1 -o gpurun_out/r2i_occ_rof_gs python profiles/run_occ.py 32 640 480 2 > /dev/null 2>&1
$NCU -k regex:k_occ_chi_fused -s 40 -c 1 -o gpurun_out/r2i_occ_chi_fused python profiles/run_occ.py 32 640 480 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
