# ncu --set full captures of the kernels rewritten late in round 2 (one GPU)
set -x
NCU="ncu --set full --clock-control none --import-source on"
python profiles/run_solve.py 32 > gpurun_out/r2p_run_solve32.txt 2>&1 || exit 1
TVL1_NO_GRAPH=1 $NCU -k regex:^k_gauss_shfl -c 2 -o gpurun_out/r2p_gauss_shfl python profiles/run_solve.py 32 > /dev/null 2>&1
TVL1_NO_GRAPH=1 $NCU -k regex:^k_zoom_in_flow -s 3 -c 1 -o gpurun_out/r2p_zoom_in python profiles/run_solve.py 32 > /dev/null 2>&1
python profiles/run_occ.py 74 640 480 1 > gpurun_out/r2p_run_occ.txt 2>&1 || exit 1
$NCU -k regex:k_occ_rof_gs_wave -s 2 -c 1 -o gpurun_out/r2p_occ_rof_gs_wave python profiles/run_occ.py 74 640 480 1 > /dev/null 2>&1
ls -la gpurun_out/r2p*.ncu-rep
