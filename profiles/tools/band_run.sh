N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_bands_over_all" > gpurun_out/r2f_bandtest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_bandtest_n$N.log
$TR --master-port 29511 bench.py --gpus $N --workload band4k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_band4k_n$N.json 2> gpurun_out/r2f_band4k_n$N.err; echo "band4k rc=$?"
$TR --master-port 29512 bench.py --gpus $N --workload band8k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_band8k_n$N.json 2> gpurun_out/r2f_band8k_n$N.err; echo "band8k rc=$?"
tail -3 gpurun_out/r2f_bandtest_n$N.log
$TR --master-port 29521 profiles/run_band_levels.py 8k 1024 2>&1 | grep "^{" > gpurun_out/r2f_band_levels_8k_n$N.json
$TR --master-port 29522 profiles/run_band_levels.py 4k 512 2>&1 | grep "^{" > gpurun_out/r2f_band_levels_4k_n$N.json
cat gpurun_out/r2f_band_levels_*_n$N.json
