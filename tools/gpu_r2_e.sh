# round-2 GPU session E: fused one-iteration-back state (no snapshot kernel), 2 CTAs/SM update kernels
set -x
mkdir -p gpurun_out
S=gpurun_out/r2e_status.txt; rm -f $S
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2e_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 500 --csv --log-file gpurun_out/r2e_sweep_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2e_ncu_bench.log 2>&1; echo "ncu rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --setup-instances 0 > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?" >> $S
cat $S; tail -8 gpurun_out/r2e_test_all.log; cut -c1-400 gpurun_out/r2e_bench_1gpu.json
