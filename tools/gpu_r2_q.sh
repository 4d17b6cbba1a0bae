# round-2 GPU session Q: single-pass spectrum rescale -- GPU suite + per-kernel times of the bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2q_test_all.log 2>&1; echo "all tests rc=$?" > gpurun_out/r2q_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r2q_bench_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2q_ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2q_status.txt
cat gpurun_out/r2q_status.txt; tail -4 gpurun_out/r2q_test_all.log
