set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 500 python -m pytest tests/test_gpu_se.py -q -p no:cacheprovider > gpurun_out/r1_test_se.log 2>&1; echo "se tests rc=$?" >> gpurun_out/r1_status.txt
timeout 300 python tools/bench_se.py --out gpurun_out/r01_state_evolution.json > gpurun_out/r1_bench_se.log 2>&1; echo "bench_se rc=$?" >> gpurun_out/r1_status.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_se_run -c 1 -o gpurun_out/r01c_se python tools/profile_se.py > gpurun_out/r1_ncu_se.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r1_status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1_status.txt
timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_se.py > gpurun_out/r1_test_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/r1_status.txt
timeout 400 python bench.py > gpurun_out/r1_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1_status.txt
cat gpurun_out/r1_status.txt; tail -5 gpurun_out/r1_test_se.log
