# round-1d GPU session: persistent single-instance sweep + State Evolution re-check
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_status.txt
timeout 400 python -m pytest tests/test_gpu_api.py -q -p no:cacheprovider -k "persistent" > gpurun_out/r2_test_persistent.log 2>&1; echo "persistent tests rc=$?" >> gpurun_out/r2_status.txt
timeout 300 python -m pytest tests/test_gpu_se.py -q -p no:cacheprovider > gpurun_out/r2_test_se.log 2>&1; echo "se tests rc=$?" >> gpurun_out/r2_status.txt
timeout 300 python tools/bench_small_configs.py > gpurun_out/r2_small.log 2>&1; echo "small configs rc=$?" >> gpurun_out/r2_status.txt
timeout 200 python tools/bench_se.py --out gpurun_out/r01d_state_evolution.json > gpurun_out/r2_bench_se.log 2>&1; echo "bench_se rc=$?" >> gpurun_out/r2_status.txt
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_se.py -k "not persistent" > gpurun_out/r2_test_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/r2_status.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_se_run -c 1 -o gpurun_out/r01d_se python tools/profile_se.py > gpurun_out/r2_ncu_se.log 2>&1; echo "ncu se rc=$?" >> gpurun_out/r2_status.txt
cat gpurun_out/r2_status.txt; tail -15 gpurun_out/r2_test_persistent.log
