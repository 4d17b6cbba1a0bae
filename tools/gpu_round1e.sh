# round-1e GPU session: persistent sweep after the memory-level-parallelism rewrite
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3_status.txt
timeout 400 python -m pytest tests/test_gpu_api.py -q -p no:cacheprovider -k "persistent or sweep_matches or early_stopping" > gpurun_out/r3_test_persistent.log 2>&1; echo "persistent tests rc=$?" >> gpurun_out/r3_status.txt
timeout 300 python tools/bench_small_configs.py > gpurun_out/r3_small.log 2>&1; echo "small configs rc=$?" >> gpurun_out/r3_status.txt
cp gpurun_out/r01_small_configs.json gpurun_out/r01e_small_configs.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_sweep_persistent -c 1 -o gpurun_out/r01e_persist python tools/profile_persistent.py > gpurun_out/r3_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r3_status.txt
cat gpurun_out/r3_status.txt; tail -5 gpurun_out/r3_test_persistent.log; tail -3 gpurun_out/r3_small.log | cut -c1-1500
