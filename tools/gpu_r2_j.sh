# round-2 GPU session J: device-resident factor-by-factor schedule on the GPU, whole suite, published protocol
set -x
mkdir -p gpurun_out
S=gpurun_out/r2j_status.txt; rm -f $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2j_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 600 python tools/bench_published_protocol.py > gpurun_out/r2j_published.log 2>&1; echo "published rc=$?" >> $S
timeout 300 python tools/bench_small_configs.py > gpurun_out/r2j_small_configs.log 2>&1; echo "small configs rc=$?" >> $S
cat $S; tail -8 gpurun_out/r2j_test_all.log; tail -12 gpurun_out/r2j_published.log | cut -c1-700; tail -5 gpurun_out/r2j_small_configs.log | cut -c1-600
