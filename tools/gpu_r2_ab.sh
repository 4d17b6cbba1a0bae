# round-2 GPU session AB: chunked updates with one memory round trip -- in-pipeline timeline, GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python tools/time_stages.py --out gpurun_out/r2ab_time_stages.json > gpurun_out/r2ab_time_stages.log 2>&1; echo "timeline rc=$?"
tail -9 gpurun_out/r2ab_time_stages.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=8 > gpurun_out/r2ab_test_all.log 2>&1; echo "all tests rc=$?"
tail -12 gpurun_out/r2ab_test_all.log
