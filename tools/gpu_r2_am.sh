# round-2 GPU session AM: last check of the in-tree build -- smoke, kernel-choice tests, multi-rank tests
set -x
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2am_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2am_smoke.log
timeout 400 python -m pytest tests/test_gpu_api.py tests/test_gpu_multi.py -x -q -m gpu -p no:cacheprovider -k "rescale_inside or chunked_update or graph_replay or batched or multi or shard" > gpurun_out/r2am_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2am_tests.log
