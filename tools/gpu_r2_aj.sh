# round-2 GPU session AJ: the kernel-choice tests with the multi-panel case added
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_api.py -x -q -m gpu -p no:cacheprovider -k "rescale_inside or chunked_update or cluster_split or graph_replay" > gpurun_out/r2aj_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2aj_tests.log
