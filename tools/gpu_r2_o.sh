# round-2 GPU session O: per-kernel times of the single large instance (config 5, one GPU, unsharded), new GPU tests
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_setup.py -m gpu -q -x -p no:cacheprovider -k "rank_deficient or fused" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2o_status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 100 -c 200 --csv --log-file gpurun_out/r2o_config5_launches.csv python -c "
import json, sys
sys.path.insert(0, '.')
from tools import bench_blocks
print(json.dumps(bench_blocks.row_sharded_block(world=1, iters=8)))
" > gpurun_out/r2o_config5.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2o_status.txt
cat gpurun_out/r2o_status.txt; tail -3 gpurun_out/r2o_tests.log; tail -2 gpurun_out/r2o_config5.log | cut -c1-600
