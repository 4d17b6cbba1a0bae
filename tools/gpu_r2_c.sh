# round-2 GPU session C: set-up kernels after the eigen-kernel fix; inner sweeps 1 vs 2; CTA waves
set -x
mkdir -p gpurun_out
for cfg in "1 8" "2 8" "1 4" "1 2" "1 16"; do
  set -- $cfg
  timeout 300 python tools/bench_setup.py --batch 16 --n 4096 --alpha 0.5 --skip-svd --skip-gram --inner $1 --waves $2 > gpurun_out/r2c_setup_b16_i$1_w$2.json 2> gpurun_out/r2c_err.log
done
timeout 300 python tools/bench_setup.py --batch 64 --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2c_setup_b64.json 2>> gpurun_out/r2c_err.log
timeout 300 python tools/bench_setup.py --batch 64 --n 1000 --alpha 0.5 --skip-svd --direct > gpurun_out/r2c_setup_n1000.json 2>> gpurun_out/r2c_err.log
timeout 300 python tools/bench_setup.py --batch 8 --n 2000 --alpha 2.0 --direct > gpurun_out/r2c_setup_n2000a2.json 2>> gpurun_out/r2c_err.log
timeout 600 python -m pytest tests/test_gpu_setup.py tests/test_gpu_api.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1
cat gpurun_out/r2c_setup_*.json | cut -c1-1500; tail -5 gpurun_out/r2c_err.log; tail -5 gpurun_out/r2c_tests.log
