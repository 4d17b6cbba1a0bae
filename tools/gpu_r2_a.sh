# round-2 GPU session A: first run of the hand-written set-up kernels.
set -x
mkdir -p gpurun_out
S=gpurun_out/r2a_status.txt; rm -f $S
timeout 600 python -m pytest tests/test_gpu_setup.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_test_setup.log 2>&1; echo "setup tests rc=$?" >> $S
timeout 300 python tools/bench_setup.py --batch 4 --n 4096 --alpha 0.5 > gpurun_out/r2a_setup_b4.json 2> gpurun_out/r2a_setup_b4.err; echo "bench b4 rc=$?" >> $S
timeout 400 python tools/bench_setup.py --batch 16 --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2a_setup_b16.json 2> gpurun_out/r2a_setup_b16.err; echo "bench b16 rc=$?" >> $S
timeout 300 python tools/bench_setup.py --batch 1 --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2a_setup_b1.json 2> gpurun_out/r2a_setup_b1.err; echo "bench b1 rc=$?" >> $S
timeout 200 python tools/bench_setup.py --batch 64 --n 1000 --alpha 0.5 --skip-svd > gpurun_out/r2a_setup_n1000.json 2> gpurun_out/r2a_setup_n1000.err; echo "bench n1000 rc=$?" >> $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_test_all.log 2>&1; echo "all tests rc=$?" >> $S
cat $S; tail -30 gpurun_out/r2a_test_setup.log; cat gpurun_out/r2a_setup_b4.json gpurun_out/r2a_setup_b16.json gpurun_out/r2a_setup_b1.json gpurun_out/r2a_setup_n1000.json; tail -5 gpurun_out/r2a_setup_b4.err; tail -8 gpurun_out/r2a_test_all.log
