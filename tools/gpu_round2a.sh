# round-2a GPU session (first of the round): what was written after round 1's GPU budget
# ran out, in the order that matters.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2a.sh'
# 1. the GPU tests that have never run on hardware (DESIGN 6 "Not yet run on a GPU"), on their own
#    so that a failure there does not hide the rest; 2. the whole GPU suite; 3. smoke + the
#    bench line; 4. the set-up variants, incl. svd_method="jacobi".
set -x
mkdir -p gpurun_out
S=gpurun_out/r02a_status.txt
rm -f $S
timeout 400 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_se_reference_examples.py -m gpu -q -p no:cacheprovider \
    > gpurun_out/r02a_test_new.log 2>&1; echo "new tests rc=$?" >> $S
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_test_all.log 2>&1; echo "all tests rc=$?" >> $S
# the north-star-shape checks again with 4 instances: CUDA-graph replay + 2-CTA clusters, a combination no other test reaches
TRB_TEST_INSTANCES=4 timeout 300 python -m pytest tests/test_gpu_sizes.py -m gpu -q -p no:cacheprovider \
    > gpurun_out/r02a_test_sizes_b4.log 2>&1; echo "sizes B=4 rc=$?" >> $S
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 400 python bench.py > gpurun_out/r02a_bench_1gpu.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> $S
timeout 300 python tools/bench_setup.py --batch 8 --n 4096 --alpha 0.5 --streams 2 4 8 --jacobi-blocks 16 \
    > gpurun_out/r02a_setup_n4096.json 2> gpurun_out/r02a_setup.err; echo "setup n4096 rc=$?" >> $S
timeout 200 python tools/bench_setup.py --batch 64 --n 1000 --alpha 0.5 --streams 4 16 --jacobi-blocks 16 \
    > gpurun_out/r02a_setup_n1000.json 2>> gpurun_out/r02a_setup.err; echo "setup n1000 rc=$?" >> $S
cat $S; tail -15 gpurun_out/r02a_test_new.log; tail -4 gpurun_out/r02a_test_all.log; tail -1 gpurun_out/r02a_smoke.log
cut -c1-600 gpurun_out/r02a_bench_1gpu.json; cat gpurun_out/r02a_setup_n4096.json gpurun_out/r02a_setup_n1000.json
# afterwards, on two GPUs (separate call):
#   gpurun --gpus 2 --timeout 600 -- 'python -m pytest tests/test_gpu_multi.py -m gpu -q'
# (row-sharded instance over peer memory / NCCL, and run_ep_sharded over NCCL, not yet run)
