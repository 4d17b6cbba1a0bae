# round-2 GPU session F (2 GPUs): multi-rank parity on two real GPUs (NCCL + peer memory over NVLink),
# bench at N=2 with the row_sharded / shared_w blocks
set -x
mkdir -p gpurun_out
S=gpurun_out/r2f_status.txt; rm -f $S
nvidia-smi -L >> $S
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2f_test_multi.log 2>&1; echo "multi (2 GPUs) rc=$?" >> $S
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench.err; echo "bench 2gpu rc=$?" >> $S
cat $S; tail -5 gpurun_out/r2f_test_multi.log; tail -5 gpurun_out/r2f_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2f_bench_2gpu.json'))
    for k in ('value','e2e','setup','row_sharded','shared_w'):
        print(k, json.dumps(d.get(k))[:1500])
except Exception as e: print('bench parse', e)
PY
