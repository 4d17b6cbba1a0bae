# round-2 GPU session I (8 GPUs): the bench line at N = 8 with the row_sharded / shared_w blocks
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2i_status.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2i_bench_8gpu.json 2> gpurun_out/r2i_bench.err; echo "bench 8gpu rc=$?" >> gpurun_out/r2i_status.txt
cat gpurun_out/r2i_status.txt; tail -5 gpurun_out/r2i_bench.err
python - <<'PY'
import json
try:
    txt=open('gpurun_out/r2i_bench_8gpu.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    for k in ('value','e2e','setup','row_sharded','shared_w'):
        print(k, json.dumps(d.get(k))[:1600])
except Exception as e: print('bench parse', e)
PY
