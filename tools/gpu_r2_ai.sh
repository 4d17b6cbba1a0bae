# round-2 GPU session AI (4 GPUs, final round-2c build): the bench line at N = 4 (the driver's scaling run visits 1, 2, 4, 8)
set -x
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2ai_bench_4gpu.json 2> gpurun_out/r2ai_bench.err; echo "bench 4gpu rc=$?" > gpurun_out/r2ai_status.txt
cat gpurun_out/r2ai_status.txt; tail -3 gpurun_out/r2ai_bench.err
python - <<'PY'
import json
try:
    txt=open('gpurun_out/r2ai_bench_4gpu.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(d['value'], d['n_gpus'], d['e2e']['value'])
    print({k:d['row_sharded'][k] for k in ('max_rel_dev_sharded_vs_unsharded','speedup_vs_one_gpu')}, d['row_sharded']['sharded']['ms_per_iter'], d['row_sharded']['sharded']['frac_of_hbm_peak'])
    print(d['shared_w']['value'], d['shared_w']['tflops_per_gpu'], d['shared_w']['max_rel_dev_dmma_vs_gemv'])
except Exception as e: print('bench parse', e)
PY
