# round-2 GPU session D: 1 GPU -- two ranks sharing the GPU (multi-rank tests over gloo + CUDA IPC),
# whole GPU suite, bench with the real-W setup block
set -x
mkdir -p gpurun_out
S=gpurun_out/r2d_status.txt; rm -f $S
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2d_test_multi.log 2>&1; echo "multi (1 GPU) rc=$?" >> $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_multi.py > gpurun_out/r2d_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_1gpu.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> $S
cat $S; tail -15 gpurun_out/r2d_test_multi.log; tail -6 gpurun_out/r2d_test_all.log; tail -2 gpurun_out/r2d_smoke.log; tail -5 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2d_bench_1gpu.json'))
    for k in ('value','e2e','setup','e2e_incl_setup','cpu_baseline','parity_vs_oracle_full_size','roofline'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e: print('bench parse', e)
PY
