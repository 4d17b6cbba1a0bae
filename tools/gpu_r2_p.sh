# round-2 GPU session P: what the driver runs at round end -- GPU suite, smoke, bench (both arms), plus the ncu launch list of the bench
set -x
mkdir -p gpurun_out
S=gpurun_out/r2p_status.txt; rm -f $S
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2p_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?" >> $S
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err ) 2> gpurun_out/r2p_time_reference.txt; echo "reference arm rc=$?" >> $S
( time timeout 900 python bench.py > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench.err ) 2> gpurun_out/r2p_time_bench.txt; echo "bench rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r2p_bench_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2p_ncu_bench.log 2>&1; echo "ncu rc=$?" >> $S
cat $S; tail -4 gpurun_out/r2p_test_all.log; tail -1 gpurun_out/r2p_smoke.log; cat gpurun_out/r2p_time_reference.txt gpurun_out/r2p_time_bench.txt | grep real; cut -c1-700 gpurun_out/r2p_bench_reference.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench_1gpu.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','setup','cpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:1200])
PY
