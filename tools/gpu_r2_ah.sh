# round-2 GPU session AH (final round-2c build): what the driver runs at round end (GPU suite, smoke, both bench arms) + ncu launch list + in-pipeline timeline
set -x
mkdir -p gpurun_out
S=gpurun_out/r2ah_status.txt; rm -f $S
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2ah_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ah_smoke.log 2>&1; echo "smoke rc=$?" >> $S
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2ah_bench_reference.json 2> gpurun_out/r2ah_bench_reference.err ) 2> gpurun_out/r2ah_time_reference.txt; echo "reference arm rc=$?" >> $S
( time timeout 900 python bench.py > gpurun_out/r2ah_bench_1gpu.json 2> gpurun_out/r2ah_bench.err ) 2> gpurun_out/r2ah_time_bench.txt; echo "bench rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r2ah_bench_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2ah_ncu_bench.log 2>&1; echo "ncu rc=$?" >> $S
timeout 600 python tools/time_stages.py --out gpurun_out/r2ah_time_stages.json > gpurun_out/r2ah_time_stages.log 2>&1; echo "timeline rc=$?" >> $S
cat $S; tail -4 gpurun_out/r2ah_test_all.log; tail -1 gpurun_out/r2ah_smoke.log; cat gpurun_out/r2ah_time_reference.txt gpurun_out/r2ah_time_bench.txt | grep real; cut -c1-600 gpurun_out/r2ah_bench_reference.json
grep default gpurun_out/r2ah_time_stages.log | tail -1; grep "nine" gpurun_out/r2ah_time_stages.log | tail -1
python - <<'PY'
import json, csv, collections
d=json.loads([l for l in open('gpurun_out/r2ah_bench_1gpu.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','setup','cpu_baseline','clocks','e2e_incl_setup'):
    print(k, json.dumps(d.get(k))[:900])
rows = list(csv.reader(open('gpurun_out/r2ah_bench_launches.csv')))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        dd = dict(zip(hdr, r))
        try: v = float(dd['Metric Value'].replace(',', ''))
        except Exception: continue
        k = dd['Kernel Name'][:48]
        agg[k][0] += 1; agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:48s} {n:4d} {t/n/1000:9.1f} us")
PY
