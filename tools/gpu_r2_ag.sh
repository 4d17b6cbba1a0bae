# round-2 GPU session AG: chunked kernels as 128 threads x 8 elements -- timeline, ncu launch list, API tests
set -x
mkdir -p gpurun_out
timeout 600 python tools/time_stages.py --out gpurun_out/r2ag_time_stages.json > gpurun_out/r2ag_time_stages.log 2>&1; echo "timeline rc=$?"
grep default gpurun_out/r2ag_time_stages.log; grep "nine" gpurun_out/r2ag_time_stages.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_x_update|k_z_update|k_factor' -c 90 --csv --log-file gpurun_out/r2ag_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2ag_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2ag_launches.csv')))
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
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_multi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2ag_test_api.log 2>&1; echo "api tests rc=$?"; tail -3 gpurun_out/r2ag_test_api.log
