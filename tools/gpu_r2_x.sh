# round-2 GPU session X: rescale computed inside the expansions (no epilogue, no cross-CTA wait), chunked x / z updates -- GPU suite, ncu launch list, benches with and without
set -x
mkdir -p gpurun_out
S=gpurun_out/r2x_status.txt; rm -f $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=8 > gpurun_out/r2x_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?" >> $S
TRB_UPDATE_KERNELS=0 TRB_FUSE_RESCALE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2x_bench_before.json 2> gpurun_out/r2x_bench_before.err; echo "bench before rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2x_bench2.json 2> gpurun_out/r2x_bench2.err; echo "bench2 rc=$?" >> $S
TRB_UPDATE_KERNELS=0 TRB_FUSE_RESCALE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2x_bench_before2.json 2> gpurun_out/r2x_bench_before2.err; echo "bench before2 rc=$?" >> $S
cat $S; tail -15 gpurun_out/r2x_test_all.log
python - <<'PY'
import csv, collections, json, glob
for f in sorted(glob.glob('gpurun_out/r2x_launches*.csv')):
    rows = list(csv.reader(open(f)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if 'Kernel Name' in r: hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try: v = float(d['Metric Value'].replace(',', ''))
            except Exception: continue
            k = d['Kernel Name'][:48]
            agg[k][0] += 1; agg[k][1] += v
    print(f)
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k:48s} {n:4d} {t/n/1000:9.1f} us")
for f in ('gpurun_out/r2x_bench.json', 'gpurun_out/r2x_bench_before.json', 'gpurun_out/r2x_bench2.json', 'gpurun_out/r2x_bench_before2.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks'])
    except Exception as e:
        print('bench parse', f, e)
PY
