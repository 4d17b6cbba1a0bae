# round-2 GPU session V: rescale folded into the projections, chunked update kernels -- GPU suite, ncu launch lists, short bench
set -x
mkdir -p gpurun_out
S=gpurun_out/r2v_status.txt; rm -f $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=8 > gpurun_out/r2v_test_all.log 2>&1; echo "all tests rc=$?" >> $S
for MASK in 7 3 0; do
TRB_UPDATE_KERNELS=$MASK timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2v_launches_mask$MASK.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2v_ncu_mask$MASK.log 2>&1; echo "ncu mask $MASK rc=$?" >> $S
done
TRB_FUSE_RESCALE=0 TRB_UPDATE_KERNELS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2v_launches_before.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2v_ncu_before.log 2>&1; echo "ncu before rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?" >> $S
cat $S; tail -15 gpurun_out/r2v_test_all.log
python - <<'PY'
import csv, collections, json, glob
for f in sorted(glob.glob('gpurun_out/r2v_launches_*.csv')):
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
try:
    d = json.loads([l for l in open('gpurun_out/r2v_bench.json') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline'])
except Exception as e:
    print('bench parse', e)
PY
