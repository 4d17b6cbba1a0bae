# round-2 GPU session AD: compute-sanitizer on the round-2c sweep kernels (memcheck on all of them; racecheck without
# the TMA-ring kernels, whose mbarrier hand-over racecheck does not model) + ncu launch list + timeline of the final build
set -x
mkdir -p gpurun_out
S=gpurun_out/r2ad_status.txt; rm -f $S
cp tools/sanitize_sweep_case.py /tmp/sanitize_sweep_case.py
timeout 300 python /tmp/sanitize_sweep_case.py > gpurun_out/r2ad_plain.log 2>&1; echo "plain rc=$?" >> $S
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/sanitize_sweep_case.py > gpurun_out/r2ad_memcheck.log 2>&1; echo "memcheck rc=$?" >> $S
NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=2000 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --kernel-regex-exclude kns=k_gemv_tma python /tmp/sanitize_sweep_case.py > gpurun_out/r2ad_racecheck.log 2>&1; echo "racecheck (no TMA kernels) rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/r2ad_bench_launches.csv python bench.py --steps 2 --warmup 1 --iters 10 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2ad_ncu_bench.log 2>&1; echo "ncu rc=$?" >> $S
timeout 600 python tools/time_stages.py --out gpurun_out/r2ad_time_stages.json > gpurun_out/r2ad_time_stages.log 2>&1; echo "timeline rc=$?" >> $S
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_primitives.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2ad_test_api.log 2>&1; echo "api tests rc=$?" >> $S
cat $S; tail -3 gpurun_out/r2ad_plain.log; tail -4 gpurun_out/r2ad_memcheck.log | cut -c1-200; tail -4 gpurun_out/r2ad_racecheck.log | cut -c1-200; tail -3 gpurun_out/r2ad_test_api.log
grep default gpurun_out/r2ad_time_stages.log | tail -1; grep "nine" gpurun_out/r2ad_time_stages.log | tail -1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2ad_bench_launches.csv')))
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
