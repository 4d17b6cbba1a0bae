# round-2 GPU session L: one-kernel-per-round Jacobi for short rows
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1; echo "setup tests rc=$?" > gpurun_out/r2l_status.txt
for b in 1 8 64; do
  timeout 200 python tools/bench_setup.py --batch $b --n 1000 --alpha 0.5 --skip-svd > gpurun_out/r2l_setup_n1000_b$b.json 2>> gpurun_out/r2l_err.log
  timeout 200 python tools/bench_setup.py --batch $b --n 1000 --alpha 0.5 --skip-svd --skip-gram --no-fused > gpurun_out/r2l_setup_n1000_b${b}_nofused.json 2>> gpurun_out/r2l_err.log
done
timeout 200 python tools/bench_setup.py --batch 64 --n 512 --alpha 0.5 --skip-svd > gpurun_out/r2l_setup_n512_b64.json 2>> gpurun_out/r2l_err.log
timeout 600 python tools/bench_published_protocol.py > gpurun_out/r2l_published.log 2>&1; echo "published rc=$?" >> gpurun_out/r2l_status.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_test_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/r2l_status.txt
cat gpurun_out/r2l_status.txt; tail -6 gpurun_out/r2l_tests.log; tail -3 gpurun_out/r2l_err.log; tail -4 gpurun_out/r2l_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_setup_*.json')):
    try:
        d=json.load(open(f)); v=d['variants']['jacobi']; p=v['parts']
        lib=d['variants'].get('gram_cusolver_eigh',{}).get('ms_per_instance')
        print(f, "B=%d ms/inst=%.2f sweeps=%d sweep_ms[2]=%.2f lib=%s orthV=%.1e"%(d['B'],v['ms_per_instance'],p['sweeps'],p['sweep_ms'][2],lib,v['orth_V']))
    except Exception as e: print(f,'ERR',e)
d=json.load(open('gpurun_out/r02_published_protocol.json'))
for r in d['rows']: print(r['alpha'], 'auto svd %.1f ms total %.1f ms | gram_eigh svd %s | gesvd svd %.1f'%(1e3*r['gpu_auto_svd_s'],1e3*r['gpu_auto_total_s'], ('%.1f'%(1e3*r['gpu_gram_eigh_svd_s'])) if 'gpu_gram_eigh_svd_s' in r else '-', 1e3*r['gpu_svd_s']))
PY
