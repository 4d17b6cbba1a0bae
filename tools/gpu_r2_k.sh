# round-2 GPU session K: CUDA-graph replay of launch-bound Jacobi sweeps
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py tests/test_gpu_primitives.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2k_status.txt
for b in 1 2; do
  timeout 200 python tools/bench_setup.py --batch $b --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2k_setup_n4096_b$b.json 2>> gpurun_out/r2k_err.log
  TRB_CUDA_GRAPHS=0 timeout 200 python tools/bench_setup.py --batch $b --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2k_setup_n4096_b${b}_nograph.json 2>> gpurun_out/r2k_err.log
done
for b in 1 8; do
  timeout 200 python tools/bench_setup.py --batch $b --n 1000 --alpha 0.5 --skip-svd > gpurun_out/r2k_setup_n1000_b$b.json 2>> gpurun_out/r2k_err.log
  TRB_CUDA_GRAPHS=0 timeout 200 python tools/bench_setup.py --batch $b --n 1000 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2k_setup_n1000_b${b}_nograph.json 2>> gpurun_out/r2k_err.log
done
timeout 600 python tools/bench_published_protocol.py > gpurun_out/r2k_published.log 2>&1; echo "published rc=$?" >> gpurun_out/r2k_status.txt
cat gpurun_out/r2k_status.txt; tail -4 gpurun_out/r2k_tests.log; tail -3 gpurun_out/r2k_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_setup_*.json')):
    try:
        d=json.load(open(f)); v=d['variants']['jacobi']; p=v['parts']
        lib=d['variants'].get('gram_cusolver_eigh',{}).get('ms_per_instance')
        print(f, "B=%d ms/inst=%.2f sweeps=%d sweep_ms[2]=%.2f lib=%s"%(d['B'],v['ms_per_instance'],p['sweeps'],p['sweep_ms'][2],lib))
    except Exception as e: print(f,'ERR',e)
d=json.load(open('gpurun_out/r02_published_protocol.json'))
for r in d['rows']: print(r['alpha'], 'auto svd %.1f ms total %.1f ms | gram_eigh svd %s | gesvd svd %.1f'%(1e3*r['gpu_auto_svd_s'],1e3*r['gpu_auto_total_s'], ('%.1f'%(1e3*r['gpu_gram_eigh_svd_s'])) if 'gpu_gram_eigh_svd_s' in r else '-', 1e3*r['gpu_svd_s']))
PY
