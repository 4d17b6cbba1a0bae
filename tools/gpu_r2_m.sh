# round-2 GPU session M: FP32 tangent in the inner Jacobi (far from convergence), fused path only for one wave of pairs
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1; echo "setup tests rc=$?" > gpurun_out/r2m_status.txt
for b in 1 8 64; do
  timeout 200 python tools/bench_setup.py --batch $b --n 1000 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2m_setup_n1000_b$b.json 2>> gpurun_out/r2m_err.log
done
for b in 16 64; do
  timeout 300 python tools/bench_setup.py --batch $b --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2m_setup_n4096_b$b.json 2>> gpurun_out/r2m_err.log
done
cat gpurun_out/r2m_status.txt; tail -4 gpurun_out/r2m_tests.log; tail -3 gpurun_out/r2m_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_setup_*.json')):
    try:
        d=json.load(open(f)); v=d['variants']['jacobi']; p=v['parts']
        print(f, "B=%d ms/inst=%.2f sweeps=%d sweep_ms[2]=%.2f orthU=%.1e orthV=%.1e resid=%.1e"%(d['B'],v['ms_per_instance'],p['sweeps'],p['sweep_ms'][2],v['orth_U'],v['orth_V'],v['residual']), ["%.0e"%x for x in p['off'][-4:]])
    except Exception as e: print(f,'ERR',e)
PY
