# round-2 GPU session R: Gram + eigenvectors in one kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_setup.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2r_tests.log 2>&1; echo "setup tests rc=$?" > gpurun_out/r2r_status.txt
for b in 16 64; do
  for m in 3 1; do
    timeout 300 python tools/bench_setup.py --batch $b --n 4096 --alpha 0.5 --skip-svd --skip-gram --fused-mask $m > gpurun_out/r2r_setup_n4096_b${b}_m$m.json 2>> gpurun_out/r2r_err.log
  done
done
for w in 2 8; do
  timeout 300 python tools/bench_setup.py --batch 16 --n 4096 --alpha 0.5 --skip-svd --skip-gram --waves $w > gpurun_out/r2r_setup_n4096_b16_w$w.json 2>> gpurun_out/r2r_err.log
done
timeout 300 python tools/bench_setup.py --batch 64 --n 1000 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2r_setup_n1000_b64.json 2>> gpurun_out/r2r_err.log
cat gpurun_out/r2r_status.txt; tail -5 gpurun_out/r2r_tests.log; tail -3 gpurun_out/r2r_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_setup_*.json')):
    try:
        d=json.load(open(f)); v=d['variants']['jacobi']; p=v['parts']
        print(f, "B=%d ms/inst=%.2f sweeps=%d sweep_ms[2]=%.2f z=%s orthV=%.1e resid=%.1e"%(d['B'],v['ms_per_instance'],p['sweeps'],p['sweep_ms'][2],p['zsplit'],v['orth_V'],v['residual']))
    except Exception as e: print(f,'ERR',e)
PY
