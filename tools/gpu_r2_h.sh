# round-2 GPU session H: set-up timing for small groups of instances (does an L2-resident group pay?),
# ncu --set full of the product DMMA GEMM and of the three set-up kernels, the GPU suite with the new tests
set -x
mkdir -p gpurun_out
for b in 1 2 3 4 6 8 32; do
  timeout 200 python tools/bench_setup.py --batch $b --n 4096 --alpha 0.5 --skip-svd --skip-gram > gpurun_out/r2h_setup_b$b.json 2>> gpurun_out/r2h_err.log
done
for w in 2 8 16; do
  timeout 200 python tools/bench_setup.py --batch 3 --n 4096 --alpha 0.5 --skip-svd --skip-gram --waves $w > gpurun_out/r2h_setup_b3_w$w.json 2>> gpurun_out/r2h_err.log
done
BENCH_GEMM_NO_PROBES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dgemm_dmma_tma -s 2 -c 2 -o gpurun_out/r2h_gemm_full -f python tools/bench_gemm.py 16384 1024 1 > gpurun_out/r2h_gemm_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobi -s 30 -c 3 -o gpurun_out/r2h_setup_full -f python tools/profile_setup.py --batch 16 > gpurun_out/r2h_setup_prof.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2h_test_all.log 2>&1; echo "all tests rc=$?" > gpurun_out/r2h_status.txt
cat gpurun_out/r2h_status.txt; tail -5 gpurun_out/r2h_test_all.log; tail -3 gpurun_out/r2h_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_setup_b*.json')):
    try:
        d=json.load(open(f)); v=d['variants']['jacobi']; p=v['parts']
        print(f, "B=%d ms/inst=%.2f sweeps=%d sweep_ms[2]=%.2f per_inst_sweep=%.2f z=%s"%(d['B'],v['ms_per_instance'],p['sweeps'],p['sweep_ms'][2],p['sweep_ms'][2]/d['B'],p['zsplit']))
    except Exception as e: print(f,'ERR',e)
PY
