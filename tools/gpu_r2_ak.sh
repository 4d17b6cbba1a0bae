# round-2 GPU session AK: ncu --set full of the projecting GEMV with the rescale inside (k_gemv_tma<NB, 2>) and of the plain expansion
set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemv_tma -s 8 -c 4 -o gpurun_out/r2ak_gemv_full -f python bench.py --steps 1 --warmup 1 --iters 6 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2ak_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2ak_gemv_full.ncu-rep --page raw --csv > gpurun_out/r2ak_gemv_raw.csv 2>/dev/null; ls -la gpurun_out/r2ak_gemv_raw.csv
