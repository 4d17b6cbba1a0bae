# round-2 GPU session AA: ncu --set full of the three update kernels of a north-star iteration
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_x_update_chunked|k_z_update_chunked|k_factor_message' -s 6 -c 3 -o gpurun_out/r2aa_updates_full -f python bench.py --steps 1 --warmup 1 --iters 6 --no-cpu-baseline --no-shortcut-modes --setup-instances 0 > gpurun_out/r2aa_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2aa_updates_full.ncu-rep
