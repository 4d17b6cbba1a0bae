# round-2 GPU session B: ncu of the three set-up kernels (launch list + one full capture each)
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed_pipe_fp64.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_jacobi -s 30 -c 45 --csv --log-file gpurun_out/r2b_setup_launches.csv python tools/profile_setup.py --batch 16 > gpurun_out/r2b_prof1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jacobi -s 30 -c 3 -o gpurun_out/r2b_setup_full -f python tools/profile_setup.py --batch 16 > gpurun_out/r2b_prof2.log 2>&1
ls -la gpurun_out/ | tail -5; tail -3 gpurun_out/r2b_prof1.log gpurun_out/r2b_prof2.log
