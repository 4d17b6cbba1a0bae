# round-2 GPU session N: where does the fused round kernel spend its 36 us?  (one 500 x 500 Gram matrix)
set -x
mkdir -p gpurun_out
TRB_CUDA_GRAPHS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_round_fused -s 40 -c 2 -o gpurun_out/r2n_fused_full -f python tools/profile_setup.py --batch 1 --m 500 --n 1000 --sweeps 3 > gpurun_out/r2n_prof.log 2>&1
TRB_CUDA_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg --clock-control none -k regex:k_jacobi -s 10 -c 40 --csv --log-file gpurun_out/r2n_launches.csv python tools/profile_setup.py --batch 1 --m 500 --n 1000 --sweeps 3 > gpurun_out/r2n_prof2.log 2>&1
tail -2 gpurun_out/r2n_prof.log; tail -5 gpurun_out/r2n_launches.csv | cut -c1-300
