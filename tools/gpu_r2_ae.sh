# round-2 GPU session AE: compute-sanitizer on the round-2c sweep kernels (memcheck on all of them; racecheck without
# the TMA-ring kernels, whose mbarrier hand-over racecheck does not model)
set -x
mkdir -p gpurun_out
S=gpurun_out/r2ae_status.txt; rm -f $S
cp tools/sanitize_sweep_case.py /tmp/sanitize_sweep_case.py
timeout 300 python /tmp/sanitize_sweep_case.py > gpurun_out/r2ae_plain.log 2>&1; echo "plain rc=$?" >> $S
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/sanitize_sweep_case.py > gpurun_out/r2ae_memcheck.log 2>&1; echo "memcheck rc=$?" >> $S
NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=2000 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --kernel-regex-exclude kns=k_gemv_tma python /tmp/sanitize_sweep_case.py > gpurun_out/r2ae_racecheck.log 2>&1; echo "racecheck (no TMA kernels) rc=$?" >> $S
cat $S; tail -3 gpurun_out/r2ae_plain.log; tail -4 gpurun_out/r2ae_memcheck.log | cut -c1-200; tail -4 gpurun_out/r2ae_racecheck.log | cut -c1-200
