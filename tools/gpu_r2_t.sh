# round-2 GPU session T: racecheck again without the TMA-ring kernels (racecheck does not model the
# mbarrier complete_tx hand-over of cp.async.bulk and reports every ring read as a hazard)
set -x
mkdir -p gpurun_out
cp tools/sanitize_case.py /tmp/sanitize_case.py
NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=2000 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 \
  --kernel-regex-exclude kns=k_jacobi_gram --kernel-regex-exclude kns=k_jacobi_rotate --kernel-regex-exclude kns=k_gemv_tma --kernel-regex-exclude kns=k_dgemm_dmma_tma --kernel-regex-exclude kns=k_jacobi_round_fused \
  python /tmp/sanitize_case.py > gpurun_out/r2t_racecheck.log 2>&1; echo "racecheck (no TMA kernels) rc=$?" > gpurun_out/r2t_status.txt
cat gpurun_out/r2t_status.txt; tail -8 gpurun_out/r2t_racecheck.log | cut -c1-300
