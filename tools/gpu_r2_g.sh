# round-2 GPU session G: whole GPU suite (batched adaptive damping, noise-row rule), parity reports
set -x
mkdir -p gpurun_out
S=gpurun_out/r2g_status.txt; rm -f $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2g_test_all.log 2>&1; echo "all tests rc=$?" >> $S
timeout 300 python tools/parity_report.py > gpurun_out/r2g_parity_report.log 2>&1; echo "parity report rc=$?" >> $S
timeout 900 python tools/parity_full_size.py config4 > gpurun_out/r2g_parity_c4.log 2>&1; echo "parity config4 rc=$?" >> $S
cp gpurun_out/r02_parity_full_size.json gpurun_out/r02_parity_full_size_config4.json
timeout 1500 python tools/parity_full_size.py config3 > gpurun_out/r2g_parity_c3.log 2>&1; echo "parity config3 rc=$?" >> $S
cp gpurun_out/r02_parity_full_size.json gpurun_out/r02_parity_full_size_config3.json
cat $S; tail -6 gpurun_out/r2g_test_all.log; tail -4 gpurun_out/r2g_parity_c4.log | cut -c1-1500; tail -4 gpurun_out/r2g_parity_c3.log | cut -c1-1500
