# round-1g GPU session: full GPU suite at HEAD, smoke, published protocol with svd_method="auto"
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r5_status.txt
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r5_test_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/r5_status.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r5_status.txt
timeout 200 python tools/bench_published_protocol.py > gpurun_out/r5_published.log 2>&1; echo "published rc=$?" >> gpurun_out/r5_status.txt
cat gpurun_out/r5_status.txt; tail -4 gpurun_out/r5_test_all.log; tail -1 gpurun_out/r5_smoke.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r01g_published_protocol.json"))
for r in d["rows"]:
    print(r["alpha"], "svd %.1f ms" % (r["gpu_total_s"] * 1e3), "auto %.1f ms" % (r["gpu_auto_total_s"] * 1e3), "x%.0f / x%.0f" % (r["speedup_vs_published"], r["speedup_vs_published_auto"]), r["mse_over_rho"], r["gpu_auto_mse_over_rho"], r["gpu_n_iter"], r["gpu_auto_n_iter"])
PY
