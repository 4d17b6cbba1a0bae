# round-1f GPU session: grid and cluster variants of the persistent sweep, full suite
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r4_status.txt
timeout 300 python -m pytest tests/test_gpu_api.py -q -p no:cacheprovider -k "persistent" > gpurun_out/r4_test_persistent.log 2>&1; echo "persistent tests rc=$?" >> gpurun_out/r4_status.txt
timeout 200 python tools/bench_small_configs.py > gpurun_out/r4_small.log 2>&1; echo "small configs rc=$?" >> gpurun_out/r4_status.txt
cp gpurun_out/r01_small_configs.json gpurun_out/r01f_small_configs.json
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not persistent" > gpurun_out/r4_test_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/r4_status.txt
cat gpurun_out/r4_status.txt; tail -12 gpurun_out/r4_test_persistent.log; tail -3 gpurun_out/r4_test_all.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r01f_small_configs.json"))
for k, v in d.items():
    print(k, {kk: (round(vv["us_per_iter"], 1), round(vv.get("us_per_iter_marginal", 0), 1)) for kk, vv in v.items() if isinstance(vv, dict) and "us_per_iter" in vv}, v["max_rel_dev_vs_oracle"])
PY
