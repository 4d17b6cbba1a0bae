# round-2 GPU session AN: a second box for the bench line of the final build
set -x
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/r2an_bench_1gpu.json 2> gpurun_out/r2an_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2an_bench_1gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks'], d['setup']['ms_per_instance'], d['e2e_incl_setup']['value'])
PY
