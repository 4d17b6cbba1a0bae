# round-2 GPU session Y: in-pipeline duration of every launch of a north-star iteration, per kernel choice
set -x
mkdir -p gpurun_out
timeout 900 python tools/time_stages.py --out gpurun_out/r2y_time_stages.json > gpurun_out/r2y_time_stages.log 2>&1; echo "rc=$?"
cat gpurun_out/r2y_time_stages.log | tail -12
