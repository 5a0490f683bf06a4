# A/B helper: VAR=<env var> VALS="a b c" bash tools/gpu_sweep.sh  -> short default bench per value (same box, back to back)
mkdir -p gpurun_out
for v in $VALS; do
env $VAR=$v timeout 600 python bench.py --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 > gpurun_out/sweep_$v.json 2> gpurun_out/sweep.err
python - <<PY
import json
d=json.loads(open("gpurun_out/sweep_$v.json").read())
print("$VAR=$v", round(d["value"],1), "views/s; kernel_ms", round(d["roofline"]["kernel_ms"],4))
PY
done
