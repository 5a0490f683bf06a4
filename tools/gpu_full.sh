# Full GPU test suite + a short default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/full_pytest.log
timeout 600 python bench.py --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 > gpurun_out/quick.json 2> gpurun_out/quick.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/quick.json").read())
    print(round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("failed", e); print(open("gpurun_out/quick.err").read()[-1500:])
PY
