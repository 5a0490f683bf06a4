# Quick GPU check used during kernel work: tcgen05 parity subset + a short default bench (48 views, no e2e / CPU legs).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "backprojection or lowres or ratio or autograd or config_G" > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/quick_pytest.log
timeout 600 python bench.py --steps 48 --e2e-steps 0 --cpu-budget 0 $QUICK_ARGS > gpurun_out/quick.json 2> gpurun_out/quick.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/quick.json").read())
    print(round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("failed", e); print(open("gpurun_out/quick.err").read()[-1500:])
PY
