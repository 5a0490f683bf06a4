# quick A/B style check: parity subset + full bench (+ optional lowres)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "backprojection or encoder_resolution or full_size or layouts or flip" > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02f_pytest.log
for f in $FEATS; do
timeout 600 python bench.py --features $f --steps 64 --e2e-steps 0 --cpu-budget 0 --pool 4 --shim-views 0 > gpurun_out/r02f_$f.json 2> gpurun_out/r02f_$f.err; echo "bench $f rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02f_$f.json").read())
    print("$f", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "view frac", round(d["roofline"]["view"]["frac"],3))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02f_$f.err").read()[-1500:])
PY
done
