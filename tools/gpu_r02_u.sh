# bp_lr MMA software pipelining: encoder-resolution parity tests, then lowres bench with stage times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "encoder or lowres or host or ratio" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
B="python bench.py --steps 60 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6"
for cfg in "lowres"; do
  timeout 300 $B --features $cfg > gpurun_out/u_$cfg.json 2> gpurun_out/u_$cfg.err; echo "features=$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/u_$cfg.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4))
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/u_$cfg.err").read()[-1500:])
PY
done
