# intra-batch early exit of finished ALU warps: parity subset, then bench (full / lowres)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "supertile or backprojection or encoder or lowres or culling or config_G or flip or ratio or host or dims or against" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/y_pytest.log
B="python bench.py --steps 60 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6"
for cfg in "full" "lowres"; do
  timeout 300 $B --features $cfg > gpurun_out/y_$cfg.json 2> gpurun_out/y_$cfg.err; echo "features=$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/y_$cfg.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), d["clocks"])
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/y_$cfg.err").read()[-1500:])
PY
done
