# Supertile binning + lister warp: parity subset, then bench (full / lowres), per-tile lists for reference.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "supertile" > gpurun_out/k_pytest0.log 2>&1; echo "supertile pytest rc=$?"; tail -4 gpurun_out/k_pytest0.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "integer or backprojection or lowres or encoder or culling or edge or known or rasterization or config_G or forward or ratio or host" > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/k_pytest.log
B="python bench.py --steps 60 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6"
for cfg in "full" "lowres"; do
  timeout 300 $B --features $cfg > gpurun_out/k.json 2> gpurun_out/k.err; echo "features=$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/k.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4))
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/k.err").read()[-1500:])
PY
done
