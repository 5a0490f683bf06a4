# Fused front end (project_pack + emit with chained scans): bit-exact tests, parity subset, bench with / without pack overlap.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "integer or backprojection or lowres or culling or edge or known or rasterization or config_G or forward" > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/h_pytest.log
B="python bench.py --steps 60 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6"
for cfg in "0 full" "1 full" "0 lowres"; do
  set -- $cfg
  timeout 300 $B --overlap-pack $1 --features $2 > gpurun_out/h.json 2> gpurun_out/h.err; echo "overlap=$1 features=$2 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/h.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4))
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/h.err").read()[-1500:])
PY
done
