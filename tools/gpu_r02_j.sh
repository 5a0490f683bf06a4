mkdir -p gpurun_out
B="python bench.py --steps 40 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6 --features lowres"
for k in 1 2 4 8; do
  GWBP_LOOK=$k timeout 300 $B > gpurun_out/j.json 2> gpurun_out/j.err; echo "look=$k rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/j.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms;", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]][:2])
except Exception as e:
    print("failed", e); print(open("gpurun_out/j.err").read()[-1500:])
PY
done
