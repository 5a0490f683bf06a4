# Experiment: persistent feature re-layout (k CTAs per SM) on a second stream next to projection + binning of the same view.
mkdir -p gpurun_out
B="python bench.py --steps 60 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6"
for cfg in "0 3 0" "1 1 1" "1 2 1" "1 3 1" "1 2 0"; do
  set -- $cfg
  GWBP_FPACK_CTAS=$2 GWBP_OVERLAP_SWAP=$3 timeout 300 $B --overlap-pack $1 > gpurun_out/ov.json 2> gpurun_out/ov.err; echo "overlap=$1 ctas=$2 pack_high_prio=$3 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ov.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4))
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/ov.err").read()[-1500:])
PY
done
