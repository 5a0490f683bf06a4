# radix sort tuning: look-back window width (experiments build), stage times at config G lowres
mkdir -p gpurun_out
B="python bench.py --steps 20 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 8 --features lowres"
for look in 8 16 32; do
  GWBP_LIB_VARIANT=exp GWBP_SORT_LOOK=$look timeout 300 $B > gpurun_out/o_$look.json 2> gpurun_out/o_$look.err; echo "look=$look rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/o_$look.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms")
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/o_$look.err").read()[-1500:])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/o_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0 --features lowres > gpurun_out/o_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/o_launches.csv | grep -v "at::" | head -14
