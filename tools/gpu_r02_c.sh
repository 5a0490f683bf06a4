# Round-2 GPU call C: full GPU suite, default + lowres bench, launch list of the lowres path
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c_pytest.log; grep "parity config\|parity lowres bilinear D=512" gpurun_out/r02c_pytest.log | cut -c1-260
for f in full lowres; do
timeout 600 python bench.py --features $f --steps 64 --e2e-steps 0 --cpu-budget 0 --pool 4 --shim-views 1 > gpurun_out/r02c_$f.json 2> gpurun_out/r02c_$f.err; echo "bench $f rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_$f.json").read())
    print("$f", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "view frac", round(d["roofline"]["view"]["frac"],3), "launches", d["gpu_launches"], "shim", d.get("shim"))
    for s in d["roofline"]["stages"] or []: print("   ", s["stage"], round(s["ms"],4), round(s.get("frac",0),3))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02c_$f.err").read()[-1500:])
PY
done
B="python bench.py --features lowres --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_lowres_launches.csv $B > gpurun_out/r02c_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02c_lowres_launches.csv 2>/dev/null | head -16
