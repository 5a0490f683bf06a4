# Hand-written radix sort + scan (no CUB): full GPU parity suite, then bench (full / lowres) with stage times and a launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_pytest.log
B="python bench.py --steps 60 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 6"
for cfg in "full" "lowres"; do
  timeout 300 $B --features $cfg > gpurun_out/t_$cfg.json 2> gpurun_out/t_$cfg.err; echo "features=$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t_$cfg.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4))
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/t_$cfg.err").read()[-1500:])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0 --features lowres > gpurun_out/t_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/t_launches.csv | grep -v "at::" | head -16
