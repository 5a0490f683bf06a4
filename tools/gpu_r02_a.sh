# Round-2 GPU call A: gsplat attempt, full GPU test suite (with parity prints), short bench, launch list.
mkdir -p gpurun_out
#bash tools/gsplat_install_attempt.sh
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log; grep "^\[parity" gpurun_out/r02a_pytest.log | head -40
timeout 600 python bench.py --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 --shim-views 1 > gpurun_out/r02a_quick.json 2> gpurun_out/r02a_quick.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02a_quick.json").read())
    print(round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "launches", d.get("gpu_launches"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02a_quick.err").read()[-1500:])
PY
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv $B > gpurun_out/r02a_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02a_launches.csv 2>/dev/null | head -30
