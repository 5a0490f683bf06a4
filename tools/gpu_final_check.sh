# Last check of the committed state: full GPU suite, smoke(), default bench (short CPU leg).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r01_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r01_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --cpu-budget 8 > gpurun_out/r01_bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r01_bench_default.json").read())
print(round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["lowres_variant"]["value"],1), d["clocks"])
PY
