# Remaining evidence with the final sources: parity report, configs C / C16 / M / M lowres / D / Q
mkdir -p gpurun_out
timeout 300 python tools/parity_report.py > gpurun_out/r02_parity.txt 2>&1; tail -4 gpurun_out/r02_parity.txt | cut -c1-220
for c in C C16 M; do timeout 600 python bench.py --config $c --steps 48 --e2e-steps 0 --cpu-budget 0 --shim-views 0 > gpurun_out/r02_bench_$c.json 2>/dev/null; done
timeout 600 python bench.py --config M --features lowres --steps 48 --e2e-steps 0 --cpu-budget 0 --shim-views 0 > gpurun_out/r02_bench_M_lowres.json 2>/dev/null
for f in full lowres; do timeout 600 python bench.py --config D --features $f --steps 48 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --pool 4 > gpurun_out/r02_bench_D_$f.json 2>/dev/null; done
timeout 600 python bench.py --config Q --steps 48 --warmup 3 --cpu-budget 20 > gpurun_out/r02_bench_Q.json 2>/dev/null
for f in C C16 M M_lowres D_full D_lowres; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read())
    print("$f", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms", "frac", round(d["roofline"]["frac"],3), "view frac", round(d["roofline"]["view"]["frac"],3))
except Exception as e:
    print("$f failed", e)
PY
done
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_Q.json").read())
print("Q", round(d["value"],1), "linear", round(d["linear_path"]["value"],1), "render ms", round(d["roofline"]["kernel_ms"],3), d["parity"])
PY
