# config D (DINOv2-shaped: 1024-d, 64 x 64 tokens, nearest): full-resolution and encoder-resolution maps
mkdir -p gpurun_out
for f in full lowres; do
  timeout 600 python bench.py --config D --features $f --steps 48 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --pool 4 > gpurun_out/r02_bench_D_$f.json 2> gpurun_out/r02_bench_D_$f.err; echo "D $f rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_D_$f.json").read())
    print("   ", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel", d["roofline"]["kernel"][:20], round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],3), d["clocks"])
    print("   ", [(s["stage"][:8], round(s["ms"],3)) for s in d["roofline"]["stages"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02_bench_D_$f.err").read()[-1500:])
PY
done
