# final default bench line (traffic / tensor-pipe share now stamped for these sources) -- 1 GPU
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_default.json").read())
r=d["roofline"]
print(round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), round(d["e2e"]["lowres_variant"]["value"],1), "clocks", d["clocks"], "launches", d["gpu_launches"])
print("roofline", r["frac"], r["traffic"], r["tensor_pipe_pct"], r["traffic_source"])
print([(s["stage"][:8], round(s["ms"],3), s.get("frac") and round(s["frac"],3)) for s in r["stages"]])
PY
