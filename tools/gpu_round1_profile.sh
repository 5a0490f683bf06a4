# Final round-1 evidence run (1 GPU): tests, default bench, reference arm, other configs, launch list, ncu full capture of
# the dominant kernel, query bench, parity report.  Everything lands in gpurun_out/; summaries are copied to profiles/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r01_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r01_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r01_bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2> gpurun_out/bench_ref.err
for c in C C16 M; do timeout 600 python bench.py --config $c --steps 48 --e2e-steps 0 --cpu-budget 0 > gpurun_out/r01_bench_$c.json 2>/dev/null; done
timeout 600 python bench.py --features lowres --steps 48 --cpu-budget 0 > gpurun_out/r01_bench_G_lowres.json 2>/dev/null
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv $B > gpurun_out/b1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_tc_kernel -s 4 -c 1 -f -o gpurun_out/r01_final_full $B > gpurun_out/b2.log 2>&1
timeout 300 python tools/query_bench.py 8 2>/dev/null | tail -1 > gpurun_out/r01_query_bench.json
timeout 300 python tools/parity_report.py > gpurun_out/r01_parity.txt 2>&1
for f in default C C16 M G_lowres; do python - <<PY
import json
d=json.loads(open("gpurun_out/r01_bench_$f.json").read())
print("$f", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms", "frac", round(d["roofline"]["frac"],3), "e2e", d.get("e2e") and round(d["e2e"]["value"],1))
PY
done
cat gpurun_out/r01_bench_reference_arm.json | head -c 300; echo; cat gpurun_out/r01_query_bench.json; tail -12 gpurun_out/r01_parity.txt
