# Re-stamp of the evidence after the last kernel change (intra-batch early exit of finished ALU warps): full GPU suite,
# default + lowres bench lines, launch lists, ncu --set full of the two contraction kernels.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/b1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches_lowres.csv $B --features lowres > gpurun_out/b1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_tc_kernel -s 4 -c 1 -f -o gpurun_out/r02_bp_tc_full $B > gpurun_out/b2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_lr_kernel -s 4 -c 1 -f -o gpurun_out/r02_bp_lr_full $B --features lowres > gpurun_out/b2.log 2>&1
timeout 600 python bench.py --features lowres --steps 96 --cpu-budget 0 --shim-views 0 > gpurun_out/r02_bench_G_lowres.json 2>/dev/null
timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
for f in default G_lowres; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read())
    print("$f", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms", "frac", round(d["roofline"]["frac"],3), "view frac", round(d["roofline"]["view"]["frac"],3), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), d.get("e2e") and round(d["e2e"]["lowres_variant"]["value"],1), d["clocks"])
except Exception as e:
    print("$f failed", e)
PY
done
