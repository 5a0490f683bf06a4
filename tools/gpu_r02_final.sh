# Round-2 evidence run (1 GPU): full GPU suite, default bench, reference arm, other configs (C, C16, M, Q, lowres),
# launch lists, ncu --set full of the two contraction kernels + project + tile sort.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
for c in C C16 M; do timeout 600 python bench.py --config $c --steps 48 --e2e-steps 0 --cpu-budget 0 --shim-views 0 > gpurun_out/r02_bench_$c.json 2>/dev/null; done
timeout 600 python bench.py --features lowres --steps 96 --cpu-budget 0 --shim-views 0 > gpurun_out/r02_bench_G_lowres.json 2>/dev/null
timeout 600 python bench.py --config M --features lowres --steps 48 --e2e-steps 0 --cpu-budget 0 --shim-views 0 > gpurun_out/r02_bench_M_lowres.json 2>/dev/null
timeout 600 python bench.py --config Q --steps 48 --warmup 3 --cpu-budget 20 > gpurun_out/r02_bench_Q.json 2>/dev/null
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/b1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches_lowres.csv $B --features lowres > gpurun_out/b1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_tc_kernel -s 4 -c 1 -f -o gpurun_out/r02_bp_tc_full $B > gpurun_out/b2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_lr_kernel -s 4 -c 1 -f -o gpurun_out/r02_bp_lr_full $B --features lowres > gpurun_out/b2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:project_pack_kernel -s 4 -c 1 -f -o gpurun_out/r02_project_full $B > gpurun_out/b2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:fpack_planar -s 4 -c 1 -f -o gpurun_out/r02_fpack_full $B > gpurun_out/b2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:radix_onesweep_kernel -s 21 -c 1 -f -o gpurun_out/r02_radix_onesweep_full $B --features lowres > gpurun_out/b2.log 2>&1
timeout 300 python tools/parity_report.py > gpurun_out/r02_parity.txt 2>&1
for f in default C C16 M G_lowres M_lowres; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read())
    print("$f", round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms", "frac", round(d["roofline"]["frac"],3), "view frac", round(d["roofline"]["view"]["frac"],3), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), "launches", d.get("gpu_launches"))
except Exception as e:
    print("$f failed", e)
PY
done
head -c 400 gpurun_out/r02_bench_reference_arm.json; echo; head -c 600 gpurun_out/r02_bench_Q.json; echo; tail -12 gpurun_out/r02_parity.txt
