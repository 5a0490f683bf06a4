# Round-2 GPU call D: ncu --set full of the two contraction kernels, config Q line, bp_lr without reductions (experiment build)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_tc_kernel -s 4 -c 1 -f -o gpurun_out/r02_bp_tc_full $B > gpurun_out/r02d_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bp_lr_kernel -s 4 -c 1 -f -o gpurun_out/r02_bp_lr_full $B --features lowres > gpurun_out/r02d_ncu2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/r02d_ncu3.log 2>&1
timeout 600 python bench.py --config Q --steps 24 --warmup 3 --cpu-budget 20 > gpurun_out/r02_bench_Q.json 2> gpurun_out/r02d_Q.err; echo "Q rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_Q.json").read())
    print("Q", round(d["value"],1), d["unit"], "linear", round(d["linear_path"]["value"],1), "e2e", round(d["e2e"]["value"],1), "render kernel ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],3), "parity", d["parity"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02d_Q.err").read()[-1500:])
PY
for dbg in 0 1; do
GWBP_LIB_VARIANT=exp GWBP_LR_DEBUG=$dbg timeout 600 python bench.py --features lowres --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 --shim-views 0 --stage-views 0 > gpurun_out/r02d_lr_dbg$dbg.json 2> gpurun_out/r02d_lr.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d_lr_dbg$dbg.json").read())
    print("lr debug=$dbg", round(d["value"],1), "views/s; kernel_ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02d_lr.err").read()[-800:])
PY
done
