# Round-2 GPU call B: lowres adjoint kernel tests + ncu of the binning kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "encoder_resolution or sh or lowres" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.log; grep "^\[parity" gpurun_out/r02b_pytest.log | cut -c1-220 | head -30
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bin_count_kernel -s 3 -c 1 -f -o gpurun_out/r02b_bin_count $B > gpurun_out/r02b_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bin_scatter_kernel -s 3 -c 1 -f -o gpurun_out/r02b_bin_scatter $B > gpurun_out/r02b_ncu2.log 2>&1
timeout 600 python bench.py --features lowres --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 --shim-views 0 > gpurun_out/r02b_lowres.json 2> gpurun_out/r02b_lowres.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b_lowres.json").read())
    print(round(d["value"],1), "views/s", round(d["ms_per_step"],3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"],4))
    for s in d["roofline"]["stages"] or []: print(s["stage"], round(s["ms"],4))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02b_lowres.err").read()[-1500:])
PY
