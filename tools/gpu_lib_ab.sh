# A/B of two builds of the library on the same box: lib/libgwbp_old.so vs lib/libgwbp_new.so
mkdir -p gpurun_out
L=3dgs-gradient-backprojection_b200/lib
for v in old new old new; do cp $L/libgwbp_$v.so $L/libgwbp.so
timeout 600 python bench.py --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 > gpurun_out/ab_$v.json 2> gpurun_out/ab.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read())
print("$v", round(d["value"],1), "views/s; kernel_ms", round(d["roofline"]["kernel_ms"],4))
PY
done
cp $L/libgwbp_new.so $L/libgwbp.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "backprojection or lowres or ratio or autograd or config_G or config_M or layouts" 2>&1 | tail -2
