# same-box A/B of two library builds (lib/libgwbp_old.so = before the intra-batch early exit, libgwbp_new.so = after), full and lowres
mkdir -p gpurun_out
L=3dgs-gradient-backprojection_b200/lib
for v in old new old new; do cp $L/libgwbp_$v.so $L/libgwbp.so
 for f in full lowres; do
  timeout 600 python bench.py --features $f --steps 48 --e2e-steps 0 --cpu-budget 0 --pool 4 --shim-views 0 --stage-views 0 > gpurun_out/ab_${v}_$f.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${v}_$f.json").read())
print("$v $f", round(d["value"],1), "views/s; kernel_ms", round(d["roofline"]["kernel_ms"],4), d["clocks"]["sm_mhz"])
PY
 done
done
cp $L/libgwbp_new.so $L/libgwbp.so
