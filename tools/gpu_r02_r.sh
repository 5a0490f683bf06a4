# 2 GPUs: closing exchange variants at config G, 20 views per rank (the driver's scaling shape)
mkdir -p gpurun_out
for coll in peer reduce_scatter; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 0 --collective $coll > gpurun_out/r_n2_$coll.json 2> gpurun_out/r_n2_$coll.err; echo "n2 $coll rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r_n2_$coll.json").read().strip().splitlines()[-1])
    print("   ", round(d["value"],1), "views/s; ms_views", round(d["ms_views"],2), "exchange_ms", round(d["exchange_ms"],3), d["exchange"], d.get("exchange_note"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r_n2_$coll.err").read()[-2500:])
PY
done
