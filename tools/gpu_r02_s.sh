# 8 GPUs: peer-memory closing exchange, config G (20 views per rank, the driver's scaling shape) and config M (24 per rank)
mkdir -p gpurun_out
run() { # name config steps extra
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --config $2 --steps $3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 0 $4 > gpurun_out/s_$1.json 2> gpurun_out/s_$1.err; echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s_$1.json").read().strip().splitlines()[-1])
    print("   ", round(d["value"],1), "views/s; ms_views", round(d["ms_views"],2), "exchange_ms", round(d["exchange_ms"],3), d["exchange"], d.get("exchange_note"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/s_$1.err").read()[-2500:])
PY
}
run n8_G_peer G 20 "--collective auto"
run n8_M_peer M 24 "--collective auto --pool 4"
