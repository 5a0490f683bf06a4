# weak-scaling point of the final code: bash tools/gpu_r02_scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 --cpu-budget 0 --shim-views 0 --stage-views 0 > gpurun_out/r02_bench_n${N}_G_final.json 2> gpurun_out/r02_bench_n${N}_G_final.err; echo "n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n${N}_G_final.json").read().strip().splitlines()[-1])
    print("   ", round(d["value"],1), "views/s; ms_views", round(d["ms_views"],2), "exchange_ms", round(d["exchange_ms"],3), d["exchange"], d.get("exchange_note"), d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02_bench_n${N}_G_final.err").read()[-2500:])
PY
