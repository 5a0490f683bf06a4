# Round-2 multi-GPU evidence (8 x B200, one box): config G at the driver's job size (20 views per rank) and config M,
# closing exchange = reduce-scatter (default); NCCL INFO log kept to show the transport (NVLS / rings).
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=INFO timeout 900 $T bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 4 --shim-views 0 > gpurun_out/r02_bench_n8_G.json 2> gpurun_out/r02_n8_G.err; echo "G rc=$?"
grep -m3 -E "NVLS|Connected all|via P2P|nranks 8" gpurun_out/r02_n8_G.err | cut -c1-200
timeout 900 $T bench.py --gpus 8 --config M --steps 24 --warmup 3 --e2e-steps 0 --shim-views 0 --pool 3 > gpurun_out/r02_bench_n8_M.json 2> gpurun_out/r02_n8_M.err; echo "M rc=$?"
timeout 900 $T bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 0 --shim-views 0 --stage-views 0 --collective allreduce > gpurun_out/r02_bench_n8_G_allreduce.json 2> gpurun_out/r02_n8_Ga.err; echo "G allreduce rc=$?"
timeout 300 $T bench.py --impl reference --gpus 8 --steps 1 --warmup 0 --cpu-budget 10 > gpurun_out/r02_bench_n8_reference.json 2> gpurun_out/r02_n8_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("n8_G","n8_M","n8_G_allreduce","n8_reference"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{f}.json").read())
        print(f, round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],3), "views ms", round(d.get("ms_views",0),2), "exchange ms", round(d.get("exchange_ms",0),2), d.get("exchange"), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), "cores", (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r02_n8_G.err | cut -c1-300
