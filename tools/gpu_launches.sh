mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv $B > gpurun_out/b1.log 2>&1; echo rc=$?
