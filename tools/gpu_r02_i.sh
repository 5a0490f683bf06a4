mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:project_pack -s 4 -c 1 -f -o gpurun_out/r02_project_pack_full $B > gpurun_out/i.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
