# ncu --set full of one radix pass (u32 keys) + the supertile pass
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2 --stage-views 0 --shim-views 0 --features lowres"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep_kernel -s 21 -c 2 -f -o gpurun_out/p_onesweep $B > gpurun_out/p_ncu1.log 2>&1; echo "ncu1 rc=$?"
