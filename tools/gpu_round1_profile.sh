# Final round-1 evidence run (1 GPU): default bench, reference arm, launch list, full ncu capture of the hot kernels.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r01_bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/bench_ref.err
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-budget 0 --pool 2"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv $B > gpurun_out/b1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bp_tc_kernel|fpack_planar_kernel|emit_kernel|project_kernel|compact_kernel" -s 15 -c 5 -f -o gpurun_out/r01_final_full $B > gpurun_out/b2.log 2>&1
tail -c 2500 gpurun_out/r01_bench_default.json; echo; tail -c 400 gpurun_out/r01_bench_reference.json; tail -n 2 gpurun_out/bench_default.err
