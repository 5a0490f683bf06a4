mkdir -p gpurun_out
timeout 600 python tools/parity_report.py > gpurun_out/r01_parity.txt 2>&1
B="python bench.py --steps 3 --warmup 3 --kernel tc --e2e-steps 0 --cpu-budget 0 --pool 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bp_tc_kernel|fpack_planar_kernel|emit_kernel|project_kernel|compact_kernel" -s 15 -c 5 -f -o gpurun_out/r01_tc_full_v4 $B > gpurun_out/b2.log 2>&1
timeout 900 python bench.py > gpurun_out/r01_bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/bench_ref.err
cat gpurun_out/r01_parity.txt; tail -c 1500 gpurun_out/r01_bench_default.json; tail -c 600 gpurun_out/r01_bench_reference.json; tail -n 3 gpurun_out/bench_default.err gpurun_out/bench_ref.err
