# Render check: full GPU test suite, render kernel timings by D, config-Q bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/full_pytest.log
for d in 64 128 256 512 768 1024; do timeout 300 python tools/render_bench.py $d 8 2>/dev/null | head -1; done
GWBP_RENDER_DEBUG=8 timeout 300 python tools/render_bench.py 512 8 2>/dev/null | head -1
GWBP_RENDER_PROF=1 timeout 300 python tools/render_bench.py 512 5 2>/dev/null | tail -4
timeout 300 python tools/query_bench.py 8 2>/dev/null | tail -1 > gpurun_out/r01_query_bench.json; cat gpurun_out/r01_query_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
