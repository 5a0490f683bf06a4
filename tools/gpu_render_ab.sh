# A/B of the weight-cache render variant (GWBP_RENDER_WCACHE=1) against the default tcgen05 render.
mkdir -p gpurun_out
GWBP_RENDER_WCACHE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "render or masks or config_G_properties or probe" > gpurun_out/quick_pytest.log 2>&1; echo "pytest(wcache) rc=$?"; tail -3 gpurun_out/quick_pytest.log
for wc in 0 1 0 1; do echo "wcache=$wc"; GWBP_RENDER_WCACHE=$wc timeout 300 python tools/render_bench.py 512 10 2>/dev/null | head -1; done
GWBP_RENDER_WCACHE=1 GWBP_RENDER_PROF=1 timeout 300 python tools/render_bench.py 512 5 2>/dev/null | tail -4
for d in 256 1024; do GWBP_RENDER_WCACHE=1 timeout 300 python tools/render_bench.py $d 5 2>/dev/null | head -1; done
