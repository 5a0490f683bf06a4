mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_peer_exchange.py -x -q -s 2>&1 | tail -4; done
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
