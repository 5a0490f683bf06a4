mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer_exchange.py -x -q -s > gpurun_out/q_pytest.log 2>&1; echo "peer pytest rc=$?"; tail -25 gpurun_out/q_pytest.log
