mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "supertile or encoder" 2>&1 | tail -8
