timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "spectral or time_step" 2>&1 | tail -15
