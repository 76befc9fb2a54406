set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generic_regime" 2>&1 | tail -2
