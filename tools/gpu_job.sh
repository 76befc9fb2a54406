timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 120 python tools/time_fwd.py 9472 24 77 256 8 6 2>&1 | grep -E "^train"
VMLMF_BWD_SPLIT=1 timeout 120 python tools/time_fwd.py 9472 24 77 256 8 6 2>&1 | grep -E "^train"
