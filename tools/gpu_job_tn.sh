timeout 300 python tools/test_tn.py 2>&1 | tail -20
