timeout 600 python -m pytest tests -q -m gpu -x -k "generic_regime or lm or group_g4" 2>&1 | tail -3
for B in 20 512; do python tools/time_lm.py $B; done
