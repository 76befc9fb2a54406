python -m pytest tests -q -m gpu -x -k "tensor_core_linear or lm" 2>&1 | tail -4
for B in 20 512; do python tools/time_lm.py $B; done
