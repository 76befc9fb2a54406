timeout 600 python -m pytest tests -q -m gpu -k "generic_regime or lm or group_g4" 2>&1 | tail -4
for B in 20 512; do python tools/time_lm.py $B; VMLMF_G_SIMT=1 python tools/time_lm.py $B; done
