timeout 600 python -m pytest tests -q -m gpu -k "generic_regime" 2>&1 | grep -E "rel-l2|passed|failed|FAILED" | head -30
