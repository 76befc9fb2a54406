timeout 900 python tools/config_sweep.py gpurun_out/r01_configs.json 2>&1 | grep -E "cfg4|graph_tokens|train_tokens|Error|Traceback" | head
