timeout 900 python tools/config_sweep.py gpurun_out/r01_configs.json 2>&1 | grep -E "config|train_seq|graph_seq|tokens|Error|error|Traceback" | head -40
