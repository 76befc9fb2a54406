python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
python tools/prof_cfg1.py 64 20
python tools/prof_cfg1.py 8192 20
python tools/config_sweep.py gpurun_out/configs_tail.json > /dev/null 2> gpurun_out/configs_tail.err; tail -3 gpurun_out/configs_tail.err
