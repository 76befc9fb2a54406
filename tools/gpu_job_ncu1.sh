for tag in "cfg5:2048 128 9 1024 64 64" "lm512:512 35 650 650 300 300" "cfg2r32:8192 24 77 256 32 32"; do
name=${tag%%:*}; args=${tag#*:}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launch_$name.csv python tools/time_r2.py $args 1 > gpurun_out/ncu_$name.log 2>&1
done
