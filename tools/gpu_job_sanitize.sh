export PYTHONUNBUFFERED=1
for tool in memcheck racecheck synccheck; do
  for c in r1_simt r1m_fused r1m_split r2_group r2_single gemm tail; do
    echo "=== $tool $c" >> gpurun_out/r02_sanitizer.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_cases.py $c 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case .* done|Error|hazard|=========.*(Invalid|Race|Barrier)" | head -12 >> gpurun_out/r02_sanitizer.log
  done
done
tail -80 gpurun_out/r02_sanitizer.log
