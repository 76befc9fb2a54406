export PYTHONUNBUFFERED=1
rm -f gpurun_out/r02_sanitizer.log
for tool in memcheck racecheck synccheck; do
  for c in r1_simt r1m_fused r1m_split r2_group r2_single r3_small gemm tail; do
    echo "=== $tool $c" >> gpurun_out/r02_sanitizer.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_cases.py $c 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case .* done|Error|hazard|=========.*(Invalid|Race|Barrier)" | head -12 >> gpurun_out/r02_sanitizer.log
  done
done
tail -80 gpurun_out/r02_sanitizer.log
