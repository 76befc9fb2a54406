timeout 120 build/test_gemm_tc 2>&1 | tail -30
echo "exit=$?"
