cd $GRAFT_REPO_ROOT
./build/test_gemm 2>&1 | grep -E "time|failed|FAIL" | head -30
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
