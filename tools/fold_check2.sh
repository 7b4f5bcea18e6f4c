cd $GRAFT_REPO_ROOT
./build/test_gemm 2>&1 | grep -E "FAIL|failed|time" | tail -8
