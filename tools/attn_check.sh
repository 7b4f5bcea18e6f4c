cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -15
timeout 300 python tools/bench_kernels.py 2>&1 | tail -4
