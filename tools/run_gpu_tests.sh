cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -40 > gpurun_out/pytest_kernels.log; tail -25 gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_towers_gpu.py -x -q 2>&1 | tail -60 > gpurun_out/pytest_towers.log; tail -40 gpurun_out/pytest_towers.log
