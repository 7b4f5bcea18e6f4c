# GPU regression: all parity tests, kernel micro-timings, full N=1 bench. Run under gpurun.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_kernels.py 2>&1 | tail -12 | tee gpurun_out/bk.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 2500 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
