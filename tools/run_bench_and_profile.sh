# Full bench (N=1) + ncu launch list + one full ncu capture of the top kernel. Run under gpurun.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 3000 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
# launch list: small workload (same kernels, same shapes per batch), serialised cold-cache timings
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --classes 32 --queries 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
# full capture of the GEMM kernel (3 launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 200 -c 4 \
    -o gpurun_out/prof_gemm -f python bench.py --classes 32 --queries 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out | tail -12
