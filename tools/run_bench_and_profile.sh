# ncu evidence for the N=1 bench command (run under gpurun, one GPU):
#   1. launch list with per-launch duration and DRAM bytes (cold-cache, serialised: shares, not absolutes)
#   2. one --set full capture of the seven kernels of one transformer layer at the bench batch size (256 images)
# bench.py runs 3 warm-up steps + 1 timed + 1 event-profiled step; ~1060 launches per step at this size.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BENCH="python bench.py --classes 32 --queries 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
if [ "$1" != "full-only" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 2000 -c 800 --csv --log-file gpurun_out/launches_v3.csv $BENCH > gpurun_out/ncu_launch_run.log 2>&1
tail -2 gpurun_out/ncu_launch_run.log
fi
# each forward launches 87 matching kernels (12 x (2 LN + 4 GEMM + attention) + ln_pre + patch-embed + ln_post);
# skip two forwards and the first two layers of the third, capture one whole layer (+1)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_pair_kernel|attention_tc_kernel|layernorm_kernel" \
    -s 190 -c 8 -o gpurun_out/prof_layer_v3 -f python tools/profile_layer.py > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log
ls -la gpurun_out | tail -4
