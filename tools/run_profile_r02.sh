# r02 ncu evidence (run under gpurun, one GPU): launch list of the bench command + one --set full capture of the five
# kernels of a transformer layer (QKV, attention, out-proj+LN, c_fc, c_proj+LN) at the bench batch (512 images).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-v1}
BENCH="python bench.py --classes 32 --queries 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 9000 --csv --log-file gpurun_out/r02_launches_$TAG.csv $BENCH > gpurun_out/ncu_launch_run.log 2>&1
tail -2 gpurun_out/ncu_launch_run.log
gzip -f gpurun_out/r02_launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_pair_kernel|gemm_tn_rowln_kernel|attention_kv_kernel" \
    -s 133 -c 6 -o gpurun_out/prof_layer_r02_$TAG -f python tools/profile_layer.py > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log
ls -la gpurun_out | tail -4
