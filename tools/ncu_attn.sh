cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 -o gpurun_out/prof_attn_tc2 -f python tools/bench_kernels.py > gpurun_out/ncu_attn2.log 2>&1
tail -3 gpurun_out/ncu_attn2.log
