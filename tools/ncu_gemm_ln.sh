cd $GRAFT_REPO_ROOT
for c in vit-qkv-pair vit-qkv-pair-ln; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_pair_kernel -s 5 -c 1 -o gpurun_out/prof_$c -f ./build/test_gemm only $c > gpurun_out/ncu_$c.log 2>&1
  tail -1 gpurun_out/ncu_$c.log
done
