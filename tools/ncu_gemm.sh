cd $GRAFT_REPO_ROOT
for spec in "10 qkv" "25 out" "40 fc" "52 proj"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_pair_kernel -s $1 -c 1 -f -o gpurun_out/prof_gemm_pair_$2 ./build/test_gemm > gpurun_out/ncu_gemm_$2.log 2>&1
  tail -2 gpurun_out/ncu_gemm_$2.log
done
ls -la gpurun_out/*.ncu-rep
