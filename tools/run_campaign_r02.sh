# r02 measurement campaign on ONE GPU (run under gpurun): smoke, bench lines of BASELINE configs 1-5, training step, reference arm,
# ncu launch list + one-layer capture.  Results land in gpurun_out/ (copied into profiles/ by hand).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-v3}
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
for c in 2 3 5 4 1; do
  st=2; if [ $c = 4 ]; then st=1; fi; if [ $c = 2 ]; then st=3; fi
  timeout 900 python bench.py --config $c --steps $st --warmup 3 > gpurun_out/r02_bench_cfg${c}_n1_$TAG.json 2> gpurun_out/r02_bench_cfg${c}_n1_$TAG.err
  tail -c 200 gpurun_out/r02_bench_cfg${c}_n1_$TAG.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_cfg${c}_n1_$TAG.json").read().strip().splitlines()[-1])
print("cfg${c}", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["kernel_ms_per_step"], round(d["roofline"]["frac"],3), d["parity"]["pass"], d["parity"]["max_abs_dlogit"], d["clocks"]["sm_mhz"], d.get("cpu_baseline",{}).get("value"))
PY
done
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r02_bench_train_n1_$TAG.json 2> gpurun_out/r02_bench_train_n1_$TAG.err
tail -c 600 gpurun_out/r02_bench_train_n1_$TAG.json | cut -c1-600; tail -c 200 gpurun_out/r02_bench_train_n1_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_$TAG.json 2> gpurun_out/r02_bench_reference_$TAG.err
tail -c 500 gpurun_out/r02_bench_reference_$TAG.json
