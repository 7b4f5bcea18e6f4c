cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 3 5; do
  st=2; if [ $c = 2 ]; then st=3; fi
  timeout 900 python bench.py --config $c --steps $st --warmup 3 > gpurun_out/r02_bench_cfg${c}_n1_v4.json 2> gpurun_out/r02_bench_cfg${c}_n1_v4.err
  tail -c 200 gpurun_out/r02_bench_cfg${c}_n1_v4.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_cfg${c}_n1_v4.json").read().strip().splitlines()[-1])
print("cfg${c}", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["kernel_ms_per_step"], round(d["roofline"]["frac"],3), d["parity"]["pass"], d["parity"]["max_abs_dlogit"], d["clocks"]["sm_mhz"], d.get("cpu_baseline",{}).get("value"))
PY
done
bash tools/run_profile_r02.sh v2
