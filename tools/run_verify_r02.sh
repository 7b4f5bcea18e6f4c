cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err
tail -c 200 gpurun_out/r02_bench_default_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_default_final.json").read().strip().splitlines()[-1])
print("default", round(d["value"]), round(d["e2e"]["value"]), d["steps"], d["warmup"], d["roofline"]["kernel_ms_per_step"], round(d["roofline"]["frac"],3), d["parity"]["pass"], d["gpu_launches"], d["clocks"])
PY
timeout 600 python bench.py --config 3 --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg3_n1_v5.json 2> gpurun_out/r02_bench_cfg3_n1_v5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_cfg3_n1_v5.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["kernel_ms_per_step"], d["parity"]["pass"], d["gpu_launches"])
PY
