# r02 multi-GPU lines (run under gpurun --gpus N): bash tools/run_scale_r02.sh N "2 3 5" [tag]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1; CFGS=$2; TAG=${3:-v4}
for c in $CFGS; do
  st=2; if [ $c = 4 ]; then st=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$c bench.py --gpus $N --config $c --steps $st --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg${c}_n${N}_$TAG.json 2> gpurun_out/r02_bench_cfg${c}_n${N}_$TAG.err
  tail -c 300 gpurun_out/r02_bench_cfg${c}_n${N}_$TAG.err | tail -2
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_cfg${c}_n${N}_$TAG.json").read().strip().splitlines()[-1])
print("cfg${c} N=$N", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["kernel_ms_per_step"], round(d["ms_per_step"],1), d["parity"]["pass"] if d.get("parity") else None)
PY
done
