cd $GRAFT_REPO_ROOT
for b in 96 128 192 256 384 512; do
  python bench.py --classes 192 --queries 6144 --batch $b --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch', d['config']['batch'], 'img/s %.0f' % d['value'], 'gemm TF %.0f' % d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], 'clk', d['clocks']['sm_mhz'])"
done
