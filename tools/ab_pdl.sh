# A/B: programmatic dependent launch on/off. Run under gpurun.
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pdl in 0 1 0 1; do
  OVMR_PDL=$pdl python bench.py --classes 192 --queries 6144 --batch 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl $pdl batch', d['config']['batch'], 'img/s %.0f' % d['value'], 'gemm TF %.0f' % d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], 'prof ms', d['roofline']['ms_per_step_with_events'], 'ms', d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
