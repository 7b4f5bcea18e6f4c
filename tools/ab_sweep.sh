# A/B: alternating sweep direction on/off, a few batch sizes. Run under gpurun.
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sw in 0 1; do for b in 127 192 256; do
  OVMR_SWEEP=$sw python bench.py --classes 192 --queries 6144 --batch $b --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sweep $sw batch', d['config']['batch'], 'img/s %.0f' % d['value'], 'gemm TF %.0f' % d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done; done
