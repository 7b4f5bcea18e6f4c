cd $GRAFT_REPO_ROOT
./build/test_gemm 2>&1 | grep -E "emit|ln-|FAIL|failed|time" | head -40
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in 0 1; do
  OVMR_FOLD_LN=$f python bench.py --classes 192 --queries 6144 --batch 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fold $f img/s %.0f' % d['value'], 'gemm TF %.0f' % d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], 'ms', d['ms_per_step'], 'clk', d['clocks']['sm_mhz'])"
done
