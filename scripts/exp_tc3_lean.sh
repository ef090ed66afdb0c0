# lean issue loop / epilogue of blend_skin_tc3_kernel: parity + timing of EPI 1 (default) and EPI 2 (SMPLPP_TC3_RING=3)
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -3
for r in 0 3; do
  echo "== SMPLPP_TC3_RING=$r"
  SMPLPP_TC3_RING=$r timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ik | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value', d['value'], 'ms/step', d['ms_per_step'], 'ms/launch', r['ms_per_launch'], 'burst', r['burst_ms_per_launch'], 'frac', r['frac'], 'diff', d['e2e']['max_abs_diff_vs_device_path'])"
  SMPLPP_TC3_RING=$r timeout 100 python scripts/tc3_debug.py 4096 2>&1 | grep 'variant 6'
done
SMPLPP_TC3_DBG=1 timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep -A4 'tc3 dbg\] cta 74' | grep -v 'stage loads\|stage seen'
