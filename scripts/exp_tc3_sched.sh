# GEMM 1 schedules (SMPLPP_TC3_RING 0 / 4 / 5) and the L2 prefetch of the first item (SMPLPP_TC3_PREFETCH)
run() {
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ik | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.3f M  ms/launch %.4f  burst %.4f  frac %.3f  diff %.2e' % (d['value']/1e6, r['ms_per_launch'], r['burst_ms_per_launch'], r['frac'], d['e2e']['max_abs_diff_vs_device_path']))"
}
echo "== default (sched 0, prefetch)"; run
echo "== no prefetch"; SMPLPP_TC3_PREFETCH=0 run
echo "== sched 1"; SMPLPP_TC3_RING=4 run
echo "== sched 2"; SMPLPP_TC3_RING=5 run
echo "== sched 1, 3+3 ring n/a; staged stores (ring 2)"; SMPLPP_TC3_RING=2 run
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -3
SMPLPP_TC3_RING=4 timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "golden or edge" 2>&1 | tail -3
