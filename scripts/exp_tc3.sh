for r in 0 1 2; do echo "== mode $r";  SMPLPP_TC3_RING=$r timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep "^variant 6"; done
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -3
