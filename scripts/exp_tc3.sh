for r in 0 1 2 3; do echo "== ring $r";  SMPLPP_TC3_RING=$r timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep "^variant [56]:"; done
SMPLPP_TC3_DBG=1 timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep -A3 "tc3 dbg" | head -4 | cut -c1-330
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -3
