for pr in 0 1; do echo "== pair $pr";  SMPLPP_TC3_PAIR=$pr timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep "^variant 6"; done
