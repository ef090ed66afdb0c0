# per-item cycles of the forward kernel (DBG instantiation) with all / half / a quarter of the SMs: L2 contention test
for g in 148 74 37; do
  echo "== grid $g"
  SMPLPP_TC3_GRID=$g SMPLPP_TC3_DBG=1 timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep -A4 'tc3 dbg\] cta 0' | grep -v 'stage loads\|stage seen' | cut -c1-330
done
