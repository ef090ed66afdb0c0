# wait-time decomposition of blend_skin_tc3_kernel (DBG instantiation): SMPLPP_TC3_DBG = 1 + 2 * mode mask
for m in 1 3 5 9 17 33 39 25 63; do
  echo "== SMPLPP_TC3_DBG=$m"
  SMPLPP_TC3_DBG=$m timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep -A4 'tc3 dbg\] cta 74' | grep -v 'stage loads\|stage seen'
done
