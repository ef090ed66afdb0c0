# skeleton decomposition of blend_skin_tc3_kernel (DBG instantiation): SMPLPP_TC3_DBG = 1 + 2 * mode mask
# 63 = no MMA / TMEM loads / stores; + 32 no stage ring, 64 no transform ring, 128 no epilogue arithmetic, 256 no matrix hand-off, 512 no drain hand-off
for mask in 0 31 63 95 159 287 543 127 255 511 1023 991 96 32 64; do
  m=$((1 + 2 * mask))
  echo "== mask $mask"
  SMPLPP_TC3_DBG=$m timeout 120 python scripts/tc3_debug.py 4096 2>&1 | grep -A4 'tc3 dbg\] cta 74' | grep -v 'stage loads\|stage seen' | cut -c1-400
done
