# forward kernel A/B in one call: smplpp_b200/lib_old.so against lib_new.so (both built here), alternating
run() {
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ik | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.3f M  ms/launch %.4f  burst %.4f  frac %.3f  diff %.2e' % (d['value']/1e6, r['ms_per_launch'], r['burst_ms_per_launch'], r['frac'], d['e2e']['max_abs_diff_vs_device_path']))"
}
for v in old new old new; do
  cp smplpp_b200/lib_$v.so smplpp_b200/libsmplpp_b200.so
  echo "== $v"; run
done
