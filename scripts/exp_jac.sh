timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench_ik
r = bench_ik.run(torch.device("cuda", 0), 0, 1, lambda x: x, torch.cuda.synchronize)
print(json.dumps({k: r[k] for k in ("jacobian", "mosh_direct")}))
PY
