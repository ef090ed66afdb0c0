# pose-blend columns on tcgen05 (420) against the FFMA phase of ik_jacobian_kernel (421): IK parity tests, IK bench leg
timeout 1200 python -m pytest tests/test_ik_gpu.py tests/test_ik_configs_gpu.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import json, os, sys
import torch
sys.path.insert(0, '.')
import bench_ik
from smplpp_b200 import capi
dev = torch.device("cuda", 0)
for v in (421, 420):
    capi.check(capi.lib().smplpp_set_forward_variant(v))
    r = bench_ik.run(dev, 0, 1, lambda x: x, torch.cuda.synchronize)
    print(v, json.dumps({k: (r[k] if not isinstance(r[k], dict) else {kk: r[k][kk] for kk in r[k] if kk in ("value", "ms_per_iter", "ms_per_step", "frames_ok", "mean_residual_m", "finite")}) for k in r if k in ("mosh_direct", "moshpp_vposer", "shared_beta", "shared_beta_vposer", "e2e")}))
PY
