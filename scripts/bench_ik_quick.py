"""IK leg of the bench alone, default two-kernel path (400 = auto) and the fused kernel (402) side by side."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_ik
from smplpp_b200 import capi
dev = torch.device("cuda", 0)
for v in ([400, 402] if "--both" in sys.argv else [400]):
    capi.check(capi.lib().smplpp_set_forward_variant(v))
    r = bench_ik.run(dev, 0, 1, lambda x: x, torch.cuda.synchronize)
    print(v, json.dumps({k: (r[k] if not isinstance(r[k], dict) else {kk: r[k][kk] for kk in r[k] if kk in ("value", "ms_per_iter", "ms_per_step", "frames_ok", "mean_residual_m", "finite")}) for k in r if k in ("mosh_direct", "moshpp_vposer", "shared_beta", "shared_beta_vposer", "e2e")}))
