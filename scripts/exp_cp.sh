timeout 600 python -m pytest tests/test_closest_gpu.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, '.')
from smplpp_b200 import api, synth
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
B = 4096
beta, theta = synth.make_forward_inputs(B, 11)
smpl.launch(beta, theta)
v = smpl.getVertex()
pts = v[:, ::168][:, :41].contiguous() + 0.01
for _ in range(2): smpl.projectPoints(pts, v)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): smpl.projectPoints(pts, v)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("closest points: B=%d n=41 F=13776: %.3f ms -> %.0f frames/s, %.2f G point-triangle pairs/s, %.0f GB/s of vertices read" % (B, ms, B / ms * 1e3, B * 41 * 13776 / ms / 1e6, B * 82680 / ms / 1e6))
PY
