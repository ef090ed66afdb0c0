timeout 600 python -m pytest tests/test_vposer_gpu.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, '.')
from smplpp_b200 import api, capi, synth
vp = api.VPoserDecoder(synth.make_vposer_params(1), device="cuda:0")
B = 16384
z = torch.randn(B, 32, device="cuda:0")
for var in (301, 300):
    capi.check(capi.lib().smplpp_set_forward_variant(var))
    for _ in range(2): vp.forward(z, jacobian=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): vp.forward(z, jacobian=True)
    b.record(); torch.cuda.synchronize()
    print("variant %d: %.3f ms per %d frames" % (var, a.elapsed_time(b) / 5, B))
PY
