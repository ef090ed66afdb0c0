"""GPU helper: device time of the standalone skinning kernels (TMA pipeline vs register kernel) at B frames."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplpp_b200 import api, capi, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
lib = capi.lib()
V = smpl.vertex_num
g = torch.Generator(device="cuda").manual_seed(0)
rest = torch.randn((B, V, 3), device="cuda", generator=g)
xf34 = torch.randn((B, 24, 3, 4), device="cuda", generator=g)
root = torch.randn((B, 3), device="cuda", generator=g)
out = torch.empty_like(rest)
st = torch.cuda.current_stream()
for var, name in ((202, "tcgen05"), (200, "ffma tma pipeline"), (201, "ffma register kernel")):
    capi.check(lib.smplpp_set_forward_variant(var))
    def run():
        capi.check(lib.smplpp_model_skinning34(smpl.handle, C.c_void_p(st.cuda_stream), C.c_int64(B), api._ptr(rest),
                                               api._ptr(xf34), api._ptr(root), api._ptr(out)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-16s B=%d: %.4f ms  %.0f GB/s algorithmic (166512 B/mesh)" % (name, B, ms, 166512 * B / ms / 1e6))
capi.check(lib.smplpp_set_forward_variant(202))
