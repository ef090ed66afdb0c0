"""Quick check + timing of the persistent tcgen05 kernel (variant 6) against the FFMA kernel (variant 1) and K2'' (5).
SMPLPP_TC3_DBG=1 prints per-item MMA-thread timestamps of two CTAs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from smplpp_b200 import api, capi, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
beta, theta = synth.make_forward_inputs(B, 11)
out = {}
for var in (1, 5, 6):
    capi.check(capi.lib().smplpp_set_forward_variant(var))
    smpl.launch(beta, theta)
    torch.cuda.synchronize()
    out[var] = smpl.getVertex()
    if os.environ.get("SMPLPP_TC3_DBG"):
        continue
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        smpl.launch(beta, theta)
    a.record()
    for _ in range(20):
        smpl.launch(beta, theta)
    b.record()
    torch.cuda.synchronize()
    print("variant %d: %.4f ms / launch (K1 + K2), %.2f M meshes/s" % (var, a.elapsed_time(b) / 20, B / (a.elapsed_time(b) / 20) / 1e3))
for var in (5, 6):
    d = (out[var] - out[1]).abs()
    print("variant %d vs FFMA: max |dv| = %.3g m, finite = %s" % (var, float(d.max()), bool(torch.isfinite(out[var]).all())))
    if float(d.max()) > 1e-5:
        bad = (d.amax(dim=2) > 1e-5)
        fr = bad.any(dim=1).nonzero().flatten()
        vt = bad.any(dim=0).nonzero().flatten()
        print("  bad frames:", fr[:20].tolist(), "... count", fr.numel(), "| bad vertices:", vt[:20].tolist(), "... count", vt.numel())
