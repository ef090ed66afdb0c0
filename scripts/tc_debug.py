"""GPU debug helper: error and device time of every forward variant (run on the B200 box)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplpp_b200 import api, capi, synth
from oracle import smpl_oracle as so

params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
beta, theta = synth.make_forward_inputs(B, 11)
v_o, j_o, _, _ = so.forward_numpy(so.SmplModel.from_params(params), beta[:64], theta[:64])
lib = capi.lib()
ref = None
for var, order in ((1, 1), (2, 1), (4, 1), (5, 1)):
    capi.check(lib.smplpp_set_forward_variant(100 + order))
    capi.check(lib.smplpp_set_forward_variant(var))
    try:
        smpl.launch(beta, theta)
        torch.cuda.synchronize()
        v = smpl.getVertex().cpu().numpy()
    except Exception as ex:
        print("variant", var, "FAILED:", ex)
        continue
    if ref is None:
        ref = v
    print("variant %d order %d: max|v - oracle| = %.3g (64 frames), max|v - ffma| = %.3g, finite=%s" % (
        var, order, np.abs(v[:64] - v_o).max(), np.abs(v - ref).max(), np.isfinite(v).all()))
    if not np.isfinite(v).all() or np.abs(v - ref).max() > 1e-4:
        d = np.abs(v - ref).max(axis=2)
        bad = np.argwhere(~(d < 1e-4))
        print("  bad entries:", len(bad), "first", bad[:8].tolist(), "frames", sorted(set(bad[:, 0].tolist()))[:10],
              "verts min/max", bad[:, 1].min(), bad[:, 1].max())
        print("  sample got", v[bad[0][0], bad[0][1]], "want", ref[bad[0][0], bad[0][1]])
    t = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            smpl.launch(beta, theta)
        b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b) / 10)
    print("  launch ms (incl. K1, 10 back to back): min %.3f  -> %.2f M meshes/s" % (min(t), B / min(t) / 1e3))
capi.check(lib.smplpp_set_forward_variant(0))
