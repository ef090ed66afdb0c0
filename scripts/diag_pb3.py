"""J of 2000 random frames by both pose-blend variants, all IK modes; run-to-run determinism of the tensor-core path."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_ik
from smplpp_b200 import api, capi, synth
dev = torch.device("cuda", 0)
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device=dev)
_, face_idx, _ = synth.make_marker_tasks(params)
ts = api.IkTaskSet(smpl, face_idx, vposer=api.VPoserDecoder(synth.make_vposer_params(1)))
F = 2000
prob = bench_ik.make_problem(smpl, ts, F, 3, dev)
gen = torch.Generator(device=dev).manual_seed(5)
for name, kw, D in (("direct", dict(), 75), ("direct+normal rows", dict(normal_task_weight=0.5), 75), ("beta", dict(optimize_beta=1, enable_qp=0), 75), ("vposer", dict(enable_vposer=1), 44),
                    ("no normals", dict(normal_offset=0.0), 75)):
    opt = api.ik_options(update_state=0, **kw)
    x0 = prob["x0"][:, :D].clone() if D == 75 else torch.cat([prob["x0"][:, :6], 0.5 * torch.randn((F, 32), generator=gen, device=dev), prob["x0"][:, 69:75]], 1).contiguous()
    x0 = x0 + 0.05 * torch.randn(x0.shape, generator=gen, device=dev)
    beta = prob["beta"][None].repeat(F, 1).contiguous() if kw.get("optimize_beta") else prob["beta"]
    outs = {}
    tn = torch.nn.functional.normalize(torch.randn((F, ts.n, 3), generator=gen, device=dev), dim=-1) if kw.get("normal_task_weight") else None
    for v in (421, 420, 420):
        capi.check(capi.lib().smplpp_set_forward_variant(v))
        th, vw = x0.clone(), prob["w0"].clone()
        status, out = ts.step(opt, th, beta, vw, prob["target"], target_normal=tn, pos_task_weight=prob["valid"], outputs=True) if tn is not None else ts.step(opt, th, beta, vw, prob["target"], pos_task_weight=prob["valid"], outputs=True)
        outs.setdefault(v, []).append((out["J"].cpu().numpy().astype(np.float64), out["delta"].cpu().numpy()))
    Jf, Jt, Jt2 = outs[421][0][0], outs[420][0][0], outs[420][1][0]
    d = np.abs(Jt - Jf).reshape(F, -1).max(1) / np.abs(Jf).reshape(F, -1).max(1)
    print("%-20s J tc vs ffma: worst frame relative %.3g (frame %d), median %.3g; tc run-to-run identical: %s; delta max diff %.3g"
          % (name, d.max(), d.argmax(), np.median(d), np.array_equal(Jt, Jt2), np.abs(outs[420][0][1] - outs[421][0][1]).max()))
