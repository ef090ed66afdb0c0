# ncu source-level capture of the forward kernel with the variant selected by SMPLPP_TC3_RING (argument 1)
cat > /tmp/prof_fwd.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from smplpp_b200 import api, synth
dev = torch.device("cuda", 0)
smpl = api.SMPL(synth.make_smpl_params(0), device=dev)
beta_h, theta_h = synth.make_forward_inputs(4096, 11)
for _ in range(6):
    smpl.launch(beta_h, theta_h)
torch.cuda.synchronize()
PY
SMPLPP_TC3_RING=$1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_skin_tc3" -s 3 -c 1 -o gpurun_out/prof_fwd_r$1 -f python /tmp/prof_fwd.py > gpurun_out/ncu_fwd_r$1.log 2>&1
tail -3 gpurun_out/ncu_fwd_r$1.log
