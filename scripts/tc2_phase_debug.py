import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from smplpp_b200 import api, capi, synth
params = synth.make_smpl_params(0)
smpl = api.SMPL(params, device="cuda:0")
beta, theta = synth.make_forward_inputs(4096, 11)
capi.check(capi.lib().smplpp_set_forward_variant(5))
for _ in range(2):
    smpl.launch(beta, theta)
torch.cuda.synchronize()
