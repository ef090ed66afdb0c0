"""Adds c4_alt_* to tests/golden/ref_ik_configs.npz: the configs[3] loop (4 frames x 10 iterations, VPoser latent state) of
the SAME compiled reference run with ONE libtorch thread instead of all of them - how far two runs of the reference itself
drift apart in this mode (see c3_alt_* in make_ref_golden_ik_configs.py; the free-running GPU test uses it as its band).

    python tests/golden/make_ref_golden_c4_alt.py        (a few minutes)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_lib  # noqa: E402
from smplpp_b200 import synth  # noqa: E402
from make_ref_golden_ik_configs import MOTION, marker_residual  # noqa: E402


def main():
    t0 = time.time()
    path = os.path.join(OUT, "ref_ik_configs.npz")
    g = dict(np.load(path))
    ref = ref_lib.RefSMPL(ref_lib.model_json_path(0))
    vp = ref_lib.RefVPoser(ref_lib.vposer_json_path(1))
    face_idx, vw0, b0 = g["face_idx"], g["vertex_weights_in"], g["beta"]
    target4, valid4, xv0 = g["c4_target"], g["c4_valid"], g["c4_theta_in"]
    F4, K4 = g["c4_residual"].shape
    n = vw0.shape[0]

    def run():
        th4 = np.zeros((F4, K4, 44), np.float32)
        res4 = np.zeros((F4, K4))
        for f in range(F4):
            x, w = xv0.copy(), vw0.copy()
            for k in range(K4):
                r = ref.ik_iteration(x, b0, face_idx, w, target4[f], pos_task_weight=valid4[f].astype(np.float64), vposer=vp,
                                     **MOTION)
                x, w = r["theta_state"], r["vertex_weights"]
                th4[f, k], res4[f, k] = x, marker_residual(r["e"], valid4[f])
            print("frame %d: residual %.5f -> %.6f m  (%.0f s)" % (f, res4[f, 0], res4[f, -1], time.time() - t0), flush=True)
        return th4, res4

    nthreads = ref_lib.get_num_threads()
    th_full, res_full = run()
    print("all threads vs the stored golden: residual deviation %.3g m" % np.abs(res_full - g["c4_residual"]).max(), flush=True)
    ref_lib.set_num_threads(1)
    th_alt, res_alt = run()
    ref_lib.set_num_threads(nthreads)
    print("reference vs itself (1 thread vs %d): max residual deviation %.3g m (relative %.3g), max state deviation %.3g"
          % (nthreads, np.abs(res_alt - g["c4_residual"]).max(), (np.abs(res_alt - g["c4_residual"]) / g["c4_residual"]).max(),
             np.abs(th_alt - g["c4_theta_traj"]).max()), flush=True)
    g.update(c4_alt_residual=res_alt, c4_alt_theta_traj=th_alt, c4_rerun_residual=res_full)
    np.savez_compressed(path, **g)


if __name__ == "__main__":
    main()
