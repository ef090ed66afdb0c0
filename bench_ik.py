"""IK leg of bench.py (lives beside it, outside the product package: `cpu_reference_ik` runs oracle/ as the checker /
CPU baseline): frame-iterations/s of the batched MoSh step (BASELINE configs[2]-[4] shaped workloads).

Synthetic 'sample_walk-shaped' mocap is generated ON THE DEVICE from the seed: ground-truth theta(t) (smooth
random walk) -> product forward pass -> 41 marker positions (15 mm normal offset) + 1 mm noise + 3.4 % dropout.
Every frame starts from the common initial pose (batched frames cannot warm-start from their predecessor)."""
from __future__ import annotations

import ctypes as C
import glob
import json
import os
import time

import numpy as np
import torch

from smplpp_b200 import api, capi, synth

ROOT = os.path.dirname(os.path.abspath(__file__))
# algorithmic FLOPs per frame-iteration of the fused IK step with the 15 mm normal offset (sparse
# forward + Jacobian over the ~600 ring vertices ~3 M, J'J 3M.D(D+1) ~0.7 M, Cholesky D^3/3 ~0.14 M), and the 63x32
# forward-mode Jacobian of the VPoser decoder (2.512.512.32 dominates)
FLOPS_DIRECT = 4.0e6
FLOPS_VPOSER_JAC = 22.0e6
# no measured fp32 figure in MEASURED_PEAKS.json: nominal CUDA-core peak 148 SM x 128 lanes x 2 x 1.965 GHz
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def ncu_counter(kernel: str, key: str):
    """per-launch counter of `kernel` from the newest committed ncu capture (profiles/*_traffic.json)"""
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f)
            if kernel in t.get(key, {}):
                return float(t[key][kernel]), os.path.basename(path)
        except Exception:
            continue
    return None, None


def ncu_counter_sum(kernels, key: str):
    """sum of a per-launch counter over the kernels of one IK iteration, from the newest capture that holds all of them"""
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f).get(key, {})
            if all(k in t for k in kernels):
                return float(sum(t[k] for k in kernels)), os.path.basename(path)
        except Exception:
            continue
    return None, None


def make_problem(smpl, tasks, frames: int, seed: int, dev):
    _, _, vw0 = synth.make_marker_tasks(smpl._params)
    n = tasks.n
    rng = np.random.default_rng(seed)
    # per-frame random poses around the walking clip statistics (cheap to generate for millions of frames)
    base = synth.make_motion(min(frames, 4096), seed)
    beta = (rng.normal(size=10) * 0.5).astype(np.float32)
    w0 = torch.as_tensor(vw0[None], device=dev).repeat(frames, 1, 1).contiguous()
    target = torch.empty((frames, n, 3), dtype=torch.float32, device=dev)
    gt = None
    for s in range(0, frames, 4096):  # chunk by chunk: a million frames never exist as one host array
        e = min(frames, s + 4096)
        chunk = base[: e - s].copy()
        chunk[:, 2:] += rng.normal(scale=0.02, size=(e - s, 23, 3)).astype(np.float32)
        if gt is None:
            gt = chunk
        smpl.launch(beta, chunk)
        target[s:e] = tasks.positions(smpl._vertices, w0[s:e], 0.015)
    gen = torch.Generator(device=dev).manual_seed(seed)
    target += 1e-3 * torch.randn(target.shape, generator=gen, device=dev)
    valid = (torch.rand((frames, n), generator=gen, device=dev) >= 0.034).float()
    target *= valid.unsqueeze(-1)  # node.cpp:682-683: missing marker -> weight 0, target zeroed
    x0 = torch.as_tensor(gt[:1].reshape(1, 75), device=dev).repeat(frames, 1).contiguous()
    return dict(beta=torch.as_tensor(beta, device=dev), w0=w0, target=target.contiguous(), valid=valid.contiguous(),
                x0=x0, gt=gt)


def time_steps(fn, iters, warmup, barrier, max_over_ranks, dev):
    stream = torch.cuda.current_stream(dev)
    for _ in range(warmup):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(iters):
        fn()
    b.record(stream)
    barrier()
    return max_over_ranks(a.elapsed_time(b)) / iters


def run(dev, rank, world, max_over_ranks, barrier, frames: int = 16384, iters: int = 10, warmup: int = 3,
        frames_total: int = 0):
    """frames = frames of THIS rank.  frames_total > 0: a strong-scaling run (configs[4]) whose ranks hold unequal
    contiguous blocks; values are then total frames / max-over-ranks time."""
    params = synth.make_smpl_params(0)
    smpl = api.SMPL(params, device=dev)
    vposer = api.VPoserDecoder(synth.make_vposer_params(1), device=dev)
    _, face_idx, _ = synth.make_marker_tasks(params)
    tasks = api.IkTaskSet(smpl, face_idx, vposer=vposer)
    prob = make_problem(smpl, tasks, frames, 20 + rank, dev)
    out = {"unit": "frame-iters/s", "frames_per_gpu": frames, "markers": tasks.n, "task_vertices": tasks.vertex_count,
           "iterations_timed": iters}
    all_frames = frames_total if frames_total > 0 else world * frames
    if frames_total > 0:
        out["frames_total"] = frames_total
        out["scaling"] = "strong"

    # (1) MoSh direct (configs[2]): theta + translation per frame (D = 75), fixed beta, normal offset 15 mm
    opt = api.ik_options()
    theta, vw = prob["x0"].clone(), prob["w0"].clone()

    def step_direct():
        tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])

    ms = time_steps(step_direct, iters, warmup, barrier, max_over_ranks, dev)
    out["mosh_direct"] = {"value": all_frames / (ms * 1e-3), "ms_per_iter": ms, "unknowns_per_frame": 75}
    # residual after the timed iterations (sanity: the solver is converging on real work)
    status, o = tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"], outputs=False), None
    del o
    out["mosh_direct"]["frames_ok"] = int((status == 0).sum().item())
    # roofline of the IK step: CUDA-core fp32 work (not HBM, not tensor) per SURVEY 8(d)
    ach = FLOPS_DIRECT * frames / (ms * 1e-3) / 1e12
    l2b, l2src = ncu_counter_sum(kernel_names(), "l2_to_sm_bytes_per_launch")
    drb, _ = ncu_counter_sum(kernel_names(), "bytes_per_launch")
    out["roofline"] = {"kernel": " + ".join(kernel_names()),
                       "bound": "fp32 (CUDA-core FFMA / issue and latency; the fp16 / fp64 tensor-pipe parts are minor; neither hbm nor tensor)",
                       "achieved": ach, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS,
                       "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no fp32 figure)",
                       "algorithmic_flops_per_frame_iter": FLOPS_DIRECT, "ms_per_launch": ms, "frames_per_launch": frames,
                       "traffic": drb, "l2_to_sm_bytes_per_launch": l2b, "traffic_source": l2src,
                       "hbm_state_bytes_per_frame_iter": 2 * 4 * 75 + 16 * tasks.n}

    # the Jacobian getter alone (smplpp_ik_jacobian, SURVEY 8(d) "Jacobian materialisation"): algorithmic bytes per frame =
    # the 3 position rows of every marker x D columns in fp32 (4 . 3M . D = 36 900 B at M = 41, D = 75) against the HBM peak.
    # What it really writes is (B, 4n, dim) with the normal-task row; what binds it is the CUDA-core work of the
    # linearisation, so the fraction is small by construction and reported for completeness.
    jac_buf_theta, jac_buf_vw = prob["x0"].clone(), prob["w0"].clone()

    def jac_only():
        tasks.jacobian(opt, jac_buf_theta, prob["beta"], jac_buf_vw, prob["target"], pos_task_weight=prob["valid"])

    ms_j = time_steps(jac_only, max(3, iters // 2), 2, barrier, max_over_ranks, dev)
    jb = 4 * 3 * tasks.n * 75
    out["jacobian"] = {"api": "smplpp_ik_jacobian (e + J in the reference layout, no solve)", "ms_per_call": ms_j,
                       "frames_per_s": all_frames / (ms_j * 1e-3), "algorithmic_bytes_per_frame": jb,
                       "written_bytes_per_frame": 4 * 4 * tasks.n * int(capi.lib().smplpp_ik_dim(C.byref(opt), C.c_int32(tasks.n))) + 4 * 4 * tasks.n,
                       "achieved_gbs": jb * frames / (ms_j * 1e-3) / 1e9, "bound": "fp32 issue / latency (not hbm)"}

    # e2e through the host-buffer C-ABI call smplpp_ik_solve_host: theta / attachments / targets / marker weights in
    # page-locked host arrays, H2D + ONE iteration + D2H of theta, weights, status, residual inside the timed region
    th_h, vw_h = api.pinned_empty((frames, 75)), api.pinned_empty((frames, tasks.n, 3))
    tg_h, pw_h = api.pinned_empty((frames, tasks.n, 3)), api.pinned_empty((frames, tasks.n))
    th0 = prob["x0"].cpu().numpy()
    vw_h[...], tg_h[...], pw_h[...] = prob["w0"].cpu().numpy(), prob["target"].cpu().numpy(), prob["valid"].cpu().numpy()
    beta_h = prob["beta"].cpu().numpy()
    e2e_iters = 1

    def host_call():
        # theta / attachments are updated in place: successive calls continue the solve from where the last one stopped
        # (resetting them here put a 5 MB host memcpy of bench bookkeeping inside the timed region)
        return tasks.solve_host(opt, e2e_iters, th_h, beta_h, vw_h, tg_h, pos_task_weight=pw_h)

    th_h[...] = th0

    for _ in range(2):
        host_call()
    barrier()
    reps = max(3, iters // 2)
    t0 = time.perf_counter()
    for _ in range(reps):
        st_h, res_h = host_call()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / reps)
    h2d = th_h.nbytes + beta_h.nbytes + vw_h.nbytes + tg_h.nbytes + pw_h.nbytes
    d2h = th_h.nbytes + vw_h.nbytes + 4 * frames + 4 * frames
    out["e2e"] = {"value": all_frames * e2e_iters / e2e_s, "unit": "frame-iters/s", "h2d_bytes_per_step": int(h2d),
                  "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "iterations_per_call": e2e_iters,
                  "api": "smplpp_ik_solve_host (targets of every frame in, theta / attachments / status / residual out), "
                         "page-locked host arrays", "frames_ok": int((st_h == 0).sum()),
                  "mean_residual_m": float(res_h[st_h == 0].mean())}

    # (2) MoSh++ with the VPoser latent prior (configs[3]): D = 44, decoder + 63x32 Jacobian in the step
    optv = api.ik_options(enable_vposer=1)
    xv = torch.zeros((frames, 44), dtype=torch.float32, device=dev)
    xv[:, :6] = prob["x0"][:, :6]
    vw2 = prob["w0"].clone()

    def step_vposer():
        tasks.step(optv, xv, prob["beta"], vw2, prob["target"], pos_task_weight=prob["valid"])

    ms = time_steps(step_vposer, max(3, iters // 2), 2, barrier, max_over_ranks, dev)
    out["moshpp_vposer"] = {"value": all_frames / (ms * 1e-3), "ms_per_iter": ms, "unknowns_per_frame": 44}

    # (3) shared-beta stage (configs[3]/[4]): Schur complement per frame + ONE all-reduce of 111 doubles
    sbeta = torch.zeros(10, dtype=torch.float32, device=dev)
    th3, vw3 = prob["x0"].clone(), prob["w0"].clone()

    def step_shared():
        tasks.shared_beta_step(opt, th3, sbeta, vw3, prob["target"], pos_task_weight=prob["valid"])

    ms = time_steps(step_shared, max(3, iters // 2), 2, barrier, max_over_ranks, dev)
    out["shared_beta"] = {"value": all_frames / (ms * 1e-3), "ms_per_iter": ms,
                          "collective": "all_reduce(sum) of 111 float64 per iteration" if world > 1 else "none (1 GPU)"}
    # (3b) configs[3] proper: VPoser latent prior (D = 44) AND the shared-beta stage
    sbeta_v = torch.zeros(10, dtype=torch.float32, device=dev)
    xv3 = torch.zeros((frames, 44), dtype=torch.float32, device=dev)
    xv3[:, :6] = prob["x0"][:, :6]
    vw4 = prob["w0"].clone()

    def step_shared_vposer():
        tasks.shared_beta_step(optv, xv3, sbeta_v, vw4, prob["target"], pos_task_weight=prob["valid"])

    ms = time_steps(step_shared_vposer, max(3, iters // 2), 2, barrier, max_over_ranks, dev)
    ach_v = (FLOPS_DIRECT + FLOPS_VPOSER_JAC) * frames / (ms * 1e-3) / 1e12
    out["shared_beta_vposer"] = {"value": all_frames / (ms * 1e-3), "ms_per_iter": ms, "unknowns_per_frame": 44,
                                 "finite": bool(torch.isfinite(xv3).all().item() and torch.isfinite(sbeta_v).all().item()),
                                 "algorithmic_tflops": ach_v,
                                 "note": "configs[3]: the per-frame part and the shared-beta system are pinned against the "
                                         "compiled reference / a dense fp64 solve in tests/test_ik_configs_gpu.py"}
    # (4) projection of the task points onto the posed mesh + re-seated face / weights (node.cpp:970-1001):
    # full forward pass of a block of frames, then smplpp_closest_points (41 points x 13776 faces per frame)
    rb = min(frames, 4096)
    smpl.launch(prob["beta"].cpu().numpy(), prob["gt"][:rb])
    verts = smpl._vertices
    pts = tasks.positions(verts, prob["w0"][:rb], 0.015)

    def reproject():
        smpl.projectPoints(pts, verts, want_weights=True)

    ms = time_steps(reproject, 5, 2, barrier, max_over_ranks, dev)
    out["reproject"] = {"value": world * rb / (ms * 1e-3), "unit": "frames/s", "ms_per_call": ms, "frames": rb,
                        "points_per_frame": tasks.n, "faces": int(params.face_indices.shape[0]),
                        "vertex_bytes_read_gbs": rb * 82680 / (ms * 1e-3) / 1e9}
    out["value"] = out["mosh_direct"]["value"]
    return out


def run_config4(dev, rank, world, max_over_ranks, barrier, frames_total: int = 1 << 20, iters: int = 2):
    """BASELINE configs[4]: `frames_total` synthetic mocap frames sharded over the ranks as contiguous blocks (strong
    scaling).  Per-frame IK steps need no exchange; the shared-beta stage all-reduces 111 doubles per iteration (NCCL over
    NVLink when world > 1); the per-frame state is gathered on every rank at the end (parallel.gather_frames)."""
    from smplpp_b200 import parallel
    start, frames = parallel.frame_block(frames_total, rank, world)
    params = synth.make_smpl_params(0)
    smpl = api.SMPL(params, device=dev)
    _, face_idx, _ = synth.make_marker_tasks(params)
    tasks = api.IkTaskSet(smpl, face_idx)
    prob = make_problem(smpl, tasks, frames, 22 + rank, dev)
    opt = api.ik_options()
    theta, vw = prob["x0"].clone(), prob["w0"].clone()

    def step_direct():
        tasks.step(opt, theta, prob["beta"], vw, prob["target"], pos_task_weight=prob["valid"])

    ms_direct = time_steps(step_direct, iters, 1, barrier, max_over_ranks, dev)
    sbeta = torch.zeros(10, dtype=torch.float32, device=dev)
    th2, vw2 = prob["x0"].clone(), prob["w0"].clone()

    def step_shared():
        tasks.shared_beta_step(opt, th2, sbeta, vw2, prob["target"], pos_task_weight=prob["valid"])

    ms_shared = time_steps(step_shared, iters, 1, barrier, max_over_ranks, dev)
    # every rank solved the same 10-dim problem from the same all-reduced message: identical betas
    beta_spread = 0.0
    if world > 1:
        import torch.distributed as dist
        lo, hi = sbeta.clone(), sbeta.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        beta_spread = float((hi - lo).abs().max().item())
    barrier()
    t0 = time.perf_counter()
    full = parallel.gather_frames(th2, frames_total)
    torch.cuda.synchronize()
    gather_s = max_over_ranks(time.perf_counter() - t0)
    ok = bool(full.shape[0] == frames_total and torch.isfinite(full).all().item())
    return {"frames_total": frames_total, "frames_this_rank": frames, "scaling": "strong", "unit": "frame-iters/s",
            "mosh_direct": {"value": frames_total / (ms_direct * 1e-3), "ms_per_iter": ms_direct},
            "shared_beta": {"value": frames_total / (ms_shared * 1e-3), "ms_per_iter": ms_shared,
                            "collective": "all_reduce(sum) of 111 float64 per iteration (NCCL)" if world > 1 else "none (1 GPU)",
                            "beta_spread_over_ranks": beta_spread},
            "final_gather": {"seconds": gather_s, "bytes": int(frames_total * 75 * 4), "gathered_rows": int(full.shape[0]),
                             "finite": ok, "collective": "all_gather over NCCL" if world > 1 else "none (1 GPU)"},
            "value": frames_total / (ms_direct * 1e-3)}


def kernel_names():
    return ("ik_restshape_tc_kernel", "ik_jacobian_kernel", "ik_poseblend_tc_kernel", "ik_solve_mma_kernel")


def cpu_reference_ik(frames: int = 2, iters: int = 2):
    """Reference CPU path (oracle/_ref harness restating node.cpp:753-968 on the compiled reference objects):
    frame-iterations/s on a bounded sample."""
    from oracle import ref_lib
    params = synth.make_smpl_params(0)
    _, face_idx, vw = synth.make_marker_tasks(params)
    ref_lib.set_num_threads(os.cpu_count() or 1)
    ref = ref_lib.RefSMPL(ref_lib.model_json_path(0))
    gt = synth.make_motion(frames + 1, 20)
    beta = np.zeros(10, np.float32)
    n = len(face_idx)
    tgt = np.random.default_rng(0).normal(size=(n, 3)).astype(np.float32) * 0.3
    t0 = time.perf_counter()
    for f in range(frames):
        x = gt[f].reshape(-1).copy()
        w = vw.copy()
        for _ in range(iters):
            r = ref.ik_iteration(x, beta, face_idx, w, tgt, normal_task_weight=0.0, phi_limit=0.0, normal_offset=0.015)
            x, w = r["theta_state"], r["vertex_weights"]
    dt = time.perf_counter() - t0
    return frames * iters / dt, ref_lib.get_num_threads(), (
        "unmodified reference objects (oracle/_ref, libtorch CPU autograd rows) + node.cpp:753-968 restated: %d frames x %d "
        "iterations, 41 markers, direct theta (D=75), 15 mm normal offset" % (frames, iters))
