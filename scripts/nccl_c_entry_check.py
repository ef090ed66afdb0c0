"""2-GPU check of the C entry smplpp_ik_shared_beta_step(..., ncclComm_t): every rank creates a RAW NCCL communicator
through libnccl's C API (ctypes; the unique id travels over the torch.distributed store), runs the shared-beta stage on
its block of frames through the one-call C entry, and the result is compared with (a) the same frames solved on ONE
rank and (b) the torch.distributed path (reduce -> dist.all_reduce -> apply).   torchrun --nproc-per-node 2 ..."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from smplpp_b200 import api, capi, parallel, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # ---- raw communicator ----
    try:
        nccl = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)  # the copy torch already loaded
    except OSError:
        import nvidia.nccl
        nccl = C.CDLL(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2"), mode=C.RTLD_GLOBAL)

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_ubyte * 128)]  # (a c_char array field would be cut at the first NUL)

    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.frombuffer(bytearray(C.string_at(C.byref(uid), 128)), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, "ncclCommInitRank failed: %d" % rc
    # ---- problem: F frames sharded as contiguous blocks ----
    params = synth.make_smpl_params(0)
    smpl = api.SMPL(params, device=dev)
    _, face_idx, vw0 = synth.make_marker_tasks(params)
    tasks = api.IkTaskSet(smpl, face_idx)
    n, F = tasks.n, 203
    gt = synth.make_motion(F, 23)
    beta_true = np.random.default_rng(6).normal(size=10).astype(np.float32)
    smpl.launch(beta_true, gt)
    w_all = torch.as_tensor(np.repeat(vw0[None], F, axis=0), device=dev).contiguous()
    target_all = tasks.positions(smpl.getVertex(), w_all, 0.015).contiguous()
    x_all = torch.as_tensor(gt.reshape(F, 75) + np.random.default_rng(7).normal(size=(F, 75)).astype(np.float32) * 0.02, device=dev)
    opt = api.ik_options()
    lib = capi.lib()

    def c_entry(theta, sbeta, vw, target, comm_ptr):
        b = theta.shape[0]
        status = torch.empty((b,), dtype=torch.int32, device=dev)
        reduced = torch.empty((111,), dtype=torch.float64, device=dev)
        need = lib.smplpp_ik_shared_beta_workspace_bytes(tasks._h, C.byref(opt), C.c_int64(b))
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        capi.check(lib.smplpp_ik_shared_beta_step(
            smpl.handle, None, tasks._h, C.byref(opt), st, C.c_int64(b), C.c_void_p(theta.data_ptr()), C.c_void_p(sbeta.data_ptr()),
            C.c_void_p(vw.data_ptr()), C.c_void_p(target.data_ptr()), None, C.c_void_p(status.data_ptr()),
            C.c_void_p(reduced.data_ptr()), comm_ptr, C.c_void_p(ws.data_ptr()), C.c_size_t(need)))
        torch.cuda.synchronize()
        return status, reduced

    s0, cnt = parallel.frame_block(F, rank, world)
    # (1) sharded, raw NCCL communicator inside the C entry
    th = x_all[s0:s0 + cnt].clone()
    sb = torch.zeros(10, device=dev)
    st1, red1 = c_entry(th, sb, w_all[s0:s0 + cnt].clone(), target_all[s0:s0 + cnt].contiguous(), comm)
    # (2) all frames on this rank alone (no communicator)
    th_full = x_all.clone()
    sb_full = torch.zeros(10, device=dev)
    st2, red2 = c_entry(th_full, sb_full, w_all.clone(), target_all, None)
    # (3) torch.distributed path
    th3 = x_all[s0:s0 + cnt].clone()
    sb3 = torch.zeros(10, device=dev)
    tasks.shared_beta_step(opt, th3, sb3, w_all[s0:s0 + cnt].clone(), target_all[s0:s0 + cnt].contiguous())
    torch.cuda.synchronize()
    d_beta = float((sb - sb_full).abs().max())
    d_theta = float((th - th_full[s0:s0 + cnt]).abs().max())
    d_red = float(((red1 - red2).abs() / red2.abs().clamp_min(1e-12)).max())
    d_torch = float((sb - sb3).abs().max()), float((th - th3).abs().max())
    gathered = parallel.gather_frames(th, F)
    d_gather = float((gathered - th_full).abs().max())
    ok = (int((st1 != 0).sum()) == 0 and d_beta < 1e-6 and d_theta < 1e-5 and d_red < 1e-9 and d_torch[0] == 0.0 and d_torch[1] == 0.0
          and d_gather < 1e-5 and gathered.shape[0] == F)
    print("rank %d/%d: C entry with ncclComm_t vs single-rank solve: |dbeta| %.2e |dtheta| %.2e, message rel %.1e; vs torch.distributed path "
          "|dbeta| %.1e |dtheta| %.1e; gather_frames over NCCL %d rows |d| %.1e -> %s"
          % (rank, world, d_beta, d_theta, d_red, d_torch[0], d_torch[1], gathered.shape[0], d_gather, "OK" if ok else "FAILED"), flush=True)
    nccl.ncclCommDestroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
