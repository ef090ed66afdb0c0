"""CPU suite for the N > 1 path: world_size-2 gloo process groups exercise the frame sharding, the shared-beta
all-reduce (one 111-double message) and the final gather (smplpp_b200/parallel.py).  The per-rank Schur reduction is
restated in numpy here; on the GPU box it is smplpp_ik_shared_beta_reduce (tests/test_ik_gpu.py covers that kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smplpp_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(F=7, D=6, seed=3):
    """Random damped normal equations with the arrow structure of the shared-beta stage."""
    rng = np.random.default_rng(seed)
    n = D + 10
    J = rng.normal(size=(F, 3 * n, n))
    e = rng.normal(size=(F, 3 * n)) * 0.05
    A = np.einsum("fri,frj->fij", J, J) + 1e-3 * np.eye(n)
    b = np.einsum("fri,fr->fi", J, e)
    return A, b, (e ** 2).sum(1)


def _solve_shared(msg, limit=0.5):
    """The replicated 10-dim box QP: min 1/2 x'Sx + r'x, |x| <= limit, with the LM term ||e||^2 on the diagonal."""
    from oracle import smpl_oracle as so
    S = msg[:100].reshape(10, 10) + msg[110] * np.eye(10)
    return so.solve_box_qp(S, msg[100:110], -limit * np.ones(10), limit * np.ones(10))


def _worker(rank, world, port, F, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, b, e2 = _problem(F)
        start, count = parallel.frame_block(F, rank, world)
        sl = slice(start, start + count)
        msg = torch.from_numpy(parallel.schur_message(A[sl], b[sl], e2[sl]))
        parallel.all_reduce_shared_beta(msg)
        dbeta = _solve_shared(msg.numpy())
        x_local = torch.from_numpy(parallel.back_substitute(A[sl], b[sl], dbeta))
        x_all = parallel.gather_frames(x_local, F)
        status = parallel.gather_frames(torch.full((count,), rank, dtype=torch.int32), F)
        if rank == 0:
            out.put((msg.numpy().copy(), dbeta, x_all.numpy().copy(), status.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_frame_block_partition():
    for total in (0, 1, 7, 4096, 1 << 20):
        for world in (1, 2, 3, 8):
            blocks = [parallel.frame_block(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        parallel.frame_block(10, 2, 2)


@pytest.mark.parametrize("F", [7, 8])
def test_shared_beta_two_ranks_equals_single_process(F):
    """Ragged (7 = 4 + 3) and even frame blocks over 2 gloo ranks: the reduced message, the shape step and every
    frame's back-substituted step must equal the single-process result over all frames."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, out)) for r in range(2)]
    for p in procs:
        p.start()
    msg, dbeta, x_all, status = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, b, e2 = _problem(F)
    msg1 = parallel.schur_message(A, b, e2)
    assert np.allclose(msg, msg1, rtol=1e-12, atol=1e-12)
    dbeta1 = _solve_shared(msg1)
    assert np.allclose(dbeta, dbeta1, atol=1e-12) and np.abs(dbeta).max() <= 0.5 + 1e-12
    assert np.allclose(x_all, parallel.back_substitute(A, b, dbeta1), atol=1e-10)
    n0 = parallel.frame_block(F, 0, 2)[1]
    assert status.tolist() == [0] * n0 + [1] * (F - n0)


def test_schur_step_equals_joint_solve_when_unconstrained():
    """Without active bounds the Schur route (reduce -> 10-dim solve -> back-substitution) is the exact minimiser of
    the coupled problem sum_f 1/2 [x_f; d]' A_f [x_f; d] + b_f' [x_f; d] (block-arrow system solved directly)."""
    A, b, e2 = _problem(F=5, D=4, seed=9)
    F, n = A.shape[0], A.shape[1]
    D = n - 10
    msg = parallel.schur_message(A, b, np.zeros(F))
    d = -np.linalg.solve(msg[:100].reshape(10, 10), msg[100:110])
    x = parallel.back_substitute(A, b, d)
    big = np.zeros((F * D + 10, F * D + 10))
    rhs = np.zeros(F * D + 10)
    for f in range(F):
        s = slice(f * D, (f + 1) * D)
        big[s, s] = A[f, :D, :D]
        big[s, F * D:] = A[f, :D, D:]
        big[F * D:, s] = A[f, D:, :D]
        big[F * D:, F * D:] += A[f, D:, D:]
        rhs[s] = b[f, :D]
        rhs[F * D:] += b[f, D:]
    sol = -np.linalg.solve(big, rhs)
    assert np.allclose(sol[F * D:], d, atol=1e-9) and np.allclose(sol[:F * D].reshape(F, D), x, atol=1e-9)
