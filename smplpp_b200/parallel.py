"""Host-side logic of the multi-GPU path: one process per GPU, frames sharded as contiguous blocks,
no data-path collective except in the shared-beta stage of MoSh++ (one all-reduce of 111 doubles per iteration) and
the final gather of the per-frame state.  Everything here runs on whatever backend the process group has (NCCL on
the B200 box, gloo in the CPU tests); the per-rank compute is smplpp_ik_shared_beta_reduce / _apply of the C ABI.

Reference: the node estimates the shape on ONE frame (node/node.cpp:652-656, 734-747, 923-928); coupling many frames
through the 10 betas is this framework's batched generalisation, so the block structure is stated here:

    per frame f :  A_ff (D x D), A_fb (D x 10), A_bb,f (10 x 10), b_f (D), b_b,f (10)
    reduced     :  S = sum_f (A_bb,f - A_bf A_ff^-1 A_fb),  r = sum_f (b_b,f - A_bf A_ff^-1 b_f),  e2 = sum_f ||e_f||^2
    message     :  [S (100) | r (10) | e2 (1)] float64  ->  all_reduce(SUM)
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

REDUCED_DOUBLES = 111  # 10 x 10 Schur block, 10 right-hand side entries, sum of squared residuals


def frame_block(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames owned by `rank`: (start, count).  The first total % world ranks get one more."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError("frame_block: bad arguments")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def all_reduce_shared_beta(reduced: torch.Tensor, group=None) -> torch.Tensor:
    """Sums the [S | r | e2] message over the ranks in place (the only collective on the data path)."""
    if reduced.dtype != torch.float64 or reduced.numel() != REDUCED_DOUBLES:
        raise ValueError("shared-beta message must be %d float64 values" % REDUCED_DOUBLES)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(reduced, op=dist.ReduceOp.SUM, group=group)
    return reduced


def gather_frames(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Final gather of a per-frame array (theta state, status ...) whose rows are sharded by frame_block."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [frame_block(total, r, world)[1] for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError("gather_frames: rank %d holds %d rows, expected %d" % (rank, local.shape[0], counts[rank]))
    width = max(counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def schur_message(A: np.ndarray, b: np.ndarray, e2: np.ndarray, n_beta: int = 10) -> np.ndarray:
    """fp64 restatement of what smplpp_ik_shared_beta_reduce sums over its local frames (used by the CPU tests and
    as documentation of the message layout): A (F, D+10, D+10) damped normal matrices with the beta block last,
    b (F, D+10), e2 (F,)."""
    F, n = A.shape[0], A.shape[1]
    D = n - n_beta
    out = np.zeros(REDUCED_DOUBLES)
    for f in range(F):
        Aff, Afb, Abb = A[f, :D, :D], A[f, :D, D:], A[f, D:, D:]
        L = np.linalg.cholesky(Aff)
        Y = np.linalg.solve(L, Afb)
        y = np.linalg.solve(L, b[f, :D])
        out[:100] += (Abb - Y.T @ Y).reshape(-1)
        out[100:110] += b[f, D:] - Y.T @ y
        out[110] += e2[f]
    return out


def back_substitute(A: np.ndarray, b: np.ndarray, dbeta: np.ndarray, n_beta: int = 10) -> np.ndarray:
    """x_f = -A_ff^-1 (b_f + A_fb dbeta) for every local frame (what smplpp_ik_shared_beta_apply does on the GPU)."""
    D = A.shape[1] - n_beta
    return np.stack([-np.linalg.solve(A[f, :D, :D], b[f, :D] + A[f, :D, D:] @ dbeta) for f in range(A.shape[0])])
