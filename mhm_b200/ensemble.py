"""Sharding of an ensemble / calibration sweep over the GPUs of one box.

Members (= parameter sets = `mhm_eval(parameterset)` evaluations, mHM/mo_mhm_eval.f90:94) are
independent, so they are dealt to ranks in contiguous blocks and the time loop needs no
collective; the only exchange is the gather of each member's gauge series
`mRM_runoff(nTimeSteps, nGauges)` on rank 0 -- the B200 analogue of the reference's MPI
master/worker `MPI_Send/Recv` of objective terms (common/mo_common_MPI_tools.F90:39-69).
"""
import numpy as np


def partition_members(n_members, world_size):
    """contiguous member ranges per rank: [(first, count), ...]; earlier ranks take the remainder"""
    base, rem = divmod(n_members, world_size)
    out, first = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < rem else 0)
        out.append((first, cnt))
        first += cnt
    return out


def gather_runoff(local_q, n_members, dist=None, device=None):
    """local_q: numpy (local_members, nGauges, nSteps) of this rank; returns on rank 0 the
    (n_members, nGauges, nSteps) array of the whole ensemble (None on other ranks).
    `dist` is torch.distributed (initialised) or None for a single process; with the NCCL backend
    pass the rank's cuda device, with gloo leave device None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert local_q.shape[0] == n_members
        return local_q
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    parts = partition_members(n_members, world)
    assert local_q.shape[0] == parts[rank][1], (local_q.shape, parts[rank])
    mx = max(c for _, c in parts)
    pad = np.zeros((mx,) + local_q.shape[1:], dtype=np.float64)
    pad[: local_q.shape[0]] = local_q
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([bufs[r][: parts[r][1]].cpu().numpy() for r in range(world)], axis=0)
