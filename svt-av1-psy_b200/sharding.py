"""Multi-GPU host logic: frames / GOPs shard across ranks with no data-path
collective; the only exchange is the reconstructed-reference all_gather.  Backend-agnostic
(torch.distributed: nccl on GPUs, gloo in the CPU tests)."""


def frames_for_rank(n_frames, rank, world, gop=1):
    """round-robin assignment of GOP-sized groups of frame indices to ranks"""
    out = []
    for g0 in range(0, n_frames, gop):
        if (g0 // gop) % world == rank:
            out.extend(range(g0, min(g0 + gop, n_frames)))
    return out


def whole_job_fps(steps_per_rank, world, max_ms):
    """weak scaling: every rank runs `steps_per_rank` frames, the job takes the slowest rank's time"""
    return world * steps_per_rank / (max_ms / 1e3)


def exchange_recon(dist, local, gathered=None):
    """reconstructed-reference exchange: every rank contributes its filtered frame, receives all"""
    import torch
    world = dist.get_world_size()
    if gathered is None:
        gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered.view(-1), local.reshape(-1))
    return gathered


def max_over_ranks(dist, values, device="cpu"):
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]
