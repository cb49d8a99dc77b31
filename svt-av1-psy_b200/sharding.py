"""Multi-GPU host logic.  Frames / GOPs shard across ranks with no data-path collective;
the one exchange of the path is the reconstructed, filtered reference picture travelling from the rank
that produced it to the ranks that will encode pictures depending on it (rest_process.c:663,735-745 ->
PictureDemux EB_PIC_REFERENCE): point-to-point, owner -> consumers, never "everyone gets everything".
Backend-agnostic (torch.distributed: nccl on GPUs, gloo in the CPU tests); used by bench.py --gpus N."""


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


def reference_consumers(rank, world, fanout=2):
    """ranks that encode pictures referencing the picture `rank` just reconstructed.  With pictures dealt
    round-robin, the next `fanout` pictures in coding order (hierarchical-B: a picture is referenced by its
    neighbours of the next temporal layer) live on the next `fanout` ranks."""
    return [(rank + k) % world for k in range(1, min(fanout, world - 1) + 1)]


def reference_producers(rank, world, fanout=2):
    """ranks whose reconstructed pictures this rank needs (inverse of reference_consumers)"""
    return [(rank - k) % world for k in range(1, min(fanout, world - 1) + 1)]


class ReconExchange:
    """Owner -> consumers exchange of reconstructed reference pictures, batched per mini-GOP.

    post(frames) sends each of this rank's `frames` (flat tensors, one per picture of the batch) to its
    consumers and receives the producers' pictures of the same batch into per-slot buffers, as ONE group of
    point-to-point operations (one NCCL group launch on the caller's current -- communication -- stream).
    Returns the work handles; wait(works) makes the current stream wait for their completion."""

    def __init__(self, dist, rank, world, like, batch, fanout=2):
        import torch
        self.dist, self.rank, self.world = dist, rank, world
        self.consumers = reference_consumers(rank, world, fanout)
        self.producers = reference_producers(rank, world, fanout)
        self.batch = batch
        # received reference pictures: [slot in batch][producer]
        self.inbox = [[torch.empty_like(like) for _ in self.producers] for _ in range(batch)]
        self.bytes_sent_per_frame = like.numel() * like.element_size() * len(self.consumers)

    def post(self, frames):
        dist = self.dist
        assert len(frames) <= self.batch
        ops = []
        # identical order on every rank: slot-major, then sends before receives (one group, so no ordering deadlock)
        for k, t in enumerate(frames):
            for c in self.consumers:
                ops.append(dist.P2POp(dist.isend, t, c))
            for j, p in enumerate(self.producers):
                ops.append(dist.P2POp(dist.irecv, self.inbox[k][j], p))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def wait(works):
        for w in works:
            w.wait()


def max_over_ranks(dist, values, device="cpu"):
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]
